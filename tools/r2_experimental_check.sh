#!/bin/bash
# Kept for the old name: the round-2 opener is split so that only the multi-GPU part is charged for two GPUs.
#   gpurun --timeout 1200 -- 'bash tools/r2_single_gpu.sh'           (INT8 probe + tests, early_pass_b, chol_alg=2, A/B bench lines)
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/r2_two_gpu.sh'      (sharded sampled path, peer_graph)
bash "$(dirname "$0")/r2_single_gpu.sh"
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then bash "$(dirname "$0")/r2_two_gpu.sh"; fi
