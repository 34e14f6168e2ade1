#!/bin/bash
# Round-2 opener, part 2 (TWO GPUs, charged twice, ~5 min):  gpurun --gpus 2 --timeout 600 -- 'bash tools/r2_two_gpu.sh'
# First hardware run of the sharded sampled path and of the NCCL-free, graph-replayed sharded sweeps (peer_graph).
mkdir -p gpurun_out
export ITCPD_EXPERIMENTAL=1
timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2_multi_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_multi_tests.log; tail -8 gpurun_out/r2_multi_tests.log
for pg in 0 1; do
  ITCPD_BENCH_PHASES=1 ITCPD_PEER_GRAPH=$pg timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 \
      bench.py --gpus 2 --steps 50 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_N2_peergraph$pg.json 2>> gpurun_out/r2_err2.log
done
python tools/r2_summary.py gpurun_out/r2_N2_*.json | tee gpurun_out/r2_summary_two_gpu.txt
