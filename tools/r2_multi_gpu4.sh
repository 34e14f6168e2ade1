#!/bin/bash
# Round-2: wider peer small all-reduce (8 GPUs, short):  gpurun --gpus 8 --timeout 600 -- 'bash tools/r2_multi_gpu4.sh'
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2q_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_tests.log; tail -3 gpurun_out/r2q_tests.log
for rep in 1 2; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29681 \
      bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/r2q_N8_$rep.json 2>> gpurun_out/r2q_err.log
done
ITCPD_BENCH_PHASES=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29681 \
      bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/r2q_N8_phases.json 2>> gpurun_out/r2q_err.log
python tools/r2_summary.py gpurun_out/r2q_*.json | tee gpurun_out/r2q_summary.txt
tail -3 gpurun_out/r2q_err.log
