#!/bin/bash
# Round-2: the owner-publishes right-looking Cholesky (ONE GPU):  gpurun --timeout 1200 -- 'bash tools/r2_chol_check.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_config_a.py tests/test_gpu_sampled.py tests/test_gpu_edge.py tests/test_gpu_golden.py -m gpu -q -x > gpurun_out/r2h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_tests.log; tail -4 gpurun_out/r2h_tests.log
B="timeout 120 python bench.py --no-cpu --no-e2e --no-extras"
for cfg in A B8 B D8; do
  steps=20; [ $cfg = A ] && steps=100; [ $cfg = B8 ] && steps=100
  $B --config $cfg --steps $steps > gpurun_out/r2h_${cfg}.json 2>> gpurun_out/r2h_err.log
done
python tools/r2_summary.py gpurun_out/r2h_*.json | tee gpurun_out/r2h_summary.txt
timeout 120 python tools/r2_sampled_profile.py nosetup | tee gpurun_out/r2h_sampled_timing.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2h_launches_B8.csv $B --config B8 --steps 3 --warmup 3 > gpurun_out/r2h_ll.log 2>&1
python tools/launch_summary.py gpurun_out/r2h_launches_B8.csv | tee gpurun_out/r2h_launch_summary_B8.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2h_launches_D8.csv $B --config D8 --steps 3 --warmup 3 > gpurun_out/r2h_ll.log 2>&1
python tools/launch_summary.py gpurun_out/r2h_launches_D8.csv | head -8
