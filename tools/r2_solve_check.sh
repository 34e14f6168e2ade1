#!/bin/bash
# Round-2: explicit-inverse row solve A/B (ONE GPU):  gpurun --timeout 1200 -- 'bash tools/r2_solve_check.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_config_a.py tests/test_gpu_sampled.py tests/test_gpu_edge.py tests/test_gpu_golden.py tests/test_julia_golden.py -m gpu -q -x > gpurun_out/r2i_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_tests.log; tail -4 gpurun_out/r2i_tests.log
B="timeout 120 python bench.py --no-cpu --no-e2e --no-extras"
for cfg in A B8 B; do
  steps=20; [ $cfg = A ] && steps=100; [ $cfg = B8 ] && steps=100
  $B --config $cfg --steps $steps > gpurun_out/r2i_${cfg}_solve1.json 2>> gpurun_out/r2i_err.log
  ITCPD_SOLVE=0 $B --config $cfg --steps $steps > gpurun_out/r2i_${cfg}_solve0.json 2>> gpurun_out/r2i_err.log
done
python tools/r2_summary.py gpurun_out/r2i_*.json | tee gpurun_out/r2i_summary.txt
python - <<'PY'
import json
for f in ("A", "B8", "B"):
    d = json.loads(open(f"gpurun_out/r2i_{f}_solve1.json").read().strip().splitlines()[-1])
    print(f, "parity", d.get("parity"))
PY
timeout 120 python tools/r2_sampled_profile.py nosetup | tee gpurun_out/r2i_sampled_timing.jsonl
ITCPD_SOLVE=0 timeout 120 python tools/r2_sampled_profile.py nosetup | tee gpurun_out/r2i_sampled_timing_solve0.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2i_launches_B8.csv $B --config B8 --steps 3 --warmup 3 > gpurun_out/r2i_ll.log 2>&1
python tools/launch_summary.py gpurun_out/r2i_launches_B8.csv > gpurun_out/r2i_launch_summary_B8.txt; cat gpurun_out/r2i_launch_summary_B8.txt
