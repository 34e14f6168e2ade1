#!/bin/bash
# Round-2: the full driver-style bench line at N = 1 (extras included), the reference arm, and the early-pass-B A/B
#   gpurun --timeout 1500 -- 'bash tools/r2_bench_check.sh'
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r2c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_tests.log; tail -4 gpurun_out/r2c_tests.log
B="timeout 120 python bench.py --no-cpu --no-e2e --no-extras"
for cfg in B B8 A; do
  steps=20; [ $cfg != B ] && steps=100
  $B --config $cfg --steps $steps > gpurun_out/r2c_${cfg}_default.json 2>> gpurun_out/r2c_err.log
  ITCPD_EARLY_B=1 $B --config $cfg --steps $steps > gpurun_out/r2c_${cfg}_earlyb.json 2>> gpurun_out/r2c_err.log
done
python tools/r2_summary.py gpurun_out/r2c_*.json | tee gpurun_out/r2c_summary.txt
( time timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_full_bench.json 2> gpurun_out/r2c_full_bench.err ) 2> gpurun_out/r2c_full_bench.time
tail -3 gpurun_out/r2c_full_bench.time; tail -5 gpurun_out/r2c_full_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c_full_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d.get("e2e"), "\nparity", d.get("parity"), "\ncpu", d.get("cpu_baseline"))
for r in d.get("extra", []):
    print(json.dumps(r)[:900])
PY
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2c_reference.json 2> gpurun_out/r2c_reference.err ) 2> gpurun_out/r2c_reference.time
tail -3 gpurun_out/r2c_reference.time; cat gpurun_out/r2c_reference.json | cut -c1-1500
