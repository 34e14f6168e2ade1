#!/bin/bash
# Round-2 final check (ONE GPU), what the driver runs at round end:  gpurun --timeout 1500 -- 'bash tools/r2_final_check.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_tests.log; tail -5 gpurun_out/r2g_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_smoke.log; tail -2 gpurun_out/r2g_smoke.log
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_full_bench.json 2> gpurun_out/r2g_full_bench.err ) 2> gpurun_out/r2g_full_bench.time
tail -3 gpurun_out/r2g_full_bench.time; tail -3 gpurun_out/r2g_full_bench.err
python tools/r2_summary.py gpurun_out/r2g_full_bench.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2g_full_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("pageable"), "parity", d["parity"]["max_abs_dfit"], d["parity"]["ok"])
for r in d.get("extra", []):
    if r["config"] == "E":
        for x in r["results"]: print("   E", x["alg"], round(x["ms_per_sweep"], 3), "ms/sweep fit", round(x["fit"], 4), x.get("setup_s"))
    else:
        print("  ", r["config"], r.get("value"), (r.get("parity") or {}).get("max_abs_dfit"), (r.get("roofline") or {}).get("frac"), r.get("error"))
PY
for cfg in B8 A; do timeout 120 python bench.py --no-cpu --no-e2e --no-extras --config $cfg --steps 100 > gpurun_out/r2g_${cfg}.json 2>> gpurun_out/r2g_err.log; done
python tools/r2_summary.py gpurun_out/r2g_B8.json gpurun_out/r2g_A.json
timeout 120 python tools/r2_sampled_profile.py nosetup | tee gpurun_out/r2g_sampled_timing.jsonl
