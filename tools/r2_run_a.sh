#!/bin/bash
# Round-2 run A (ONE GPU): sampled path after the kernel fixes, INT8 dynamic-range tests, full bench line with every extra record,
# ncu evidence (full capture of the INT8 contraction and of the fused sampled kernel; launch list of a default config-B run)
#   gpurun --timeout 1500 -- 'bash tools/r2_run_a.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampled.py tests/test_gpu_dense.py -m gpu -q -x -k "sampled or leverage or i8 or generator or per_hook or graph" > gpurun_out/r2f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_tests.log; tail -5 gpurun_out/r2f_tests.log
timeout 300 python tools/r2_sampled_profile.py > gpurun_out/r2f_sampled_timing.jsonl 2> gpurun_out/r2f_err.log; cat gpurun_out/r2f_sampled_timing.jsonl
for cfg in B8 A; do
  timeout 120 python bench.py --no-cpu --no-e2e --no-extras --config $cfg --steps 100 > gpurun_out/r2f_${cfg}_default.json 2>> gpurun_out/r2f_err.log
done
( time timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_full_bench.json 2> gpurun_out/r2f_full_bench.err ) 2> gpurun_out/r2f_full_bench.time
tail -3 gpurun_out/r2f_full_bench.time; tail -3 gpurun_out/r2f_full_bench.err
python tools/r2_summary.py gpurun_out/r2f_*.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2f_full_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("pageable"))
for r in d.get("extra", []):
    if r["config"] == "E":
        for x in r["results"]: print("   E", json.dumps(x)[:300])
    else:
        print("  ", r["config"], r.get("value"), r.get("parity", {}).get("max_abs_dfit") if r.get("parity") else None, r.get("roofline", {}).get("frac"), r.get("error"))
PY
# ncu: launch list of the default path (shares), full capture of the INT8 contraction (opt-in) and of the sampled MTTKRP kernel
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches_configB.csv $CMD > gpurun_out/r2f_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r2f_launches_configB.csv > gpurun_out/r2f_launch_summary_configB.txt 2>&1; cat gpurun_out/r2f_launch_summary_configB.txt
ITCPD_GEMM_I8=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:partial_gemm_i8 -s 4 -c 2 -o gpurun_out/r2f_prof_i8 $CMD > gpurun_out/r2f_ncu_full_i8.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:partial_gemm_kernel -s 4 -c 2 -o gpurun_out/r2f_prof_dmma $CMD > gpurun_out/r2f_ncu_full_dmma.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sampled_mttkrp -s 6 -c 1 -o gpurun_out/r2f_prof_sampled python tools/r2_sampled_profile.py nosetup > gpurun_out/r2f_ncu_full_sampled.log 2>&1
for f in i8 dmma sampled; do
ncu -i gpurun_out/r2f_prof_$f.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) > 2:
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
    idx = [hdr.index(w) for w in want if w in hdr]
    for r in rows[2:]:
        print({hdr[i]: r[i] for i in idx})
PY
done
