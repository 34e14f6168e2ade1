#!/bin/bash
# Round-2 profiling pass for the INT8 contraction (ONE GPU, run AFTER tools/r2_single_gpu.sh is green):
#   gpurun --timeout 900 -- 'bash tools/r2_ncu.sh'
# 1. launch list of a short config-B run with the pre-packed INT8 contraction (shares per kernel; cold-cache, serialised)
# 2. one `--set full` capture of the dominant kernel (partial_gemm_i8p_kernel) with source correlation (-lineinfo is on)
# Expectations to hold the capture against (profiles/r1_i8_model.txt): 6.4 GB read + 0.54 GB written per launch = the
# algorithmic bytes; DRAM throughput > 80 % of peak; tensor pipe ~ 75 %; duration 1.0 - 1.3 ms.
mkdir -p gpurun_out
export ITCPD_GEMM_I8=${ITCPD_GEMM_I8:-2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_configB_i8.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_configB_i8.csv > gpurun_out/r2_launch_summary_configB_i8.txt 2>&1; cat gpurun_out/r2_launch_summary_configB_i8.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:partial_gemm_i8 -s 4 -c 2 -o gpurun_out/r2_prof_i8 $CMD > gpurun_out/r2_ncu_full.log 2>&1
ncu -i gpurun_out/r2_prof_i8.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) > 2:
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]
    idx = [hdr.index(w) for w in want if w in hdr]
    for r in rows[2:]:
        print({hdr[i]: r[i] for i in idx})
PY
