#!/bin/bash
# Per-kernel launch times (ncu, cold-cache, serialised: shares, not absolutes) and live phase tables (CUDA events, no profiler)
# for the single-GPU configurations:  gpurun --timeout 900 -- 'bash tools/r2_launch_lists.sh'
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e"
for cfg in B8 A B; do
  ITCPD_BENCH_PHASES=1 timeout 120 $B --config $cfg --steps 20 > gpurun_out/r2_phases_$cfg.json 2>> gpurun_out/r2_ll_err.log
done
for chol in 1 2; do
  ITCPD_CHOL=$chol timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches_B8_chol$chol.csv $B --config B8 --steps 3 --warmup 3 > gpurun_out/r2_ll.log 2>&1
  python tools/launch_summary.py gpurun_out/r2_launches_B8_chol$chol.csv > gpurun_out/r2_launch_summary_B8_chol$chol.txt 2>&1
done
ITCPD_CHOL=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches_A.csv $B --config A --steps 3 --warmup 3 > gpurun_out/r2_ll.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_A.csv > gpurun_out/r2_launch_summary_A.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches_B.csv $B --config B --steps 3 --warmup 3 > gpurun_out/r2_ll.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_B.csv > gpurun_out/r2_launch_summary_B.txt 2>&1
python tools/r2_summary.py gpurun_out/r2_phases_*.json
tail -n 30 gpurun_out/r2_launch_summary_B8_chol1.txt gpurun_out/r2_launch_summary_B8_chol2.txt
