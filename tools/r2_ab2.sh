#!/bin/bash
# Round-2 A/B of the latency-chain changes (ONE GPU):  gpurun --timeout 1200 -- 'bash tools/r2_ab2.sh'
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r2b_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_tests.log; tail -5 gpurun_out/r2b_tests.log
B="timeout 120 python bench.py --no-cpu --no-e2e --no-extras"
for cfg in B8 A B D8; do
  steps=20; [ $cfg = A ] && steps=100; [ $cfg = B8 ] && steps=100
  $B --config $cfg --steps $steps > gpurun_out/r2b_${cfg}_default.json 2>> gpurun_out/r2b_err.log
  ITCPD_SOLVE=0 $B --config $cfg --steps $steps > gpurun_out/r2b_${cfg}_solve0.json 2>> gpurun_out/r2b_err.log
  ITCPD_CHOL=1 $B --config $cfg --steps $steps > gpurun_out/r2b_${cfg}_chol1.json 2>> gpurun_out/r2b_err.log
  ITCPD_CHOL=2 $B --config $cfg --steps $steps > gpurun_out/r2b_${cfg}_chol2.json 2>> gpurun_out/r2b_err.log
done
# early pass B: with and without the CUDA graph (does the overlap survive the replay?)
ITCPD_EARLY_B=1 $B --config B --steps 20 > gpurun_out/r2b_B_earlyb_graph.json 2>> gpurun_out/r2b_err.log
ITCPD_EARLY_B=1 ITCPD_NO_GRAPH=1 $B --config B --steps 20 > gpurun_out/r2b_B_earlyb_nograph.json 2>> gpurun_out/r2b_err.log
ITCPD_NO_GRAPH=1 $B --config B --steps 20 > gpurun_out/r2b_B_nograph.json 2>> gpurun_out/r2b_err.log
ITCPD_BENCH_PHASES=1 $B --config B8 --steps 20 > gpurun_out/r2b_phases_B8.json 2>> gpurun_out/r2b_err.log
python tools/r2_summary.py gpurun_out/r2b_*.json | tee gpurun_out/r2b_summary.txt
tail -3 gpurun_out/r2b_err.log
