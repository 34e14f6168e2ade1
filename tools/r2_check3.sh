#!/bin/bash
# Round-2 check after the fused mode tail (TWO GPUs, charged twice):  gpurun --gpus 2 --timeout 1500 -- 'bash tools/r2_check3.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_tests.log; tail -6 gpurun_out/r2d_tests.log
B="timeout 120 python bench.py --no-cpu --no-e2e --no-extras"
for cfg in B B8 A; do
  steps=20; [ $cfg != B ] && steps=100
  $B --config $cfg --steps $steps > gpurun_out/r2d_${cfg}_default.json 2>> gpurun_out/r2d_err.log
  ITCPD_FUSED_TAIL=0 $B --config $cfg --steps $steps > gpurun_out/r2d_${cfg}_tail0.json 2>> gpurun_out/r2d_err.log
done
ITCPD_BENCH_PHASES=1 $B --config B8 --steps 20 > gpurun_out/r2d_phases_B8.json 2>> gpurun_out/r2d_err.log
for t in 1 0; do
ITCPD_FUSED_TAIL=$t timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 \
      bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r2d_N2_tail$t.json 2>> gpurun_out/r2d_err.log
done
python tools/r2_summary.py gpurun_out/r2d_*.json | tee gpurun_out/r2d_summary.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2d_N2_tail1.json").read().strip().splitlines()[-1])
print("N2 parity", d.get("parity"))
PY
tail -3 gpurun_out/r2d_err.log
