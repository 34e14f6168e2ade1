#!/bin/bash
# Round-1 last GPU call: validate the team Cholesky (ITCPD_CHOL=1) against the whole single-GPU suite, then time it.
# Every step has its own timeout; results land in gpurun_out/.
mkdir -p gpurun_out
export ITCPD_CHOL=1
timeout 75 python -m pytest tests/test_gpu_dense.py tests/test_gpu_golden.py tests/test_gpu_edge.py tests/test_gpu_sampled.py tests/test_gpu_config_a.py -x -q > gpurun_out/fc_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/fc_tests.log
tail -3 gpurun_out/fc_tests.log
B="timeout 40 python bench.py --no-cpu --no-e2e"
ITCPD_CHOL=1 $B --steps 30 > gpurun_out/fc_B_chol1.json 2>> gpurun_out/fc_err.log
ITCPD_CHOL=0 $B --config B8 --steps 50 > gpurun_out/fc_B8_chol0.json 2>> gpurun_out/fc_err.log
ITCPD_CHOL=1 $B --config B8 --steps 50 > gpurun_out/fc_B8_chol1.json 2>> gpurun_out/fc_err.log
ITCPD_CHOL=1 ITCPD_NO_GRAPH=1 $B --config B8 --steps 50 > gpurun_out/fc_B8_chol1_nograph.json 2>> gpurun_out/fc_err.log
ITCPD_CHOL=0 $B --config A --steps 50 > gpurun_out/fc_A_chol0.json 2>> gpurun_out/fc_err.log
ITCPD_CHOL=1 $B --config A --steps 50 > gpurun_out/fc_A_chol1.json 2>> gpurun_out/fc_err.log
for f in gpurun_out/fc_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"], 2), "sweeps/s", round(d["ms_per_step"], 4), "ms gemm", round(d["roofline"]["launch_ms"], 4))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
ITCPD_CHOL=1 ITCPD_NO_GRAPH=1 timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_B8.csv \
    python bench.py --config B8 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/fc_ncu.log 2>&1
echo "ncu rc=$?"
