#!/bin/bash
# Round-2 closing check on TWO GPUs:  gpurun --gpus 2 --timeout 400 -- 'bash tools/r2_final_check_n2.sh'
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_tests.log; tail -3 gpurun_out/r2h_tests.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29683 \
      bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h_N2.json 2> gpurun_out/r2h_err.log
python tools/r2_summary.py gpurun_out/r2h_N2.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2h_N2.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "parity", d.get("parity", {}).get("max_abs_dfit"), "extra", [(r["config"], r.get("value"), r.get("error")) for r in d.get("extra", [])])
PY
tail -2 gpurun_out/r2h_err.log
