#!/bin/bash
# Round-2 final multi-GPU record (8 GPUs):  gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_multi_gpu3.sh'
mkdir -p gpurun_out
run() {  # run <tag> <nproc> <extra bench args...>
  tag=$1; n=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29671 \
      bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r2p_$tag.json 2>> gpurun_out/r2p_err.log
}
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2p_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_tests.log; tail -3 gpurun_out/r2p_tests.log
run N8 8
run N4 4
run N2 2
ITCPD_BENCH_PHASES=1 run N8_phases 8 --no-extras
ITCPD_BENCH_PHASES=1 run N4_phases 4 --no-extras
ITCPD_BENCH_PHASES=1 run N2_phases 2 --no-extras
python tools/r2_summary.py gpurun_out/r2p_*.json | tee gpurun_out/r2p_summary.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2p_N?.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, "value", round(d["value"], 2), "parity", (d.get("parity") or {}).get("max_abs_dfit"), [(r["config"], round(r.get("value", 0), 2), r.get("error")) for r in d.get("extra", [])])
PY
tail -3 gpurun_out/r2p_err.log
