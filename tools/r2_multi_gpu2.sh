#!/bin/bash
# Round-2 second multi-GPU run (8 GPUs):  gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_multi_gpu2.sh'
mkdir -p gpurun_out
run() {  # run <tag> <nproc> <extra bench args...>
  tag=$1; n=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29661 \
      bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r2n_$tag.json 2>> gpurun_out/r2n_err.log
}
run N8_tail1 8 --no-extras
ITCPD_FUSED_TAIL=0 run N8_tail0 8 --no-extras
ITCPD_FUSED_TAIL=0 ITCPD_CHOL=2 run N8_tail0_chol2 8 --no-extras
ITCPD_FUSED_TAIL=0 run N4_tail0 4
ITCPD_FUSED_TAIL=0 run N8_tail0_full 8
python tools/r2_summary.py gpurun_out/r2n_*.json | tee gpurun_out/r2n_summary.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2n_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, "value", round(d["value"], 2), "parity", (d.get("parity") or {}).get("max_abs_dfit"), [ (r["config"], round(r.get("value", 0), 2)) for r in d.get("extra", [])])
PY
tail -3 gpurun_out/r2n_err.log
