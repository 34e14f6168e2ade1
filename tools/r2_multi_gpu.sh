#!/bin/bash
# Round-2 multi-GPU check (N GPUs of one box, charged N x):  gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_multi_gpu.sh 8'
N=${1:-8}
mkdir -p gpurun_out
run() {  # run <tag> <nproc> <extra bench args...>
  tag=$1; n=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29641 \
      bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r2m_$tag.json 2>> gpurun_out/r2m_err.log
}
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_tests.log; tail -3 gpurun_out/r2m_tests.log
run N${N}_default $N
ITCPD_PEER_GRAPH=0 run N${N}_peergraph0 $N --no-extras
ITCPD_BENCH_PHASES=1 run N${N}_phases $N --no-extras
ITCPD_PEER=0 run N${N}_nccl $N --no-extras
[ $N -gt 2 ] && run N2_default 2
python tools/r2_summary.py gpurun_out/r2m_*.json | tee gpurun_out/r2m_summary.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2m_*_default.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, "value", d["value"], "parity", d.get("parity"))
    for r in d.get("extra", []):
        print("   extra", json.dumps(r)[:700])
PY
tail -5 gpurun_out/r2m_err.log
