#!/bin/bash
# Round-2 opener: first hardware run of everything written after round 1's GPU budget was spent.
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/r2_experimental_check.sh'
# 1. the whole GPU suite with the opt-in tests (right-looking Cholesky chol_alg=2, sharded sampled path, peer_graph)
# 2. A/B timings: chol_alg 1 vs 2 on the per-rank slab and config A; 2-GPU sweeps with and without peer_graph
mkdir -p gpurun_out
# 0. single-instruction probe of the INT8 tensor-core building blocks (build it first: see the header of tools/i8_probe.cu)
[ -x tools/i8_probe ] && timeout 30 ./tools/i8_probe > gpurun_out/r2_i8_probe.txt 2>&1; tail -6 gpurun_out/r2_i8_probe.txt
export ITCPD_EXPERIMENTAL=1
# 0b. the INT8 contraction alone first (both variants, incl. split-K shapes), so its verdict survives whatever follows
timeout 180 python -m pytest tests/test_gpu_dense.py -m gpu -q -k gemm_i8 > gpurun_out/r2_i8_tests.log 2>&1
echo "i8 tests rc=$?" >> gpurun_out/r2_i8_tests.log
tail -8 gpurun_out/r2_i8_tests.log
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2_exp_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_exp_tests.log
tail -15 gpurun_out/r2_exp_tests.log
B="timeout 60 python bench.py --no-cpu --no-e2e"
for chol in 1 2; do
  ITCPD_CHOL=$chol $B --config B8 --steps 50 > gpurun_out/r2_B8_chol$chol.json 2>> gpurun_out/r2_err.log
  ITCPD_CHOL=$chol $B --config A --steps 50 > gpurun_out/r2_A_chol$chol.json 2>> gpurun_out/r2_err.log
done
# INT8 tensor-core contraction (draft): config B with and without it
ITCPD_GEMM_I8=1 $B --steps 20 > gpurun_out/r2_B_gemm_i8.json 2>> gpurun_out/r2_err.log
ITCPD_GEMM_I8=2 $B --steps 20 > gpurun_out/r2_B_gemm_i8_prepacked.json 2>> gpurun_out/r2_err.log
$B --steps 20 > gpurun_out/r2_B_dmma.json 2>> gpurun_out/r2_err.log
# pass B on its own stream under mode 1's update (early_pass_b), alone and on top of the pre-packed INT8 contraction
ITCPD_EARLY_B=1 $B --steps 20 > gpurun_out/r2_B_dmma_earlyb.json 2>> gpurun_out/r2_err.log
ITCPD_EARLY_B=1 ITCPD_GEMM_I8=2 $B --steps 20 > gpurun_out/r2_B_gemm_i8_prepacked_earlyb.json 2>> gpurun_out/r2_err.log
for pg in 0 1; do
  ITCPD_BENCH_PHASES=1 ITCPD_PEER_GRAPH=$pg timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 \
      bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r2_N2_peergraph$pg.json 2>> gpurun_out/r2_err.log
done
for f in gpurun_out/r2_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"], 2), "sweeps/s", round(d["ms_per_step"], 4), "ms")
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
