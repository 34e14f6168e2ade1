#!/bin/bash
# Round-2 opener, part 1 (ONE GPU, ~12 min):  gpurun --timeout 1200 -- 'bash tools/r2_single_gpu.sh'
# First hardware run of what was written after round 1's GPU budget was spent.  Ordered so that the most valuable verdicts
# survive a time-out; everything lands in gpurun_out/r2_*.
mkdir -p gpurun_out
# 0. single-instruction probe of the INT8 tensor-core building blocks; diagnoses itself on a mismatch (tools/i8_probe.cu)
if [ ! -x tools/i8_probe ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -DITCPD_I8_PROBE -o tools/i8_probe tools/i8_probe.cu > gpurun_out/r2_i8_probe_build.log 2>&1
fi
timeout 60 ./tools/i8_probe > gpurun_out/r2_i8_probe.txt 2>&1; tail -25 gpurun_out/r2_i8_probe.txt
export ITCPD_EXPERIMENTAL=1
# 1. the INT8 contraction alone (both variants, ragged, 3 rank blocks, split-K), then the other single-GPU opt-in tests
timeout 200 python -m pytest tests/test_gpu_dense.py -m gpu -q -x -k "gemm_i8" > gpurun_out/r2_i8_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_i8_tests.log; tail -6 gpurun_out/r2_i8_tests.log
timeout 200 python -m pytest tests/test_gpu_dense.py tests/test_gpu_config_a.py -m gpu -q -k "early_pass_b or right_looking or experimental_contraction or single_sweep_calls" > gpurun_out/r2_exp_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_exp_tests.log; tail -8 gpurun_out/r2_exp_tests.log
# 2. A/B bench lines (config B unless stated), 20 timed sweeps each
B="timeout 90 python bench.py --no-cpu --no-e2e --steps 20"
$B > gpurun_out/r2_B_dmma.json 2>> gpurun_out/r2_err.log
ITCPD_EARLY_B=1 $B > gpurun_out/r2_B_dmma_earlyb.json 2>> gpurun_out/r2_err.log
ITCPD_GEMM_I8=2 $B > gpurun_out/r2_B_i8_prepacked.json 2>> gpurun_out/r2_err.log
ITCPD_GEMM_I8=2 ITCPD_EARLY_B=1 $B > gpurun_out/r2_B_i8_prepacked_earlyb.json 2>> gpurun_out/r2_err.log
ITCPD_GEMM_I8=1 $B > gpurun_out/r2_B_i8_on_the_fly.json 2>> gpurun_out/r2_err.log
for chol in 1 2 3; do   # B8 / A at R = 64 and 50; the rank-128 kernel shows in config D slabs (bench --config D8 if present)
  ITCPD_CHOL=$chol $B --config B8 --steps 50 > gpurun_out/r2_B8_chol$chol.json 2>> gpurun_out/r2_err.log
  ITCPD_CHOL=$chol $B --config A --steps 50 > gpurun_out/r2_A_chol$chol.json 2>> gpurun_out/r2_err.log
done
ITCPD_GEMM_I8=2 $B --config B8 --steps 50 > gpurun_out/r2_B8_i8_prepacked.json 2>> gpurun_out/r2_err.log
# config D's per-rank slab (R = 128: the two-thread right-looking Cholesky, two rank blocks in one INT8 launch, one SM left to the factorisation)
for chol in 1 2 3; do ITCPD_CHOL=$chol $B --config D8 --steps 20 > gpurun_out/r2_D8_chol$chol.json 2>> gpurun_out/r2_err.log; done
ITCPD_CHOL=2 ITCPD_GEMM_I8=2 $B --config D8 --steps 20 > gpurun_out/r2_D8_i8_chol2.json 2>> gpurun_out/r2_err.log
ITCPD_CHOL=2 ITCPD_GEMM_I8=2 ITCPD_I8_SPARE_SMS=1 $B --config D8 --steps 20 > gpurun_out/r2_D8_i8_chol2_spare1.json 2>> gpurun_out/r2_err.log
python tools/r2_summary.py gpurun_out/r2_*.json | tee gpurun_out/r2_summary.txt
# per-iteration loop of the reference-facing API (decompose + FitCheck, one itcpd_sweep(1) per iteration) on config A: graph_single off / on
for gs in 0 1; do
  ITCPD_GRAPH_SINGLE=$gs timeout 120 python - > gpurun_out/r2_A_decompose_graph_single$gs.txt 2>&1 <<'PY'
import time, numpy as np, itcpd
rng = np.random.default_rng(0)
T = np.asfortranarray(rng.standard_normal((200, 200, 200)))
nT = float(np.linalg.norm(T))
cp0 = itcpd.random_CPD(T, 50, rng=np.random.default_rng(1))
for rep in range(3):
    chk = itcpd.FitCheck(0.0, 100, nT)
    t0 = time.perf_counter()
    itcpd.als_optimize(T, cp0, check=chk)
    dt = time.perf_counter() - t0
    print(f"decompose-style loop, config A, 100 sweeps: {dt * 1e3:.1f} ms  ({100 / dt:.0f} sweeps/s, upload included)  final fit {chk.final_fit:.6f}")
PY
  tail -1 gpurun_out/r2_A_decompose_graph_single$gs.txt
done
# 3. the rest of the single-GPU suite with the opt-in tests on (regression of the default path included)
timeout 400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/r2_all_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_all_tests.log; tail -5 gpurun_out/r2_all_tests.log
