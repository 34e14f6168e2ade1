"""First-light / microbenchmark script run on the GPU box (not part of the product path)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itcpd
from oracle import cpals

out = {}
eng = itcpd.Engine(0)
out["device"] = eng.device_info()
out["dmma_tflops"] = eng.probe_dmma_peak()
out["dfma_tflops"] = eng.probe_dfma_peak()
print(json.dumps(out), flush=True)

def relerr(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))

# parity diagnostics on a small problem, each mode / both swizzle modes
rng = np.random.default_rng(0)
for dims, R in [((32, 32, 32), 8), ((20, 30, 40), 50), ((64, 64, 64), 64)]:
    T = np.asfortranarray(rng.standard_normal(dims)); cp = cpals.random_CPD(T, R, np.random.default_rng(1))
    for swz in (1, 0):
        eng.set_option("swizzle", swz)
        eng.set_tensor(T); eng.set_cpd(cp.factors, cp.lam)
        errs = []
        for n in range(len(dims)):
            M = eng.mttkrp(n); Mo = cpals.mttkrp_krp_normal(T, cp.factors, n)
            errs.append(relerr(M, Mo))
        print("parity", dims, R, "swizzle", swz, errs, flush=True)
    eng.set_option("swizzle", 1)

try:
    import torch
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    for _ in range(2): torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["cublas_dgemm_8192_tflops"] = 2 * 8192**3 / (best * 1e-3) / 1e12
    # the MTTKRP-shaped DGEMM: (2^20 x 1024) x (1024 x 64)
    a = torch.randn(1024, 1 << 20, dtype=torch.float64, device="cuda").t(); b = torch.randn(1024, 64, dtype=torch.float64, device="cuda")
    for _ in range(2): torch.matmul(a, b)
    torch.cuda.synchronize(); best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["cublas_dgemm_mttkrp_shape_tflops"] = 2 * (1 << 20) * 1024 * 64 / (best * 1e-3) / 1e12
    out["cublas_dgemm_mttkrp_shape_ms"] = best
    del a, b
    torch.cuda.empty_cache()
except Exception as ex:  # pragma: no cover
    out["torch_error"] = repr(ex)
print(json.dumps(out), flush=True)

# config B timing
for dims, R in [((1024, 1024, 1024), 64)]:
    t0 = time.time(); eng.generate_tensor(dims, seed=0); eng.set_rank(R); eng.random_cpd(1); eng.compute_grams(); eng.synchronize()
    print("generate s", time.time() - t0, flush=True)
    eng.set_option("time_gemm", 1)
    for warps in (8, 4):
        eng.set_option("tile_warps", warps)
        eng.sweep(2)
        eng.gemm_timing(True)
        t0 = time.time(); inner, n2 = eng.sweep(5); dt = (time.time() - t0) / 5
        ms, n = eng.gemm_timing(True)
        P = float(np.prod(dims))
        out[f"cfgB_w{warps}"] = {"sweep_ms": dt * 1e3, "gemm_ms": ms, "gemm_launches": n, "gemm_tflops": 2 * R * P / (ms * 1e-3) / 1e12,
                                 "gemm_GBs": 8 * P / (ms * 1e-3) / 1e9}
        print(json.dumps(out[f"cfgB_w{warps}"]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/first_light.json", "w"), indent=1)
