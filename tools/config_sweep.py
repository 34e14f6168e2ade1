"""Timing of every BASELINE.json single-GPU configuration (device resident), printed as JSON lines."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itcpd
from bench import CONFIGS, init_factors

eng = itcpd.Engine(0)
peak = eng.probe_dmma_peak()
for name in sys.argv[1:] or ["A", "B", "C", "D"]:
    dims, R = CONFIGS[name]["dims"], CONFIGS[name]["rank"]
    P = float(np.prod(dims))
    eng.generate_tensor(dims, seed=0)
    eng.set_cpd(init_factors(dims, R), np.ones(R))
    eng.compute_grams()
    flush = P * 8 < (256 << 20)
    eng.sweep(3)
    eng.set_option("time_gemm", 1); eng.gemm_timing(True)
    K = 20 if P < 2**32 else 5
    eng.event_record(0)
    if flush:
        for _ in range(K):
            eng.flush_l2(); eng.sweep_async(1)
    else:
        eng.sweep_async(K)
    eng.event_record(1)
    eng.synchronize()
    ms = eng.event_elapsed_ms(0, 1) / K
    gms, gn = eng.gemm_timing(True)
    eng.set_option("time_gemm", 0)
    print(json.dumps({"config": name, "dims": dims, "rank": R, "ms_per_sweep": ms, "sweeps_per_s": 1e3 / ms, "gemm_ms": gms,
                      "gemm_tflops": 2 * R * P / (gms * 1e-3) / 1e12, "gemm_frac_of_dmma_peak": 2 * R * P / (gms * 1e-3) / 1e12 / peak,
                      "gemm_GBs": 8 * P / (gms * 1e-3) / 1e9, "sweep_roofline_frac": (4 * R * P / (peak * 1e12)) / (ms * 1e-3),
                      "dmma_peak": peak}), flush=True)
