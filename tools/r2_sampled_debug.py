"""per-call timing of the device-resident sampled sweep inside the host driver (debugging aid)"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itcpd
from bench import init_factors
dims, R = (1024, 1024, 1024), 64
P = float(np.prod(dims))
eng = itcpd.Engine(0)
eng.generate_lowrank_tensor(dims, R, seed=11, noise=0.1 * np.sqrt(R) / np.sqrt(P))
cp0 = itcpd.CPD(init_factors(dims, R, seed=1), np.ones(R))
orig = eng.sampled_sweep_async
log = []
def wrapped(*a, **k):
    eng.synchronize(); t0 = time.perf_counter(); l0 = eng.launch_count
    orig(*a, **k)
    eng.synchronize(); log.append((round(1e3 * (time.perf_counter() - t0), 2), eng.launch_count - l0, eng.last_solve_status(0)))
eng.sampled_sweep_async = wrapped
for ns in (640, 4096):
    for per_hook in (False, True):
        als = itcpd.compute_als(eng, cp0, alg=itcpd.LevScoreSampled(ns), normal=True, check=itcpd.NoCheck(20), seed=5)
        als.additional_items["per_hook"] = per_hook
        log.clear()
        eng.synchronize(); t0 = time.perf_counter(); cp = itcpd.optimize(cp0, als); eng.synchronize(); dt = time.perf_counter() - t0
        print(ns, "per_hook" if per_hook else "fused", round(1e3 * dt / 20, 3), "ms/sweep; calls:", log, "solve paths", als.solve_paths[:5], flush=True)
eng.close()
