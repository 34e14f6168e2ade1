"""sampled sweeps timed in a handle that has just run dense sweeps (the state bench.py's config-E record measures in)"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itcpd
from bench import init_factors
dims, R, sweeps = (1024, 1024, 1024), 64, 20
P = float(np.prod(dims))
eng = itcpd.Engine(0)
eng.generate_lowrank_tensor(dims, R, seed=11, noise=0.1 * np.sqrt(R) / np.sqrt(P))
nT = eng.tensor_norm()
cp0 = itcpd.CPD(init_factors(dims, R, seed=1), np.ones(R))
itcpd.als_optimize(eng, cp0, check=itcpd.FitCheck(0.0, 3, nT))      # dense sweeps first
for ns in (640, 4096):
    itcpd.als_optimize(eng, cp0, alg=itcpd.LevScoreSampled(ns), normal=True, check=itcpd.NoCheck(3), seed=4)
    als = itcpd.compute_als(eng, cp0, alg=itcpd.LevScoreSampled(ns), normal=True, check=itcpd.NoCheck(sweeps), seed=5)
    eng.synchronize(); t0 = time.perf_counter(); cp = itcpd.optimize(cp0, als); eng.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"alg": f"LevScoreSampled({ns}) after dense sweeps", "ms_per_sweep": 1e3 * dt / sweeps}), flush=True)
eng.close()
