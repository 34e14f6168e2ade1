"""where the SE-QRCS set-up of config E goes (ITCPD_TRACE_SETUP=1 prints the library's own marks; the host driver's calls are timed here)"""
import os, sys, time
import numpy as np
os.environ["ITCPD_TRACE_SETUP"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itcpd
from bench import init_factors
dims, R = (1024, 1024, 1024), 64
P = float(np.prod(dims))
eng = itcpd.Engine(0)
cp0 = itcpd.CPD(init_factors(dims, R, seed=1), np.ones(R))
log = []
def wrap(name):
    orig = getattr(eng, name)
    def f(*a, **k):
        eng.synchronize(); t0 = time.perf_counter(); r = orig(*a, **k); eng.synchronize()
        log.append((name, round(1e3 * (time.perf_counter() - t0), 1)))
        return r
    setattr(eng, name, f)
for name in ("seqrcs_modes", "set_projector", "drop_tensor", "set_cpd"):
    wrap(name)
for rep in range(int(os.environ.get("REPS", 3))):
    eng.generate_lowrank_tensor(dims, R, seed=11, noise=0.1 * np.sqrt(R) / np.sqrt(P))
    log.clear()
    eng.synchronize(); t0 = time.perf_counter()
    als = itcpd.compute_als(eng, cp0, alg=itcpd.SEQRCSPivProjected(1, 4096, (1, 2, 3), (128,) * 3), check=itcpd.NoCheck(5), seed=9)
    eng.synchronize(); print("setup", rep, round(time.perf_counter() - t0, 3), "s; host driver calls (ms):", log, flush=True)
eng.close()
