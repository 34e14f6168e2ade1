"""Config E workload for profiling the sampled path on one B200: leverage-score sampled sweeps (640 / 4096 samples) and the SE-QRCS
pivot set-up on 1024^3 rank 64 (planted tensor + noise).  Prints JSON lines; run it plain for timings and under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:...` for the per-kernel table
(tools/r2_sampled_ncu.sh)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itcpd
from bench import init_factors

small = "small" in sys.argv
setup = "nosetup" not in sys.argv
dims, R = ((256, 256, 256), 32) if small else ((1024, 1024, 1024), 64)
P = float(np.prod(dims))
noise = 0.1 * np.sqrt(R) / np.sqrt(P)
sweeps = 20
eng = itcpd.Engine(0)
eng.generate_lowrank_tensor(dims, R, seed=11, noise=noise)
nT = eng.tensor_norm()
cp0 = itcpd.CPD(init_factors(dims, R, seed=1), np.ones(R))


def timed(fn):
    eng.synchronize(); t0 = time.perf_counter(); out = fn(); eng.synchronize(); return out, time.perf_counter() - t0


for ns in (10 * R, 64 * R):
    for per_hook in (False, True):
        warm = itcpd.compute_als(eng, cp0, alg=itcpd.LevScoreSampled(ns), normal=True, check=itcpd.NoCheck(3), seed=4)   # untimed: buffers + graph
        warm.additional_items["per_hook"] = per_hook
        itcpd.optimize(cp0, warm)
        als = itcpd.compute_als(eng, cp0, alg=itcpd.LevScoreSampled(ns), normal=True, check=itcpd.NoCheck(sweeps), seed=5)
        als.additional_items["per_hook"] = per_hook
        cp, dt = timed(lambda: itcpd.optimize(cp0, als))
        eng.set_cpd(cp.factors, cp.lam)
        print(json.dumps({"alg": f"LevScoreSampled({ns})", "driver": "per-mode host calls" if per_hook else "device-resident sweep (graph)",
                          "ms_per_sweep": 1e3 * dt / sweeps, "fit": 1.0 - eng.residual_norm() / nT}), flush=True)
if setup:
    ns, ksk = 64 * R, 2 * R
    als, ts = timed(lambda: itcpd.compute_als(eng, cp0, alg=itcpd.SEQRCSPivProjected(1, ns, (1, 2, 3), (ksk,) * 3), check=itcpd.NoCheck(sweeps), seed=9))
    cp, dt = timed(lambda: itcpd.optimize(cp0, als))
    print(json.dumps({"alg": f"SEQRCSPivProjected(1,{ns},rank_vect={ksk})", "setup_s": ts, "ms_per_sweep": 1e3 * dt / sweeps}), flush=True)
eng.close()
