"""BASELINE.json config E: randomized CP-ALS (leverage-score sampling, SE-QRCS pivot sampling) on 1024^3 rank 64
vs exact ALS, on a planted rank-64 + noise tensor generated on the device.  Prints JSON lines."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itcpd
from bench import init_factors

small = len(sys.argv) > 1 and sys.argv[1] == "small"
dims, R = ((128, 128, 128), 16) if small else ((1024, 1024, 1024), 64)
P = float(np.prod(dims))
noise = 0.1 * np.sqrt(R) / np.sqrt(P)          # noise norm = 10 % of the signal norm
sweeps = 20
eng = itcpd.Engine(0)
out = []


def planted():
    t0 = time.perf_counter()
    eng.generate_lowrank_tensor(dims, R, seed=11, noise=noise)
    return time.perf_counter() - t0


def fit_of(cp):
    eng.set_cpd(cp.factors, cp.lam)
    return 1.0 - eng.residual_norm() / eng.tensor_norm()


gen_s = planted()
nT = eng.tensor_norm()
cp0 = itcpd.CPD(init_factors(dims, R, seed=1), np.ones(R))
print(json.dumps({"generate_s": gen_s, "norm": nT, "noise_sigma": noise}), flush=True)

# exact ALS
chk = itcpd.FitCheck(0.0, sweeps, nT)
t0 = time.perf_counter(); cp = itcpd.als_optimize(eng, cp0, check=chk); dt = time.perf_counter() - t0
out.append({"alg": "exact (KRPFreeNormal -> B200 tree)", "sweeps": sweeps, "s_per_sweep": dt / sweeps, "fit": chk.history[-1], "fit_sweep5": chk.history[4]})
print(json.dumps(out[-1]), flush=True)

# leverage-score sampled ALS
for ns in (10 * R, 64 * R):
    t0 = time.perf_counter()
    cp = itcpd.als_optimize(eng, cp0, alg=itcpd.LevScoreSampled(ns), normal=True, check=itcpd.NoCheck(sweeps), seed=5)
    dt = time.perf_counter() - t0
    out.append({"alg": f"LevScoreSampled({ns})", "sweeps": sweeps, "s_per_sweep": dt / sweeps, "fit": fit_of(cp),
                "algorithmic_bytes_per_sweep": 3 * 8.0 * (dims[0] * ns + 2 * ns * R + ns * R)})
    print(json.dumps(out[-1]), flush=True)

# SE-QRCS pivot-projected ALS (setup: sketch + 2 QRCPs per mode on the device)
for ns, ksk in ((64 * R, 2 * R),):
    t0 = time.perf_counter()
    als = itcpd.compute_als(eng, cp0, alg=itcpd.SEQRCSPivProjected(1, ns, (1, 2, 3), (ksk,) * 3), check=itcpd.NoCheck(sweeps), seed=9)
    setup = time.perf_counter() - t0
    t0 = time.perf_counter(); cp = itcpd.optimize(cp0, als); dt = time.perf_counter() - t0
    planted()  # the setup released the dense tensor; regenerate it (same seed) to evaluate the exact fit
    out.append({"alg": f"SEQRCSPivProjected(1,{ns}, rank_vect={ksk})", "setup_s": setup, "sweeps": sweeps, "s_per_sweep": dt / sweeps,
                "fit": fit_of(cp), "effective_ranks": als.additional_items["effective_ranks"]})
    print(json.dumps(out[-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"dims": dims, "rank": R, "noise_sigma": noise, "results": out}, open("gpurun_out/config_E.json", "w"), indent=1)
