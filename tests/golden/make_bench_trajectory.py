"""Full-size oracle trajectories on bench.py's own synthetic inputs (run in the build container, CPU only):

    python tests/golden/make_bench_trajectory.py B 32      # 1024^3 rank 64: ~20 s per sweep on 8 cores, 17 GB of RAM
    python tests/golden/make_bench_trajectory.py A 100     # 200^3 rank 50
    python tests/golden/make_bench_trajectory.py B512 20   # 512^3 rank 64 (tests/test_gpu_fullsize.py)

The tensor is regenerated on the CPU with oracle/synth.py (the numpy restatement of the device's counter-based
generator), the initial factors are bench.py's (numpy default_rng(1), column-normalised as src/cpd.jl:48-60), and the
sweeps are the oracle's (oracle/cpals.py: KRPNormal shape, dpstrf(tol=1e-6) + dgeqp3 fallback, FitCheck scalars).
Output: tests/golden/bench_trajectory_<config>.json with <T,That>, ||That||^2 and the fit of every sweep.  bench.py
compares the sweeps it ran against this file and prints max |dfit| in `config.parity`, at every N: that is the
driver-visible statement that the GPU path follows the CPU restatement of the reference at full size.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import cpals, synth  # noqa: E402

CONFIGS = {"A": ((200, 200, 200), 50), "B": ((1024, 1024, 1024), 64), "B512": ((512, 512, 512), 64), "S": ((256, 256, 256), 32)}


def main():
    name = sys.argv[1]
    nsweeps = int(sys.argv[2])
    dims, R = CONFIGS[name]
    t0 = time.time()
    T = synth.generate_tensor(dims, seed=0)
    nT = float(np.sqrt(np.sum(T * T)))
    print(f"tensor {dims} generated in {time.time() - t0:.1f} s, norm {nT!r}", flush=True)
    factors = synth.init_factors(dims, R, seed=1)
    cp = cpals.CPD([f.copy(order="F") for f in factors], np.ones(R))
    inner, norm2 = [], []

    class Rec(cpals.FitCheck):
        def update(self, i, n2, Rk, verbose=False):
            inner.append(float(i))
            norm2.append(float(n2))
            print(f"sweep {len(inner)}: {time.time() - t0:.1f} s", flush=True)
            return super().update(i, n2, Rk, verbose)

    chk = Rec(0.0, nsweeps, nT)
    cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=chk)
    out = {"config": name, "dims": list(dims), "rank": R, "tensor": "oracle/synth.py generate_tensor(dims, seed=0)",
           "factors": "oracle/synth.py init_factors(dims, R, seed=1), lambda = 1", "ref_norm": nT,
           "source": "oracle/cpals.py als_optimize(alg=KRPNormal(), check=FitCheck(0, nsweeps, norm(T)))",
           "inner": inner, "norm2": norm2, "fit": [float(x) for x in chk.history]}
    path = os.path.join(HERE, f"bench_trajectory_{name}.json")
    json.dump(out, open(path, "w"), indent=0)
    print("wrote", path)


if __name__ == "__main__":
    main()
