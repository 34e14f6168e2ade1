"""Generates the golden fixtures under tests/golden/ from the CPU oracle (run in the build container).
The reference ships no golden vectors and Julia is absent, so these pin the ORACLE (against drift) and give
the GPU box fixed input/output pairs that do not depend on numpy's RNG implementation."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import cpals, sampled  # noqa: E402


def dense():
    rng = np.random.default_rng(20261017)
    dims, R = (9, 8, 7), 5
    T = np.asfortranarray(rng.standard_normal(dims))
    cp = cpals.random_CPD(T, R, rng)
    chk = cpals.FitCheck(0.0, 30, float(np.linalg.norm(T)))
    cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=chk)
    out = {"dims": list(dims), "rank": R, "T": T.reshape(-1, order="F").tolist(),
           "factors": [f.reshape(-1, order="F").tolist() for f in cp.factors],
           "mttkrp": [cpals.mttkrp_krp_normal(T, cp.factors, n).reshape(-1, order="F").tolist() for n in range(3)],
           "fits": chk.history}
    json.dump(out, open(os.path.join(HERE, "dense_als.json"), "w"))


def sparse_sign():
    out = {}
    for name, inj in (("sparse_sign", False), ("sparsestack", True)):
        vals, rows, cs = sampled.sparse_sign_call(40, 25, 3, inj, which="ref", seed=12345)
        out[name] = {"l": 40, "n": 25, "s": 3, "seed": 12345, "rows": rows.tolist(),
                     "signs": [int(np.sign(v)) if np.isfinite(v) else 0 for v in vals], "colstarts": cs.tolist()}
    json.dump(out, open(os.path.join(HERE, "sparse_sign_ref.json"), "w"))


if __name__ == "__main__":
    dense()
    sparse_sign()
    print("golden fixtures written")
