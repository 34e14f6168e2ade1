"""one R = 64 solve through the right-looking pivoted Cholesky, for an ncu capture of that kernel"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itcpd
from oracle import cpals
rng = np.random.default_rng(3)
T = np.asfortranarray(rng.standard_normal((96, 80, 72)))
cp = cpals.random_CPD(T, 64, rng)
eng = itcpd.Engine(0)
eng.set_option("chol_alg", 2)
eng.set_tensor(T); eng.set_cpd(cp.factors, cp.lam); eng.compute_grams()
for rep in range(3):
    eng.gram_hadamard(0, fetch=False); eng.mttkrp(0, fetch=False)
    print(eng.solve(0, 1e-6))
eng.close()
