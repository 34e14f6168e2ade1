"""BASELINE.json configs[0] ("A": dense random 200x200x200, rank 50, normal-equation CP-ALS, 100 sweeps) -- the
reference's own CPU-runnable case -- at FULL size: the north-star parity criterion, per-sweep fit within 1e-9 of the
oracle over 100 sweeps from identical initial factors, and the README stopping rule FitCheck(1e-3, 100, norm(T))."""
import numpy as np
import pytest

from oracle import cpals

pytestmark = pytest.mark.gpu


def test_config_a_fit_trajectory_100_sweeps_and_readme_rule(engine):
    import itcpd

    dims, R, nsweeps = (200, 200, 200), 50, 100
    rng = np.random.default_rng(0)
    T = np.asfortranarray(rng.standard_normal(dims))
    cp = cpals.random_CPD(T, R, np.random.default_rng(1))
    nT = float(np.linalg.norm(T))
    ref = cpals.FitCheck(0.0, nsweeps, nT)
    cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=ref)
    chk = itcpd.FitCheck(0.0, nsweeps, nT)
    itcpd.als_optimize(T, itcpd.CPD(cp.factors, cp.lam), check=chk)
    d = np.abs(np.array(chk.history) - np.array(ref.history))
    assert d.shape == (nsweeps,)
    assert d.max() <= 1e-9, (d.max(), int(d.argmax()))
    # README.md:96-129 rule: stops when |dfit| < 1e-3 twice in a row; same sweep count and final fit as the oracle
    r2 = cpals.FitCheck(1e-3, 100, nT)
    cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=r2)
    c2 = itcpd.FitCheck(1e-3, 100, nT)
    itcpd.als_optimize(T, itcpd.CPD(cp.factors, cp.lam), check=c2)
    assert c2.total_iter == r2.total_iter and abs(c2.final_fit - r2.final_fit) <= 1e-9


@pytest.mark.parametrize("gemm_i8,early_b", [(1, 0), (2, 0), (0, 1), (2, 1)])
def test_config_a_trajectory_with_the_experimental_contraction_paths(engine, gemm_i8, early_b):
    """the same north-star criterion (per-sweep fit within 1e-9 of the oracle over 100 sweeps) with the MTTKRP on the INT8
    tensor cores (6-digit base-256 split: 5e-14-level MTTKRP error) and / or pass B overlapped with mode 1's update"""
    import itcpd

    dims, R, nsweeps = (200, 200, 200), 50, 100
    rng = np.random.default_rng(0)
    T = np.asfortranarray(rng.standard_normal(dims))
    cp = cpals.random_CPD(T, R, np.random.default_rng(1))
    nT = float(np.linalg.norm(T))
    ref = cpals.FitCheck(0.0, nsweeps, nT)
    cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=ref)
    engine.set_option("gemm_i8", gemm_i8)
    engine.set_option("early_pass_b", early_b)
    try:
        engine.set_tensor(T)
        chk = itcpd.FitCheck(0.0, nsweeps, nT)
        itcpd.als_optimize(engine, itcpd.CPD(cp.factors, cp.lam), check=chk)
        d = np.abs(np.array(chk.history) - np.array(ref.history))
        assert d.shape == (nsweeps,) and d.max() <= 1e-9, (d.max(), int(d.argmax()))
    finally:
        engine.set_option("gemm_i8", 0)
        engine.set_option("early_pass_b", 0)
