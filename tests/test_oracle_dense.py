"""CPU tests of the dense-path oracle: the reference's own property tests (test/cp_als.jl,
test/basic_features.jl) re-stated against oracle/cpals.py, plus the committed golden fixtures."""
import json
import os

import numpy as np
import pytest

from oracle import cpals

HERE = os.path.dirname(os.path.abspath(__file__))


def test_row_norm_unit_columns():  # test/basic_features.jl:1-16
    rng = np.random.default_rng(0)
    X = rng.standard_normal((20, 30))
    A, lam = cpals.row_norm(X)
    assert np.allclose(np.sum(A * A, axis=0), 1.0, rtol=0, atol=1e-14)
    assert np.allclose(A * lam[None, :], X)


def test_khatri_rao_against_triple_loop():  # test/basic_features.jl:18-58 (pairwise had_contract)
    rng = np.random.default_rng(1)
    A, B = rng.standard_normal((20, 40)), rng.standard_normal((30, 40))
    K = cpals.khatri_rao([A, B])
    ref = np.empty((20, 30, 40))
    for i in range(20):
        for j in range(30):
            ref[i, j, :] = A[i, :] * B[j, :]
    assert np.linalg.norm(K - ref.reshape(600, 40, order="F")) / np.linalg.norm(ref) < 1e-7


def test_reconstruct_against_loops():  # test/basic_features.jl:60-92
    rng = np.random.default_rng(2)
    cp = cpals.random_CPD((4, 5, 6), 3, rng)
    rec = cpals.reconstruct(cp)
    ref = np.zeros((4, 5, 6))
    for r in range(3):
        ref += cp.lam[r] * np.einsum("i,j,k->ijk", cp.factors[0][:, r], cp.factors[1][:, r], cp.factors[2][:, r])
    assert np.allclose(rec, ref, rtol=5 * np.finfo(float).eps * 10, atol=1e-15)


def test_two_mttkrp_formulations_agree():
    rng = np.random.default_rng(3)
    T = np.asfortranarray(rng.standard_normal((9, 8, 7, 6)))
    cp = cpals.random_CPD(T, 5, rng)
    for n in range(4):
        a = cpals.mttkrp_krp_normal(T, cp.factors, n)
        b = cpals.mttkrp_krp_free(T, cp.factors, n)
        assert np.linalg.norm(a - b) / np.linalg.norm(a) < 1e-13


def test_ldiv_solve_cholesky_and_fallback():
    rng = np.random.default_rng(4)
    A = rng.standard_normal((30, 8))
    G = A.T @ A
    B = rng.standard_normal((8, 5))
    info = {}
    X = cpals.ldiv_solve(G, B, info)
    assert info["path"] == "cholesky"
    assert np.allclose(G @ X, B)
    A[:, 5] = A[:, 1]
    G = A.T @ A
    X = cpals.ldiv_solve(G, B, info)
    assert info["path"] == "qrcp" and info["rank"] == 7
    Xp = np.linalg.pinv(G) @ B  # min-norm least squares
    assert np.allclose(X, Xp, atol=1e-8)


def test_fitcheck_state_machine():  # fit_check.jl:24-66
    def step(chk, inner, fact_square):  # check_converge = iter += 1 (fit_check.jl:25) + the scalar state machine
        chk.iter += 1
        return chk.update(inner, fact_square, 3)

    chk = cpals.FitCheck(1e-3, 5, 10.0)
    assert chk.lastfit == 1.0
    # residual^2 = 100 + m - 2 i ; fits 0.5, 0.5004, 0.5006 -> converged on the third call
    for fit, expect in [(0.5, False), (0.5004, False), (0.5006, True)]:
        resid = (1 - fit) * 10.0
        done = step(chk, (100 + 1.0 - resid ** 2) / 2, 1.0)
        assert done is expect
    assert chk.iter == 0 and chk.lastfit == 0 and abs(chk.final_fit - 0.5006) < 1e-12 and chk.total_iter == 3
    # max_counter reached: state reset but returns False (fit_check.jl:57-65)
    chk = cpals.FitCheck(1e-9, 2, 10.0)
    assert step(chk, 40.0, 1.0) is False
    assert step(chk, 45.0, 1.0) is False
    assert chk.iter == 0 and chk.total_iter == 2 and chk.lastfit == 0
    with pytest.raises(RuntimeError):
        step(chk, float("nan"), 1.0)


def test_nocheck_returns_true_at_max():  # no_check.jl:9-20
    chk = cpals.NoCheck(3)
    assert [chk.check_converge(None, np.ones(2), None) for _ in range(3)] == [False, False, True]
    assert chk.iter == 0


def test_overcomplete_als_reconstructs():  # test/cp_als.jl:9-43 scaled (10x12x14, R = 140)
    rng = np.random.default_rng(5)
    T = np.asfortranarray(rng.standard_normal((10, 12, 14)))
    nT = np.linalg.norm(T)
    cp = cpals.random_CPD(T, 140)
    opt = cpals.als_optimize(T, cp, alg=cpals.KRPNormal())
    assert np.linalg.norm(cpals.reconstruct(opt) - T) / nT < 5e-7
    chk = cpals.FitCheck(1e-6, 100, nT)
    opt = cpals.als_optimize(T, cp, alg=cpals.KRPFreeNormal(), check=chk, maxiter=None)
    assert np.linalg.norm(cpals.reconstruct(opt) - T) / nT < 1e-5
    opt = cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=cpals.NoCheck(10))
    assert np.linalg.norm(cpals.reconstruct(opt) - T) / nT < 1e-1


def test_final_fit_equals_explicit_fit():  # test/itensor_network_cpd.jl:104-108 (rtol 1e-3)
    rng = np.random.default_rng(6)
    T = np.asfortranarray(rng.standard_normal((12, 11, 10)))
    nT = np.linalg.norm(T)
    chk = cpals.FitCheck(1e-4, 60, nT)
    opt = cpals.als_optimize(T, cpals.random_CPD(T, 6), alg=cpals.KRPNormal(), check=chk)
    explicit = 1 - np.linalg.norm(T - cpals.reconstruct(opt)) / nT
    assert abs(chk.final_fit - explicit) < 1e-3 * abs(explicit) + 1e-12


def test_rank_adaptive():  # test/cp_als.jl:106-115 scaled
    rng = np.random.default_rng(7)
    T = np.asfortranarray(rng.standard_normal((6, 7, 8)))
    cp = cpals.decompose_adaptive(T, 1e-3, 56, start_rank=28, rank_step=28, alg=cpals.KRPNormal())
    assert np.linalg.norm(cpals.reconstruct(cp) - T) / np.linalg.norm(T) < 1e-3


def test_golden_dense_fixture():
    """tests/golden/dense_als.json was generated by tests/golden/make_golden.py from this oracle; it pins the
    oracle against silent drift (and is what the GPU path is compared with on the GPU box)."""
    g = json.load(open(os.path.join(HERE, "golden", "dense_als.json")))
    T = np.array(g["T"]).reshape(g["dims"], order="F")
    factors = [np.array(f).reshape((d, g["rank"]), order="F") for f, d in zip(g["factors"], g["dims"])]
    cp = cpals.CPD(factors, np.ones(g["rank"]))
    for n in range(len(g["dims"])):
        M = cpals.mttkrp_krp_normal(T, factors, n)
        assert np.allclose(M.reshape(-1, order="F"), np.array(g["mttkrp"][n]), rtol=1e-13, atol=1e-13)
    chk = cpals.FitCheck(0.0, len(g["fits"]), float(np.linalg.norm(T)))
    cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=chk)
    assert np.max(np.abs(np.array(chk.history) - np.array(g["fits"]))) < 1e-11
