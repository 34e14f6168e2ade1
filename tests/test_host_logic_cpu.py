"""CPU tests of the host-side mirror (itensorcpd.jl_b200/host.py) with a fake C-ABI handle (tests/fake_engine.py):
the optimize loop, the convergence state machines fed by library scalars, the sampled-solver setups and the
rank-adaptive decompose must reproduce the oracle's control flow exactly (optimize.jl:6-35, fit_check.jl, decompose.jl)."""
import numpy as np
import pytest

import itcpd
from oracle import cpals, sampled

from fake_engine import FakeEngine


def problem(dims=(9, 8, 7), R=4, seed=0):
    rng = np.random.default_rng(seed)
    T = np.asfortranarray(rng.standard_normal(dims))
    cp = cpals.random_CPD(T, R, np.random.default_rng(seed + 1))
    return T, cp


def fake_with(T):
    e = FakeEngine()
    e.set_tensor(T)
    return e


@pytest.mark.parametrize("per_hook", [False, True])
def test_optimize_loop_matches_oracle(per_hook):
    T, cp = problem()
    nT = float(np.linalg.norm(T))
    ref = cpals.FitCheck(1e-5, 40, nT)
    o_ref = cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=ref)
    chk = itcpd.FitCheck(1e-5, 40, nT)
    als = itcpd.compute_als(fake_with(T), itcpd.CPD(cp.factors, cp.lam), check=chk)
    als.additional_items["per_hook"] = per_hook
    o = itcpd.optimize(itcpd.CPD(cp.factors, cp.lam), als)
    assert chk.history == ref.history and chk.total_iter == ref.total_iter and chk.final_fit == ref.final_fit
    for a, b in zip(o.factors, o_ref.factors):
        assert np.array_equal(a, b)
    assert np.array_equal(o.lam, o_ref.lam)


def test_default_algorithm_and_nocheck_default_maxiter():
    """als_optimizer.jl:45-47: alg defaults to KRPFreeNormal, check to NoCheck(maxiter or 100)."""
    T, cp = problem()
    e = fake_with(T)
    als = itcpd.compute_als(e, itcpd.CPD(cp.factors, cp.lam))
    assert isinstance(als.mttkrp_alg, itcpd.KRPFreeNormal) and isinstance(als.check, itcpd.NoCheck) and als.check.max_counter == 100
    als = itcpd.compute_als(e, itcpd.CPD(cp.factors, cp.lam), maxiter=7)
    assert als.check.max_counter == 7
    itcpd.optimize(itcpd.CPD(cp.factors, cp.lam), als)
    assert sum(n for name, n in e.calls if name == "sweep") == 7 and als.check.iter == 0


def test_partially_consumed_check_shortens_the_next_run():
    """optimize.jl:8,15,31: the while guard starts from check.iter."""
    T, cp = problem()
    chk = itcpd.NoCheck(10)
    chk.iter = 6
    e = fake_with(T)
    itcpd.als_optimize(e, itcpd.CPD(cp.factors, cp.lam), check=chk)
    assert sum(n for name, n in e.calls if name == "sweep") == 4
    ref = cpals.NoCheck(10)
    ref.iter = 6
    count = []
    als = cpals.compute_als(T, cp, alg=cpals.KRPNormal(), check=ref)
    cpals.optimize(cp, als, on_mode=lambda f, m, fa, l: count.append(f))
    assert len(count) == 4 * 3


def test_reused_fitcheck_prints_delta_equal_fit_first(capsys):
    """README.md:100-103 behaviour: after a finished run lastfit is 0, so the first delta of the next run equals the fit."""
    T, cp = problem()
    chk = itcpd.FitCheck(1e-3, 50, float(np.linalg.norm(T)))
    e = fake_with(T)
    o = itcpd.als_optimize(e, itcpd.CPD(cp.factors, cp.lam), check=chk)
    assert chk.lastfit == 0 and chk.iter == 0
    itcpd.als_optimize(e, o, check=chk, verbose=True)
    first = capsys.readouterr().out.strip().splitlines()[0].split("\t")
    assert abs(float(first[2]) - float(first[3])) < 1e-15


def test_decompose_api_errors_and_rank_adaptive():
    T, cp = problem((6, 7, 8), 3, seed=3)
    with pytest.raises(TypeError):
        itcpd.decompose(T, 5, solver=T)                      # test/cp_als.jl:15
    class S(itcpd.CPDOptimizer):
        pass
    with pytest.raises(RuntimeError, match="OptimizerError"):
        itcpd.decompose(T, 5, solver=S())                    # decompose.jl:27-29
    e = fake_with(T)
    out = itcpd.decompose(e, 1e-3, 56, start_rank=28, rank_step=28, rng=np.random.default_rng(3))
    ref = cpals.decompose_adaptive(T, 1e-3, 56, start_rank=28, rank_step=28, rng=np.random.default_rng(3), alg=cpals.KRPNormal())
    assert out.rank == ref.rank
    for a, b in zip(out.factors, ref.factors):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("Check", ["CPDiffCheck", "CPAngleCheck"])
def test_diff_and_angle_state_machines_match_oracle(Check):
    T, cp = problem((10, 9, 8), 3, seed=5)
    c1 = getattr(itcpd, Check)(1e-4, 50)
    itcpd.als_optimize(fake_with(T), itcpd.CPD(cp.factors, cp.lam), alg=itcpd.KRPNormal(), check=c1)
    c2 = getattr(cpals, Check)(1e-4, 50)
    cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=c2)
    assert c1.total_iter == c2.total_iter and c1.final_fit == c2.final_fit and c1.iter == c2.iter == 0


def test_sampled_solvers_control_flow():
    rng = np.random.default_rng(6)
    A = cpals.reconstruct(cpals.random_CPD((10, 11, 9), 3, rng))
    cp0 = cpals.random_CPD(A, 3, rng)
    start = itcpd.CPD(cp0.factors, cp0.lam)
    nA = np.linalg.norm(A)
    for alg, kw in [(itcpd.LevScoreSampled(120), dict(normal=True)), (itcpd.LevScoreSampled(120), {}),
                    (itcpd.BlockLevScoreSampled(120, 3), dict(normal=True)),
                    (itcpd.QRPivProjected(90), {}), (itcpd.SEQRCSPivProjected(1, 90, (1, 2, 3), (8, 8, 8)), dict(seed=1)),
                    (itcpd.KSEQRCSPivProjected(1, (90,), (1, 2, 3), 5), dict(seed=2))]:
        e = fake_with(A)
        o = itcpd.als_optimize(e, start, alg=alg, check=itcpd.CPDiffCheck(1e-6, 60), rng=np.random.default_rng(9), **kw)
        assert np.linalg.norm(A - cpals.reconstruct(cpals.CPD(o.factors, o.lam))) / nA < 0.2, type(alg).__name__
        if isinstance(alg, (itcpd.QRPivProjected, itcpd.SEQRCSPivProjected, itcpd.KSEQRCSPivProjected)):
            # ALS(ITensor(inds(target)), ...) drops the ALS object's reference to the tensor, not the caller's: a tensor that lives in an
            # engine the CALLER passed in stays resident, so that a second set-up on the same engine (the rank-adaptive loop,
            # decompose.jl:51-66) finds it again (it used to be dropped, and the second rank step failed)
            assert e.T is not None
            o2 = itcpd.als_optimize(e, start, alg=alg, check=itcpd.NoCheck(3), rng=np.random.default_rng(9), **kw)
            assert np.all(np.isfinite(o2.factors[0]))


def test_update_samples_bookkeeping():  # test/rand_cp_als.jl:28-36
    T, cp = problem((8, 9, 10), 6, seed=7)
    e = fake_with(T)
    als = itcpd.compute_als(e, itcpd.CPD(cp.factors, cp.lam), alg=itcpd.QRPivProjected(60), check=itcpd.FitCheck(1e-6, 5, 1.0), trunc_tol=4)
    als2 = itcpd.update_samples(T, als, 70, reshuffle=False)
    assert itcpd.stop(als2.mttkrp_alg) == 70 and itcpd.start(als2.mttkrp_alg) == 1 and type(als2.mttkrp_alg) is itcpd.QRPivProjected
    assert als.additional_items["effective_ranks"][0] < 8 and als2.additional_items["projects_tensors"][0].shape == (70, 2)
    itcpd.optimize(itcpd.CPD(cp.factors, cp.lam), als2)   # FitCheck with a sampled solver: runs max_counter sweeps
    assert als2.check.iter == 0


def test_fitcheck_disabled_warning_for_sampled(capsys):  # ProjectionAlgorithm.jl:30-51
    T, cp = problem((8, 9, 7), 3, seed=8)
    chk = itcpd.FitCheck(1e-3, 3, float(np.linalg.norm(T)))
    itcpd.als_optimize(fake_with(T), itcpd.CPD(cp.factors, cp.lam), alg=itcpd.LevScoreSampled(60), normal=True, check=chk)
    assert "FitCheck is not enabled" in capsys.readouterr().out and chk.iter == 0


def test_cpd_container_semantics():  # cpd.jl:7-46, test/basic_features.jl:95-106
    a = itcpd.random_CPD((4, 5, 6), 3, np.random.default_rng(1))
    assert len(a) == 3 and itcpd.cp_rank(a) == 3 and a.copy() == a and [f.shape for f in a] == [(4, 3), (5, 3), (6, 3)]
    b = a.copy()
    b.factors[0][0, 0] += 1
    assert not (a == b)
