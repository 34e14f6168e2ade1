"""GPU parity tests of the sampled / randomized path through the C-ABI against the oracle (oracle/sampled.py).
Gathers, index maps and sketches are exact (bit-for-bit or 1e-12); sampling is checked statistically, exactly as
the reference's own tests do (test/pivot_mapping.jl, test/SEQRCS_test.jl, test/rand_cp_als.jl)."""
import numpy as np
import pytest

from oracle import cpals, sampled

pytestmark = pytest.mark.gpu


def problem(dims, R, seed):
    rng = np.random.default_rng(seed)
    T = np.asfortranarray(rng.standard_normal(dims))
    cp = cpals.random_CPD(T, R, np.random.default_rng(seed + 1))
    return T, cp, rng


@pytest.mark.parametrize("dims", [(10, 15, 6), (5, 10, 15, 6), (13, 7, 9)])
def test_pivot_hadamard_exact(engine, dims):  # test/pivot_mapping.jl:56-86
    T, cp, rng = problem(dims, 7, 1)
    engine.set_tensor(T)
    engine.set_cpd(cp.factors, cp.lam)
    for mode in range(len(dims)):
        rd = [dims[m] for m in range(len(dims)) if m != mode]
        cols = rng.integers(1, int(np.prod(rd)) + 1, size=57)
        piv = sampled.column_to_multi_coords(cols, rd)
        K = engine.pivot_hadamard(mode, piv)
        Ko = sampled.pivot_hadamard([f for m, f in enumerate(cp.factors) if m != mode], piv)
        assert np.array_equal(K, Ko)
        full = cpals.khatri_rao([f for m, f in enumerate(cp.factors) if m != mode])
        assert np.max(np.abs(K - full[cols - 1])) < 1e-15


@pytest.mark.parametrize("dims", [(10, 15, 6), (5, 10, 15, 6), (13, 7, 9)])
def test_gather_fibers_exact(engine, dims):  # fused_flatten_sample, test/pivot_mapping.jl:88-103
    T, cp, rng = problem(dims, 3, 2)
    engine.set_tensor(T)
    for mode in range(len(dims)):
        rd = [dims[m] for m in range(len(dims)) if m != mode]
        cols = rng.integers(1, int(np.prod(rd)) + 1, size=41)
        piv = sampled.column_to_multi_coords(cols, rd)
        S = engine.gather_fibers(mode, piv)
        assert np.array_equal(S, cpals.unfold(T, mode)[:, cols - 1])
        assert np.array_equal(S, sampled.fused_flatten_sample(T, mode, piv))


def test_out_of_range_pivot_is_rejected(engine):
    import itcpd

    T, cp, rng = problem((6, 7, 8), 3, 3)
    engine.set_tensor(T)
    engine.set_cpd(cp.factors, cp.lam)
    bad = np.array([[1, 9]], dtype=np.int64)  # second remaining mode of mode 0 has extent 8
    with pytest.raises(itcpd.ItcpdError):
        engine.gather_fibers(0, bad)


@pytest.mark.parametrize("inj", [False, True])
def test_sketch_unfolding_matches_dense(engine, inj):  # test/pivot_mapping.jl:110-119, SEQRCS_test.jl:20-25
    import itcpd

    T, cp, rng = problem((12, 10, 14), 3, 4)
    engine.set_tensor(T)
    for mode in range(3):
        n = T.size // T.shape[mode]
        l, s = 25, 3
        rows0, vals, cs = itcpd.sparse_sign_matrix(l, n, s, injective=inj, seed=31 + mode)
        A_sk = engine.sketch_unfolding(mode, l, s, rows0, vals)
        Ao = sampled.sketched_matricization(T, mode, l, rows0 + 1, vals, s)
        assert np.linalg.norm(A_sk - Ao) < 1e-12
        import scipy.sparse as sp
        om = sp.csc_matrix((vals, (rows0, np.repeat(np.arange(n), s))), shape=(l, n))
        assert np.linalg.norm(A_sk - cpals.unfold(T, mode) @ om.toarray().T) < 1e-12


@pytest.mark.parametrize("rows,R", [(40, 6), (300, 50), (5, 9), (64, 64)])
def test_leverage_scores(engine, rows, R):  # probability.jl:3-10
    rng = np.random.default_rng(5)
    dims = (rows, 7, 5)
    T = np.asfortranarray(rng.standard_normal(dims))
    cp = cpals.random_CPD(T, R, rng)
    engine.set_tensor(T)
    engine.set_cpd(cp.factors, cp.lam)
    p = engine.leverage_scores(0)
    po = sampled.compute_leverage_score_probability(cp.factors[0])
    assert np.max(np.abs(p - po)) < 1e-10
    assert abs(p.sum() - 1.0) < 1e-10 or rows < R


def test_weighted_sampling_distribution(engine):  # probability.jl:12-33 (statistical)
    rng = np.random.default_rng(6)
    dims = (30, 20, 10)
    T = np.asfortranarray(rng.standard_normal(dims))
    cp = cpals.random_CPD(T, 4, rng)
    # make mode 1's leverage very non-uniform
    f = [x.copy() for x in cp.factors]
    f[1][3, :] *= 30
    f[1], _ = cpals.row_norm(f[1])
    engine.set_tensor(T)
    engine.set_cpd(f, cp.lam)
    nsamp = 40000
    piv = engine.sample_factor_matrices(0, nsamp, seed=9)
    assert piv.shape == (nsamp, 2) and piv.dtype == np.int64
    assert piv[:, 0].min() >= 1 and piv[:, 0].max() <= 20 and piv[:, 1].min() >= 1 and piv[:, 1].max() <= 10
    for col, mode in ((0, 1), (1, 2)):
        p = sampled.compute_leverage_score_probability(f[mode])
        freq = np.bincount(piv[:, col] - 1, minlength=len(p)) / nsamp
        assert np.max(np.abs(freq - p)) < 5 * np.sqrt(p.max() / nsamp) + 2e-3
    # a different seed gives different samples, the same seed the same samples
    assert np.array_equal(piv, engine.sample_factor_matrices(0, nsamp, seed=9))
    assert not np.array_equal(piv, engine.sample_factor_matrices(0, nsamp, seed=10))


def test_sampled_update_matches_oracle_normal_equations(engine):  # ProjectionAlgorithm.jl:57-68, normal=true
    T, cp, rng = problem((14, 12, 10), 5, 7)
    engine.set_tensor(T)
    engine.set_cpd(cp.factors, cp.lam)
    mode = 1
    piv = np.asfortranarray(np.stack([rng.integers(1, 15, size=200), rng.integers(1, 11, size=200)], axis=1).astype(np.int64))
    engine.sampled_update(mode, piv)
    K = sampled.pivot_hadamard([cp.factors[0], cp.factors[2]], piv)
    Ts = sampled.fused_flatten_sample(T, mode, piv)
    X = cpals.ldiv_solve(K.T @ K, np.asfortranarray((Ts @ K).T)).T
    Ao, lo = cpals.row_norm(X)
    assert np.linalg.norm(engine.get_factor(mode) - Ao) / np.linalg.norm(Ao) < 1e-10
    assert np.linalg.norm(engine.get_lambda() - lo) / np.linalg.norm(lo) < 1e-10
    assert np.linalg.norm(engine.get_gram(mode) - cpals.gram(Ao)) < 1e-10


def test_lev_score_sampled_als_statistical(engine):
    """test/rand_cp_als.jl:99-150 scaled: leverage-score sampled ALS on an exactly low-rank tensor comes within 10 %
    of the exact-ALS error; CPDiffCheck is the stopping rule (FitCheck is disabled for sampled solvers)."""
    import itcpd

    rng = np.random.default_rng(8)
    A = cpals.reconstruct(cpals.random_CPD((24, 26, 22), 4, rng))
    nA = np.linalg.norm(A)
    cp0 = cpals.random_CPD(A, 3, rng)
    exact = itcpd.als_optimize(A, itcpd.CPD(cp0.factors, cp0.lam), check=itcpd.CPDiffCheck(1e-5, 100), alg=itcpd.KRPNormal())
    e_exact = np.linalg.norm(A - itcpd.reconstruct(exact)) / nA
    ok = False
    for attempt in range(5):
        o = itcpd.als_optimize(A, itcpd.CPD(cp0.factors, cp0.lam), alg=itcpd.LevScoreSampled(400), normal=True,
                               check=itcpd.CPDiffCheck(1e-5, 100), seed=attempt)
        e = np.linalg.norm(A - itcpd.reconstruct(o)) / nA
        if abs(e_exact - e) / e_exact < 0.1:
            ok = True
            break
    assert ok, (e_exact, e)


def test_angle_check_and_fitcheck_warning_for_sampled(engine, capsys):
    import itcpd

    rng = np.random.default_rng(10)
    A = np.asfortranarray(rng.standard_normal((12, 13, 3)))
    cp0 = cpals.random_CPD(A, 5, rng)
    o = itcpd.als_optimize(A, itcpd.CPD(cp0.factors, cp0.lam), alg=itcpd.LevScoreSampled(100), normal=True,
                           check=itcpd.CPAngleCheck(1e-5, 30))
    assert np.linalg.norm(A - itcpd.reconstruct(o)) / np.linalg.norm(A) < 1.0
    chk = itcpd.FitCheck(1e-3, 4, float(np.linalg.norm(A)))
    itcpd.als_optimize(A, itcpd.CPD(cp0.factors, cp0.lam), alg=itcpd.LevScoreSampled(100), normal=True, check=chk)
    assert "FitCheck is not enabled" in capsys.readouterr().out  # ProjectionAlgorithm.jl:30-36
    assert chk.iter == 0
