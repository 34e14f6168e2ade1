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
        # the sparse-matrix variant (pivot_mapping.jl:90-104; SEQRCS(...; use_omega = true)): same kernel, same summation order
        A_om = engine.sketch_unfolding_omega(mode, om)
        assert np.array_equal(A_om, A_sk)
        assert np.linalg.norm(A_om - sampled.sketched_matricization_omega(T, mode, om)) < 1e-12


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


@pytest.mark.parametrize("cond", [1e3, 1e6])
def test_leverage_scores_ill_conditioned(engine, cond):
    """Late ALS sweeps produce nearly collinear factor columns.  The reference takes a Householder QR (probability.jl:6); the device
    route is Cholesky-QR with a second-order Neumann correction (solve.cu: leverage_impl), accurate to ~ cond * eps where the plain
    Gram + Cholesky route of round 1 lost cond^2 * eps (1e-4 at cond 1e6)."""
    rng = np.random.default_rng(17)
    rows, R = 300, 20
    U, _ = np.linalg.qr(rng.standard_normal((rows, R)))
    V, _ = np.linalg.qr(rng.standard_normal((R, R)))
    A = np.asfortranarray((U * np.logspace(0, -np.log10(cond), R)[None, :]) @ V.T)
    assert abs(np.linalg.cond(A) / cond - 1) < 1e-6
    engine.set_tensor(np.zeros((rows, 3, 2), order="F"))
    engine.set_cpd([A, np.ones((3, R), order="F"), np.ones((2, R), order="F")], np.ones(R))
    p = engine.leverage_scores(0)
    q, _ = np.linalg.qr(A)
    po = np.sum(q * q, axis=1) / R
    assert np.max(np.abs(p - po)) < 1e-9 * np.max(po), (np.max(np.abs(p - po)) / np.max(po))
    assert abs(p.sum() - 1.0) < 1e-9


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


# ---------------------------------------------------------------------------------------------------------
# pivot-projected solvers: device QRCP, SE-QRCS, cached projectors
# ---------------------------------------------------------------------------------------------------------
def _rank_k_error(A, p, k):
    """|| A[:,p] - Q_k R_k || / ||A|| for the QR of the pivoted matrix truncated to k columns."""
    Ap = A[:, p - 1]
    Q, R = np.linalg.qr(Ap[:, :k])
    return np.linalg.norm(Ap - Q @ (Q.T @ Ap), 2) / np.linalg.norm(A, 2)


@pytest.mark.parametrize("m,n", [(12, 40), (50, 3000), (40, 25), (64, 64), (131, 500), (259, 400), (515, 700), (1029, 1200), (2053, 300), (2200, 2300)])
def test_device_qrcp_matches_lapack(engine, m, n):
    """qr(A, ColumnNorm()): same pivot order and |diag(R)| as LAPACK dgeqp3 on generic matrices.  The row counts walk through every
    register-resident instantiation of the apply kernel (2 ... 64 values per lane, odd sizes: unaligned column starts) and past it
    (more than 2048 rows: the two-pass kernel)."""
    rng = np.random.default_rng(20)
    A = np.asfortranarray(rng.standard_normal((m, n)) * np.exp(rng.standard_normal(n))[None, :])
    piv, rd = engine.qrcp_matrix(A)
    _, R, p = sampled.qrcp(A)
    k = min(m, n)
    assert sorted(piv.tolist()) == list(range(1, n + 1))
    assert np.array_equal(piv[:k], p[:k])
    assert np.max(np.abs(np.abs(rd) - np.abs(np.diag(R)))) < 1e-10 * np.abs(R[0, 0])


def test_device_qrcp_unfolding_all_modes(engine):
    rng = np.random.default_rng(21)
    T = np.asfortranarray(rng.standard_normal((9, 10, 11)) * np.exp(rng.standard_normal((1, 10, 1))))
    engine.set_tensor(T)
    for mode in range(3):
        piv, rd = engine.qrcp_unfolding(mode)
        _, R, p = sampled.qrcp(cpals.unfold(T, mode))
        k = T.shape[mode]
        assert np.array_equal(piv[:k], p[:k]) and sorted(piv.tolist()) == list(range(1, T.size // k + 1))
        assert np.max(np.abs(np.abs(rd) - np.abs(np.diag(R)))) < 1e-10 * np.abs(R[0, 0])


@pytest.mark.parametrize("inj", [False, True])
def test_device_seqrcs_matches_oracle_and_qrcp_quality(engine, inj):
    """test/SEQRCS_test.jl:32-47 scaled: with the same libc rand() stream the device SE-QRCS must select the same
    candidate set as the oracle, and its rank-k approximation error must be within 1e-2 of plain QRCP."""
    rng = np.random.default_rng(22)
    A = np.asfortranarray(rng.standard_normal((50, 3000)))
    k, l, s, t = 40, 750, 1, 40
    engine.set_tensor(A)
    piv, rd, ncand = engine.seqrcs(0, l, s, t, injective=inj, seed=17)
    info = {}
    Q, R, p = sampled.seqrcs_tensor(A, 0, l, s, t, use_omega=False, injective=inj, which="ref", seed=17, info=info)
    assert ncand == info["subset"]
    assert sorted(piv.tolist()) == list(range(1, 3001))
    assert set(piv[:ncand].tolist()) == set(p[:ncand].tolist())      # same candidate columns
    assert np.array_equal(piv[ncand:], p[ncand:])                     # same remainder order
    assert np.array_equal(piv[:k], p[:k])                             # same leading pivots
    _, _, p_act = sampled.qrcp(A)
    assert abs(_rank_k_error(A, piv, k) - _rank_k_error(A, p_act, k)) <= 1e-2
    # the sparse-matrix variant (SEQRCS.jl:89-134, use_omega = true): same sketch, candidates listed in increasing column order
    piv_o, _, ncand_o = engine.seqrcs(0, l, s, t, injective=inj, seed=17, use_omega=True)
    _, _, p_o = sampled.seqrcs_tensor(A, 0, l, s, t, use_omega=True, injective=inj, which="ref", seed=17)
    assert ncand_o == ncand and set(piv_o[:ncand].tolist()) == set(piv[:ncand].tolist())
    assert np.array_equal(piv_o[:k], p_o[:k]) and np.array_equal(piv_o[ncand:], p_o[ncand:])
    engine.set_option("seqrcs_use_omega", 0)


def test_seqrcs_modes_equals_per_mode_calls_and_threaded_sort(engine, monkeypatch):
    """itcpd_seqrcs_modes (generator + row sort of mode n+1 on a helper thread while the device factorises mode n) against one
    itcpd_seqrcs per mode under the same seeds: identical pivot lists, with the stable counting sort forced onto 5 threads."""
    rng = np.random.default_rng(41)
    T = np.asfortranarray(rng.standard_normal((24, 30, 20)))
    engine.set_tensor(T)
    modes, ls, ss, ts = [0, 1, 2], [90, 110, 70], [2, 3, 1], [12, 10, 9]
    single = [engine.seqrcs(m, l, s, t, seed=50 + m) for m, l, s, t in zip(modes, ls, ss, ts)]
    monkeypatch.setenv("ITCPD_SKETCH_THREADS", "5")
    both = engine.seqrcs_modes(modes, ls, ss, ts, seeds=[50, 51, 52])
    for (p1, r1, c1), (p2, r2, c2) in zip(single, both):
        assert c1 == c2 and np.array_equal(p1, p2) and np.array_equal(r1, r2)
    # one stream across the call (seed only the first mode) == per-mode calls that let the stream run on
    import ctypes
    ctypes.CDLL(None).srand(77)
    cont = [engine.seqrcs(m, l, s, t) for m, l, s, t in zip(modes, ls, ss, ts)]
    both = engine.seqrcs_modes(modes, ls, ss, ts, seeds=[77, None, None])
    for (p1, r1, c1), (p2, r2, c2) in zip(cont, both):
        assert c1 == c2 and np.array_equal(p1, p2) and np.array_equal(r1, r2)
    # the explicit-unfolding path of the set-up (tiled transpose, then contiguous columns) against the in-place strided gather
    engine.set_option("sketch_unfold", 0)
    inplace = engine.seqrcs_modes(modes, ls, ss, ts, seeds=[50, 51, 52])
    engine.set_option("sketch_unfold", 1)
    for (p1, r1, c1), (p2, r2, c2) in zip(single, inplace):
        assert c1 == c2 and np.array_equal(p1, p2) and np.array_equal(r1, r2)
    # the sketch itself under the threaded sort, against the oracle
    import itcpd
    rows0, vals, cs = itcpd.sparse_sign_matrix(110, 24 * 20, 3, seed=3)
    A_sk = engine.sketch_unfolding(1, 110, 3, rows0, vals)
    Ao = sampled.sketched_matricization(T, 1, 110, rows0 + 1, vals, 3)
    assert np.linalg.norm(A_sk - Ao) <= 1e-12 * np.linalg.norm(Ao)


@pytest.mark.parametrize("dims", [(23, 37, 41), (64, 32, 96), (7, 5, 3, 9), (33, 65)])
def test_sketch_from_explicit_unfolding_is_bitwise_the_inplace_sketch(engine, dims):
    """option sketch_unfold: the tiled transpose (partial tiles, a padded leading dimension, order 2 and 4) followed by the sketch on
    contiguous columns must give exactly the numbers of the strided in-place gather, for every mode."""
    import itcpd
    rng = np.random.default_rng(sum(dims))
    T = np.asfortranarray(rng.standard_normal(dims))
    engine.set_tensor(T)
    try:
        for mode in range(len(dims)):
            n = T.size // dims[mode]
            l, s = 3 * dims[mode] + 1, 2
            rows0, vals, cs = itcpd.sparse_sign_matrix(l, n, s, seed=9 + mode)
            engine.set_option("sketch_unfold", 0)
            a = engine.sketch_unfolding(mode, l, s, rows0, vals)
            engine.set_option("sketch_unfold", 2)
            b = engine.sketch_unfolding(mode, l, s, rows0, vals)
            assert np.array_equal(a, b), mode
            assert np.linalg.norm(a - sampled.sketched_matricization(T, mode, l, rows0 + 1, vals, s)) <= 1e-12 * np.linalg.norm(a)
    finally:
        engine.set_option("sketch_unfold", 1)


def test_projected_update_matches_oracle(engine):
    T, cp, rng = problem((14, 12, 10), 5, 23)
    engine.set_tensor(T)
    engine.set_cpd(cp.factors, cp.lam)
    mode = 2
    piv = np.asfortranarray(np.stack([rng.integers(1, 15, size=90), rng.integers(1, 13, size=90)], axis=1).astype(np.int64))
    engine.set_projector(mode, piv)
    engine.drop_tensor()
    import itcpd
    with pytest.raises(itcpd.ItcpdError):
        engine.mttkrp(0)  # the dense tensor is gone, exactly like ALS(ITensor(inds(target)), ...)
    engine.projected_update(mode)
    K = sampled.pivot_hadamard([cp.factors[0], cp.factors[1]], piv)
    Ts = sampled.fused_flatten_sample(T, mode, piv)
    X = cpals.ldiv_solve(K.T @ K, np.asfortranarray((Ts @ K).T)).T
    Ao, lo = cpals.row_norm(X)
    assert np.linalg.norm(engine.get_factor(mode) - Ao) / np.linalg.norm(Ao) < 1e-10
    assert np.linalg.norm(engine.get_lambda() - lo) / np.linalg.norm(lo) < 1e-10


def test_pivot_projected_als_like_reference_tests(engine):
    """test/rand_cp_als.jl:28-96 scaled: QRPivProjected / SEQRCSPivProjected + update_samples bookkeeping; on an exactly
    low-rank tensor the projected solvers come within 10 % of the exact-ALS error."""
    import itcpd

    rng = np.random.default_rng(24)
    A = cpals.reconstruct(cpals.random_CPD((12, 13, 11), 4, rng))
    nA = np.linalg.norm(A)
    cp0 = cpals.random_CPD(A, 3, rng)
    start_cp = itcpd.CPD(cp0.factors, cp0.lam)
    exact = itcpd.als_optimize(A, start_cp, check=itcpd.CPDiffCheck(1e-5, 100), alg=itcpd.KRPNormal())
    e_exact = np.linalg.norm(A - itcpd.reconstruct(exact)) / nA

    als = itcpd.compute_als(A, start_cp, alg=itcpd.QRPivProjected(100), check=itcpd.CPDiffCheck(1e-5, 100), trunc_tol=4)
    als = itcpd.update_samples(A, als, 120, reshuffle=False)
    assert itcpd.stop(als.mttkrp_alg) == 120 and itcpd.start(als.mttkrp_alg) == 1
    assert type(als.mttkrp_alg) is itcpd.QRPivProjected
    assert als.additional_items["effective_ranks"][0] < 12
    itcpd.optimize(start_cp, als)

    for alg, kw in [(itcpd.QRPivProjected(140), {}),
                    (itcpd.SEQRCSPivProjected(1, 140, (1, 2, 3), (10, 10, 10)), dict(seed=3)),
                    (itcpd.SEQRCSPivProjected((1,), (140,), (1, 2, 3), (10, 10, 10)), dict(seed=4, injective=True, shuffle_pivots=False))]:
        ok = False
        for attempt in range(5):
            o = itcpd.als_optimize(A, start_cp, alg=alg, check=itcpd.CPDiffCheck(1e-5, 100), rng=np.random.default_rng(100 + attempt), **kw)
            e = np.linalg.norm(A - itcpd.reconstruct(o)) / nA
            if abs(e_exact - e) / e_exact < 0.1:
                ok = True
                break
        assert ok, (type(alg).__name__, e_exact, e)


@pytest.mark.parametrize("rankdef", [False, True])
def test_sampled_update_normal_false_matches_oracle_lstsq(engine, rankdef):
    """ProjectionAlgorithm.jl:63-64: qr(K, ColumnNorm()) \\ T_s' (pivoted-QR min-norm least squares on the tall sampled KRP)."""
    T, cp, rng = problem((14, 12, 10), 5, 25)
    f = [x.copy() for x in cp.factors]
    if rankdef:  # duplicate a column in both other factors -> K has two identical columns -> rank R-1
        f[0][:, 4] = f[0][:, 1]
        f[2][:, 4] = f[2][:, 1]
    engine.set_tensor(T)
    engine.set_cpd(f, cp.lam)
    mode = 1
    piv = np.asfortranarray(np.stack([rng.integers(1, 15, size=120), rng.integers(1, 11, size=120)], axis=1).astype(np.int64))
    engine.sampled_update(mode, piv, normal=False)
    K = sampled.pivot_hadamard([f[0], f[2]], piv)
    Ts = sampled.fused_flatten_sample(T, mode, piv)
    X = cpals.ldiv_solve(K, np.asfortranarray(Ts.T)).T      # oracle: LAPACK dgelsy, rcond = R * eps
    Ao, lo = cpals.row_norm(X)
    assert np.linalg.norm(engine.get_factor(mode) - Ao) / np.linalg.norm(Ao) < 1e-9
    assert np.linalg.norm(engine.get_lambda() - lo) / np.linalg.norm(lo) < 1e-9


def test_default_levscore_uses_normal_false_like_reference(engine):
    """optimizers/.../krp_lev_score_sampled.jl:7: LevScoreSampled defaults to normal=false."""
    import itcpd

    rng = np.random.default_rng(26)
    A = cpals.reconstruct(cpals.random_CPD((20, 18, 16), 3, rng))
    cp0 = cpals.random_CPD(A, 3, rng)
    als = itcpd.compute_als(A, itcpd.CPD(cp0.factors, cp0.lam), alg=itcpd.LevScoreSampled(300), check=itcpd.CPDiffCheck(1e-6, 60))
    assert als.additional_items["normal"] is False
    o = itcpd.optimize(itcpd.CPD(cp0.factors, cp0.lam), als)
    assert np.linalg.norm(A - itcpd.reconstruct(o)) / np.linalg.norm(A) < 5e-2
    o = itcpd.als_optimize(A, itcpd.CPD(cp0.factors, cp0.lam), alg=itcpd.QRPivProjected(200), normal=False, check=itcpd.CPDiffCheck(1e-6, 60))
    assert np.linalg.norm(A - itcpd.reconstruct(o)) / np.linalg.norm(A) < 5e-2


def test_device_seqrcs_krp_matches_oracle(engine):
    """SEQRCS.jl:184-241 (KRP-structured): same candidates / pivots as the oracle under the same rand() stream
    (test/SEQRCS_test.jl:49-83 compares it with the tensor version within 1e-2)."""
    rng = np.random.default_rng(30)
    dims, R = (12, 20, 25), 10
    cp = cpals.random_CPD(dims, R, rng)
    engine.generate_tensor(dims, seed=1)       # shape carrier: only the factors matter here
    engine.set_cpd(cp.factors, cp.lam)
    mode, l, s, t = 0, 120, 2, 10
    piv, rd, ncand = engine.seqrcs_krp(mode, l, s, t, seed=41)
    Q, Rm, p = sampled.seqrcs_krp([cp.factors[1], cp.factors[2]], l, s, t, which="ref", seed=41)
    n = 20 * 25
    assert sorted(piv.tolist()) == list(range(1, n + 1))
    assert set(piv[:ncand].tolist()) == set(p[:ncand].tolist()) and np.array_equal(piv[ncand:], p[ncand:])
    assert np.array_equal(piv[:R], p[:R])
    assert np.max(np.abs(np.abs(rd) - np.abs(np.diag(Rm)))) < 1e-10 * abs(Rm[0, 0])


def test_kseqrcs_projected_als(engine):
    """test/rand_cp_als.jl:85-96 scaled: KSEQRCSPivProjected on an exactly low-rank tensor (retry loop as in the reference)."""
    import itcpd

    rng = np.random.default_rng(31)
    A = cpals.reconstruct(cpals.random_CPD((12, 13, 11), 4, rng))
    nA = np.linalg.norm(A)
    cp0 = cpals.random_CPD(A, 3, rng)
    start_cp = itcpd.CPD(cp0.factors, cp0.lam)
    exact = itcpd.als_optimize(A, start_cp, check=itcpd.CPDiffCheck(1e-5, 100), alg=itcpd.KRPNormal())
    e_exact = np.linalg.norm(A - itcpd.reconstruct(exact)) / nA
    for kw in (dict(), dict(normal=False, injective=True)):
        ok = False
        for attempt in range(10):
            o = itcpd.als_optimize(A, start_cp, alg=itcpd.KSEQRCSPivProjected(1, (140,), (1, 2, 3), 5), check=itcpd.CPDiffCheck(1e-5, 100),
                                   rng=np.random.default_rng(200 + attempt), seed=attempt, **kw)
            e = np.linalg.norm(A - itcpd.reconstruct(o)) / nA
            if abs(e_exact - e) / e_exact < 0.1:
                ok = True
                break
        assert ok, (kw, e_exact, e)


def test_cpd_diff_terms_match_oracle(engine):
    """CPDiffCheck / CPAngleCheck scalars on the device (cp_cp_contract + norm_factors, cp_diff_check.jl:31-32)."""
    rng = np.random.default_rng(40)
    dims, R = (14, 9, 11, 5), 6
    prev = cpals.random_CPD(dims, R, rng)
    curr = cpals.random_CPD(dims, R, rng)
    prev.lam = rng.uniform(0.5, 2.0, R)
    curr.lam = rng.uniform(0.5, 2.0, R)
    engine.generate_tensor(dims, seed=2)
    engine.set_cpd(prev.factors, prev.lam)
    engine.cpd_snapshot()
    engine.set_cpd(curr.factors, curr.lam)
    inner, sq = engine.cpd_diff_terms()
    inner_o = float(prev.lam @ cpals.cp_cp_inner(prev.factors, curr.factors) @ curr.lam)
    sq_o = cpals.norm_factors([cpals.gram(f) for f in curr.factors], curr.lam)
    assert abs(inner - inner_o) < 1e-12 * max(1.0, abs(inner_o)) and abs(sq - sq_o) < 1e-12 * sq_o
    # explicit check against the dense tensors
    assert abs(inner - float(np.sum(cpals.reconstruct(prev) * cpals.reconstruct(curr)))) < 1e-10 * abs(inner_o) + 1e-12


def test_diff_check_trajectory_matches_oracle(engine):
    """The CPDiffCheck state machine driven by device scalars stops at the same sweep, with the same value, as the oracle's."""
    import itcpd

    T, cp, rng = problem((12, 13, 11), 4, 41)
    c1 = itcpd.CPDiffCheck(1e-4, 60)
    itcpd.als_optimize(T, itcpd.CPD(cp.factors, cp.lam), alg=itcpd.KRPNormal(), check=c1)
    c2 = cpals.CPDiffCheck(1e-4, 60)
    cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=c2)
    assert c1.total_iter == c2.total_iter and abs(c1.final_fit - c2.final_fit) < 1e-9
    a1 = itcpd.CPAngleCheck(1e-4, 60)
    itcpd.als_optimize(T, itcpd.CPD(cp.factors, cp.lam), alg=itcpd.KRPNormal(), check=a1)
    a2 = cpals.CPAngleCheck(1e-4, 60)
    cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=a2)
    assert a1.total_iter == a2.total_iter and abs(a1.final_fit - a2.final_fit) < 1e-6


def test_block_lev_score_sampled_als(engine):
    """test/rand_cp_als.jl:126-150 scaled: BlockLevScoreSampled with CPDiffCheck on an exactly low-rank tensor."""
    import itcpd

    rng = np.random.default_rng(42)
    A = cpals.reconstruct(cpals.random_CPD((24, 26, 22), 4, rng))
    nA = np.linalg.norm(A)
    cp0 = cpals.random_CPD(A, 3, rng)
    start = itcpd.CPD(cp0.factors, cp0.lam)
    exact = itcpd.als_optimize(A, start, check=itcpd.CPDiffCheck(1e-5, 100), alg=itcpd.KRPNormal())
    e_exact = np.linalg.norm(A - itcpd.reconstruct(exact)) / nA
    ok = False
    for attempt in range(5):
        o = itcpd.als_optimize(A, start, alg=itcpd.BlockLevScoreSampled(400, 4), normal=True, check=itcpd.CPDiffCheck(1e-5, 100), seed=attempt)
        e = np.linalg.norm(A - itcpd.reconstruct(o)) / nA
        if abs(e_exact - e) / e_exact < 0.1:
            ok = True
            break
    assert ok, (e_exact, e)
    piv = o and itcpd.package.host.block_sample_factor_matrices(20, [np.ones(8) / 8, np.ones(6) / 6, np.ones(5) / 5], 4, 1, np.random.default_rng(0))
    assert piv.shape == (20, 2) and piv.min() >= 1 and piv[:, 0].max() <= 8 and piv[:, 1].max() <= 5
    # rows of a block are consecutive in the blocked (first non-skipped) mode
    assert np.all(np.diff(piv[:4, 0]) == 1) and len(set(piv[:4, 1].tolist())) == 1


@pytest.mark.parametrize("normal", [True, False])
@pytest.mark.parametrize("dims,R,ns", [((30, 26, 22), 6, 120), ((40, 18, 12, 9), 5, (90, 64, 77, 100))])
def test_device_sampled_sweeps_equal_the_per_mode_calls(engine, dims, R, ns, normal):
    """itcpd_sampled_sweep_async (draws, gathers, sampled solves and leverage refresh with no host round trip, replayed from a CUDA
    graph, seeds from a device-side draw counter) against the hook-by-hook driver (itcpd_sample_factor_matrices + itcpd_sampled_update
    per mode): same seeds, same kernels -> bitwise equal factors, also across the plain-sweep -> graph-replay transition."""
    import itcpd

    rng = np.random.default_rng(3)
    A = np.asfortranarray(cpals.reconstruct(cpals.random_CPD(dims, R, rng)) + 0.01 * rng.standard_normal(dims))
    cp0 = cpals.random_CPD(dims, R, rng)
    out = {}
    for per_hook in (True, False):
        engine.set_tensor(A)
        als = itcpd.compute_als(engine, itcpd.CPD(cp0.factors, cp0.lam), alg=itcpd.LevScoreSampled(ns), normal=normal, check=itcpd.NoCheck(7), seed=4)
        als.additional_items["per_hook"] = per_hook
        launches0 = engine.launch_count
        o = itcpd.optimize(itcpd.CPD(cp0.factors, cp0.lam), als)
        out[per_hook] = (o, engine.launch_count - launches0, [engine.leverage_scores(n) for n in range(len(dims))])
    for a, b in zip(out[True][0].factors, out[False][0].factors):
        assert np.array_equal(a, b)
    assert np.array_equal(out[True][0].lam, out[False][0].lam)
    for a, b in zip(out[True][2], out[False][2]):
        assert np.array_equal(a, b)
    assert out[False][1] > 0
    assert all(np.all(np.isfinite(f)) for f in out[False][0].factors)


def test_sampled_mttkrp_kernel_matches_gather_then_multiply(engine):
    """sampled.cu: sampled_mttkrp_kernel (fibres read through the pivots, 64 x 64 register-tiled, split over sample chunks) against
    numpy on the gathered unfolding, for every mode, ragged sizes, more samples than one chunk"""
    T, cp, rng = problem((70, 33, 21), 37, 11)
    engine.set_tensor(T)
    engine.set_cpd(cp.factors, cp.lam)
    for mode in range(3):
        others = [m for m in range(3) if m != mode]
        nsamp = 1500 + 17 * mode
        piv = np.asfortranarray(np.stack([rng.integers(1, T.shape[m] + 1, size=nsamp) for m in others], axis=1).astype(np.int64))
        engine.sampled_update(mode, piv)
        K = sampled.pivot_hadamard([cp.factors[m] for m in others], piv)
        Ts = sampled.fused_flatten_sample(T, mode, piv)
        X = cpals.ldiv_solve(K.T @ K, np.asfortranarray((Ts @ K).T)).T
        Ao, lo = cpals.row_norm(X)
        assert np.linalg.norm(engine.get_factor(mode) - Ao) / np.linalg.norm(Ao) < 1e-9, mode
        engine.set_factor(mode, cp.factors[mode])
