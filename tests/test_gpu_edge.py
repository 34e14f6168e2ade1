"""Edge cases of the C-ABI on the GPU: argument errors, degenerate shapes, the pivot-tolerance boundary of the
pivoted Cholesky (dpstrf semantics, absolute tol), NaN propagation to the host-side `throw("Error NAN")`."""
import numpy as np
import pytest

from oracle import cpals

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_argument_errors_are_reported_not_crashes(engine):
    import itcpd

    rng = np.random.default_rng(0)
    T = np.asfortranarray(rng.standard_normal((6, 5, 4)))
    engine.set_tensor(T)
    cp = cpals.random_CPD(T, 3, rng)
    engine.set_cpd(cp.factors, cp.lam)
    with pytest.raises(itcpd.ItcpdError) as e:
        engine.mttkrp(3, fetch=False)          # mode out of range (checked by the library, not the wrapper)
    assert e.value.code == 2
    with pytest.raises(itcpd.ItcpdError):
        engine.solve(0)                        # solve before the MTTKRP of that mode
    with pytest.raises(itcpd.ItcpdError):
        engine.set_option("no_such_option", 1)
    with pytest.raises(itcpd.ItcpdError):
        engine.set_option("tile_warps", 5)
    with pytest.raises(itcpd.ItcpdError):
        engine.projected_update(1)             # no projector cached
    with pytest.raises(AssertionError):
        engine.set_factor(0, np.zeros((7, 3)))  # wrong shape is caught by the host mirror


@pytest.mark.parametrize("dims,R", [((1, 9, 8), 3), ((9, 1, 8), 3), ((9, 8, 1), 3), ((2, 2, 2), 1), ((3, 300), 2), ((17, 1), 1), ((5, 1, 1, 6), 2)])
def test_degenerate_shapes(engine, dims, R):
    """Extents of 1, rank 1, order 2: everything still matches the oracle."""
    rng = np.random.default_rng(1)
    T = np.asfortranarray(rng.standard_normal(dims))
    cp = cpals.random_CPD(T, R, rng)
    engine.set_tensor(T)
    engine.set_cpd(cp.factors, cp.lam)
    for n in range(len(dims)):
        assert relerr(engine.mttkrp(n), cpals.mttkrp_krp_normal(T, cp.factors, n)) < 1e-12
    engine.compute_grams()
    inner, norm2 = engine.sweep(4)
    chk = cpals.FitCheck(0.0, 4, float(np.linalg.norm(T)))
    cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=chk)
    nT = float(np.linalg.norm(T))
    fits = 1.0 - np.sqrt(np.abs(nT * nT + norm2 - 2 * np.abs(inner))) / nT
    # the fit formula cancels catastrophically when the model is (numerically) exact: |dfit| ~ sqrt(eps) there
    tol = np.where(np.array(chk.history) > 1 - 1e-6, 1e-6, 1e-9)
    assert np.all(np.abs(fits - np.array(chk.history)) < tol), (fits, chk.history)


def test_rank_above_one_r_block_and_large_rank_solve(engine):
    """R = 200 > 64: several r-blocks in the GEMM, E = 8 entries per lane in the row solve, Cholesky in shared memory;
    R = 300: the factorisation falls back to its global-memory workspace."""
    for R in (200, 300):
        rng = np.random.default_rng(2)
        T = np.asfortranarray(rng.standard_normal((40, 30, 20)))
        cp = cpals.random_CPD(T, R, rng)
        engine.set_tensor(T)
        engine.set_cpd(cp.factors, cp.lam)
        engine.compute_grams()
        inner, norm2 = engine.sweep(3)
        chk = cpals.FitCheck(0.0, 3, float(np.linalg.norm(T)))
        cpals.als_optimize(T, cp, alg=cpals.KRPNormal(), check=chk)
        nT = float(np.linalg.norm(T))
        fits = 1.0 - np.sqrt(np.abs(nT * nT + norm2 - 2 * np.abs(inner))) / nT
        assert np.max(np.abs(fits - np.array(chk.history))) < 1e-7, (R, fits, chk.history)


def test_cholesky_tolerance_boundary_matches_lapack_decision(engine):
    """Pivots straddling the absolute 1e-6 tolerance (cholesky_epsilon, src/ITensorCPD.jl:2): the device takes the
    Cholesky path exactly when LAPACK dpstrf(tol=1e-6) reports full rank, and the two solutions agree."""
    rng = np.random.default_rng(3)
    dims, R = (30, 25, 20), 6
    T = np.asfortranarray(rng.standard_normal(dims))
    base = cpals.random_CPD(T, R, rng)
    engine.set_tensor(T)
    decisions = []
    for eps in (1e-2, 3e-3, 2e-3, 1.2e-3, 8e-4, 5e-4, 1e-4, 1e-6):
        f = [x.copy() for x in base.factors]
        # make column 5 nearly parallel to column 0 in modes 0 and 2: Gamma_1 = G_0 * G_2 gets a smallest pivot ~ eps^2-ish
        for m in (0, 2):
            v = f[m][:, 0] + eps * f[m][:, 5]
            f[m][:, 5] = v / np.linalg.norm(v)
        engine.set_cpd(f, base.lam)
        engine.compute_grams()
        engine.gram_hadamard(1, fetch=False)
        engine.mttkrp(1, fetch=False)
        path, rank = engine.solve(1, 1e-6)
        grams = [cpals.gram(x) for x in f]
        info = {}
        Xo = cpals.solve_ls_problem(cpals.compute_krp_gram(grams, 1), cpals.mttkrp_krp_normal(T, f, 1), info)
        decisions.append((eps, path, info["path"]))
        assert (path == 0) == (info["path"] == "cholesky"), decisions
        if path == 0:
            assert rank == info["rank"] == R   # on the QRCP path the device reports the xGELSY rank, the oracle the dpstrf rank
        engine.normalize(1)
        Ao, _ = cpals.row_norm(Xo)
        cond = np.linalg.cond(cpals.compute_krp_gram(grams, 1))
        assert relerr(engine.get_factor(1), Ao) < 1e-12 * max(cond, 1e3), (eps, path, cond)
    assert {d[1] for d in decisions} == {0, 1}, decisions  # both branches were exercised


def test_last_solve_status_reports_the_path_taken(engine):
    """itcpd_last_solve_status: (path, rank) of the most recent R x R solve -- Cholesky for a full-rank Gram-Hadamard, the pivoted-QR
    fallback of ldiv_solve.jl:19-21 with the numerical rank when two columns are duplicated; after whole sweeps one slot per mode."""
    rng = np.random.default_rng(21)
    T = np.asfortranarray(rng.standard_normal((30, 25, 20)))
    cp = cpals.random_CPD(T, 12, rng)
    engine.set_tensor(T)
    engine.set_cpd(cp.factors, cp.lam)
    engine.compute_grams()
    engine.gram_hadamard(1, fetch=False); engine.mttkrp(1, fetch=False)
    assert engine.solve(1, 1e-6) == (0, 12) and engine.last_solve_status(0) == (0, 12)
    f = [x.copy() for x in cp.factors]
    for m in range(3):
        f[m][:, 7] = f[m][:, 2]
        f[m][:, 11] = f[m][:, 5]
    engine.set_cpd(f, cp.lam)
    engine.compute_grams()
    engine.gram_hadamard(1, fetch=False); engine.mttkrp(1, fetch=False)
    assert engine.solve(1, 1e-6) == (1, 10) and engine.last_solve_status(0) == (1, 10)
    engine.set_cpd(cp.factors, cp.lam)
    engine.compute_grams()
    engine.sweep(2)
    assert [engine.last_solve_status(m) for m in range(3)] == [(0, 12)] * 3


def test_nan_fit_raises_on_host_like_reference(engine):
    """row_norm has no zero guard (row_norm.jl:19-21): a zero column -> NaN -> throw("Error NAN") (fit_check.jl:40-42)."""
    import itcpd

    rng = np.random.default_rng(4)
    T = np.asfortranarray(rng.standard_normal((8, 7, 6)))
    T[:, :, :] = 0.0
    T[0, 0, 0] = 1.0   # rank-1 tensor with rank-3 model: some solved columns vanish -> 0/0
    cp = cpals.random_CPD(T, 3, rng)
    f = [x.copy() for x in cp.factors]
    f[1][:, 2] = 0.0
    chk = itcpd.FitCheck(1e-9, 30, 1.0)
    with pytest.raises(RuntimeError, match="Error NAN"):
        itcpd.als_optimize(T, itcpd.CPD(f, cp.lam), check=chk)


@pytest.mark.parametrize("dims", [(256, 250, 330), (255, 250, 331)])   # 169 / 168 MB: above the staging threshold; even and odd leading dimension
def test_pageable_tensor_upload_is_staged_and_exact(dims):
    """api.cu: upload_to_device -- a pageable host tensor (what a Julia Array is) goes through pinned staging buffers filled by four host
    threads; the device copy must be bit-for-bit the host array, for itcpd_set_tensor and for itcpd_als_from_host, and equal to what the
    single cudaMemcpyAsync path (staged_upload = 0) leaves."""
    import itcpd

    rng = np.random.default_rng(12)
    T = np.asfortranarray(rng.standard_normal(dims))
    with itcpd.Engine(0) as eng:
        eng.set_tensor(T)
        assert np.array_equal(eng.get_tensor(), T)
        n1 = eng.tensor_norm()
        eng.set_option("staged_upload", 0)
        eng.set_tensor(T)
        assert np.array_equal(eng.get_tensor(), T) and eng.tensor_norm() == n1
        eng.set_option("staged_upload", 1)
        f = [np.asfortranarray(x / np.linalg.norm(x, axis=0)) for x in (rng.standard_normal((d, 8)) for d in dims)]
        a = eng.als_from_host(T, f, 3)
        eng.set_option("staged_upload", 0)
        b = eng.als_from_host(T, f, 3)
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
