"""CPU tests of the sampled-path oracle: re-statement of test/pivot_mapping.jl and test/SEQRCS_test.jl, the
bit-exact pin of the sparse-sign generators against the reference's own C (oracle/_ref), and the statistical
ALS tests of test/rand_cp_als.jl at reduced size."""
import json
import os

import numpy as np
import pytest

from oracle import cpals, sampled

HERE = os.path.dirname(os.path.abspath(__file__))
HAVE_REF = os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libsparse_sign_ref.so"))


def test_column_to_multi_coords_literal_pivots():
    """test/pivot_mapping.jl:4-28 and :30-53 -- the literal pivot lists of the reference, exact equality."""
    rng = np.random.default_rng(0)
    T = np.asfortranarray(rng.standard_normal((4, 5, 6, 3)))
    Tm = np.reshape(T, (4, 90), order="F")
    cols = [1, 4, 7, 12, 29, 30, 8, 21, 17, 42, 62, 86, 72]
    coords = sampled.column_to_multi_coords(cols, (5, 6, 3))
    for p, c in enumerate(cols):
        b, cc, d = coords[p] - 1
        assert np.array_equal(Tm[:, c - 1], T[:, b, cc, d])
    T = np.asfortranarray(rng.standard_normal((3, 6, 8, 2)))
    Tm = cpals.unfold(T, 2)
    cols = [1, 3, 7, 14, 29, 30, 10, 22, 35, 8, 11]
    coords = sampled.column_to_multi_coords(cols, (3, 6, 2))
    for p, c in enumerate(cols):
        a, b, d = coords[p] - 1
        assert np.array_equal(Tm[:, c - 1], T[a, b, :, d])
    assert np.array_equal(sampled.multi_coords_to_column((3, 6, 2), coords), np.array(cols))  # :105-107


def test_pivot_hadamard_and_fused_flatten_sample():  # test/pivot_mapping.jl:56-103
    rng = np.random.default_rng(1)
    A, B, C = rng.standard_normal((10, 7)), rng.standard_normal((15, 7)), rng.standard_normal((6, 7))
    K = cpals.khatri_rao([A, B, C])
    cols = rng.integers(1, 10 * 15 * 6 + 1, size=40)
    coords = sampled.column_to_multi_coords(cols, (10, 15, 6))
    assert np.array_equal(sampled.pivot_hadamard([A, B, C], coords), K[cols - 1, :])
    T = np.asfortranarray(rng.standard_normal((5, 10, 15, 6)))
    for k in range(4):
        rd = [T.shape[m] for m in range(4) if m != k]
        cols = rng.integers(1, int(np.prod(rd)) + 1, size=23)
        S = sampled.fused_flatten_sample(T, k, sampled.column_to_multi_coords(cols, rd))
        assert np.array_equal(S, cpals.unfold(T, k)[:, cols - 1])
        assert np.array_equal(sampled.fused_flatten_sample(T, k, cols), S)


def test_transform_alpha():  # pivot_mapping.jl:52-55
    T = np.arange(4 * 5 * 6, dtype=float).reshape((4, 5, 6), order="F")
    v = T.reshape(-1, order="F")
    # mode 1 (0-based) fibres: stride 4, extent 5; column alpha (1-based) of the unfolding starts at this position
    U = cpals.unfold(T, 1)
    for alpha in range(1, 25):
        pos = sampled.transform_alpha_to_vectorized_tensor_position(alpha, 5, 4)
        assert v[pos - 1] == U[0, alpha - 1]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (no reference sources on this box)")
@pytest.mark.parametrize("inj", [False, True])
@pytest.mark.parametrize("l,n,s", [(1200, 10000, 8), (50, 31, 2), (40, 62, 1), (7, 5, 9), (21294, 3000, 7)])
def test_sparse_sign_port_bit_exact_vs_reference_c(inj, l, n, s):
    a = sampled.sparse_sign_call(l, n, s, inj, "port", seed=2024)
    b = sampled.sparse_sign_call(l, n, s, inj, "ref", seed=2024)
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)  # NaN = entries the reference leaves unwritten


def test_sparse_sign_golden_fixture():
    g = json.load(open(os.path.join(HERE, "golden", "sparse_sign_ref.json")))
    for name, inj in (("sparse_sign", False), ("sparsestack", True)):
        e = g[name]
        vals, rows, cs = sampled.sparse_sign_call(e["l"], e["n"], e["s"], inj, "port", seed=e["seed"])
        assert rows.tolist() == e["rows"] and cs.tolist() == e["colstarts"]
        assert [int(np.sign(v)) if np.isfinite(v) else 0 for v in vals] == e["signs"]


def test_sparse_sign_structure():  # test/SEQRCS_test.jl:6-18
    m, n, s, l = 10, 10000, 8, 1200
    rows1, vals, om = sampled.sparse_sign_matrix(l, n, s, omega=True, seed=7)
    assert om.shape == (l, n)
    assert np.all(np.diff(om.tocsc().indptr) == s)
    assert np.all(np.isin(vals, [1 / np.sqrt(s), -1 / np.sqrt(s)]))
    # matrix-free sketch == omega * A (:20-28)
    rng = np.random.default_rng(3)
    A = rng.standard_normal((n, m))
    T = np.asfortranarray(A.T)  # 10 x 10000 "tensor"
    A_sk = sampled.sketched_matricization(T, 0, l, rows1, vals, s)
    assert np.linalg.norm(om @ A - A_sk.T) < 1e-10
    assert np.linalg.norm(sampled.sketched_matricization_omega(T, 0, om) - A_sk) < 1e-10
    ratio = np.linalg.svd(A_sk.T, compute_uv=False) / np.linalg.svd(A, compute_uv=False)
    assert np.all((0.5 <= ratio) & (ratio <= 1.5))


def test_sketched_matricization_all_modes():  # test/pivot_mapping.jl:110-119
    rng = np.random.default_rng(4)
    T = np.asfortranarray(rng.standard_normal((6, 7, 8)))
    for k in range(3):
        n = T.size // T.shape[k]
        rows1, vals, om = sampled.sparse_sign_matrix(20, n, 3, omega=True, seed=11 + k)
        dense = cpals.unfold(T, k) @ om.toarray().T
        assert np.linalg.norm(sampled.sketched_matricization(T, k, 20, rows1, vals, 3) - dense) < 1e-12
        assert np.linalg.norm(sampled.sketched_matricization_omega(T, k, om) - dense) < 1e-12


def test_omega_hadamard():  # test/pivot_mapping.jl:121-158
    rng = np.random.default_rng(5)
    A, B = rng.standard_normal((12, 6)), rng.standard_normal((9, 6))
    _, _, om = sampled.sparse_sign_matrix(30, 108, 4, omega=True, seed=13)
    assert np.linalg.norm(sampled.omega_hadamard([A, B], om) - om @ cpals.khatri_rao([A, B])) < 1e-12


def test_seqrcs_rank_k_error_close_to_qrcp():  # test/SEQRCS_test.jl:32-47 scaled (50 x 3000, k = 40)
    rng = np.random.default_rng(6)
    A = np.asfortranarray(rng.standard_normal((50, 3000)))
    k = 40
    Qa, Ra, pa = sampled.qrcp(A)
    err_act = np.linalg.norm(A[:, pa - 1] - Qa[:, :k] @ Ra[:k, :], 2) / np.linalg.norm(A, 2)
    for use_omega in (False, True):
        Q, R, p = sampled.seqrcs_tensor(A, 0, 750, 1, 40, use_omega=use_omega, seed=17)
        assert sorted(p.tolist()) == list(range(1, 3001))
        # complete R as the reference does with compute_r=true (SEQRCS.jl:124-130)
        nsub = R.shape[1]
        Rfull = np.hstack([R, Q.T @ A[:, p[nsub:] - 1]])
        err = np.linalg.norm(A[:, p - 1] - Q[:, :k] @ Rfull[:k, :], 2) / np.linalg.norm(A, 2)
        assert abs(err - err_act) <= 1e-2


def test_leverage_scores_sum_to_one():  # probability.jl:3-10
    rng = np.random.default_rng(7)
    A = rng.standard_normal((40, 6))
    p = sampled.compute_leverage_score_probability(A)
    assert abs(p.sum() - 1) < 1e-12 and np.all(p >= 0)
    s = sampled.samples_from_probability_vector(p, 2000, rng)
    assert s.min() >= 1 and s.max() <= 40
    blk = sampled.block_sample_factor_matrices(30, [p, p, p], 4, 1, rng)
    assert blk.shape == (30, 2) and blk.min() >= 1 and blk.max() <= 40


def test_sampled_als_statistical():  # test/rand_cp_als.jl:74-150 scaled
    rng = np.random.default_rng(8)
    # exactly low-rank target: sampled solvers must come within 10 % of exact-ALS error
    A = cpals.reconstruct(cpals.random_CPD((12, 13, 11), 4, rng))
    cp = cpals.random_CPD(A, 3, rng)
    chk = cpals.CPDiffCheck(1e-5, 100)
    opt = cpals.als_optimize(A, cp, alg=cpals.KRPNormal(), check=chk)
    exact = np.linalg.norm(A - cpals.reconstruct(opt)) / np.linalg.norm(A)
    for alg, kw in [(sampled.QRPivProjected(140), {}),
                    (sampled.SEQRCSPivProjected(1, 140, (1, 2, 3), (10, 10, 10)), {}),
                    (sampled.LevScoreSampled(140), dict(normal=True))]:
        ok = False
        for attempt in range(5):  # the reference wraps these in retry loops (rand_cp_als.jl:1-14)
            o = cpals.als_optimize(A, cp, alg=alg, check=cpals.CPDiffCheck(1e-5, 100), rng=np.random.default_rng(100 + attempt), **kw)
            err = np.linalg.norm(A - cpals.reconstruct(o)) / np.linalg.norm(A)
            if abs(exact - err) / exact < 0.1:
                ok = True
                break
        assert ok, type(alg).__name__


def test_update_samples_bookkeeping():  # test/rand_cp_als.jl:28-36
    rng = np.random.default_rng(9)
    T = np.asfortranarray(rng.standard_normal((8, 9, 10)))
    cp = cpals.random_CPD(T, 20, rng)
    als = cpals.compute_als(T, cp, alg=sampled.QRPivProjected(60), check=cpals.FitCheck(1e-6, 5, np.linalg.norm(T)), trunc_tol=4)
    als2 = sampled.update_samples(T, als, 70, reshuffle=False)
    assert als2.mttkrp_alg.End == 70 and als2.mttkrp_alg.Start == 1
    assert type(als2.mttkrp_alg) is sampled.QRPivProjected
    assert als.additional_items["effective_ranks"][0] < 8
    assert als2.additional_items["projects_tensors"][0].shape == (70, 2)
    cpals.optimize(cp, als2)
