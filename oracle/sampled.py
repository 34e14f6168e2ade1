"""Sampled / randomized CP-ALS oracle (leverage-score and pivot-projected solvers).

TEST INFRASTRUCTURE (see oracle/__init__.py).  **parity unpinned** for everything that draws from
Julia RNG streams (StatsBase.sample, randperm): those cannot be reproduced without Julia; the
sampling here uses numpy Generators with the SAME distribution.  The deterministic pieces (index
maps, gathers, sketches, sparse-sign generators) are exact restatements and are pinned by tests.

Index conventions follow the reference: pivot / sample matrices are int64, **1-based**,
shape (nsamp, N-1), column m holds the coordinate in the m-th *remaining* mode.
All citations are file:line into /root/reference.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Optional, Sequence

import numpy as np
from scipy.linalg import qr as _scipy_qr

from . import cpals
from .cpals import ALS, CPD, asf, ldiv_solve, unfold

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------------------
# index maps   (src/algebra/pivot_mapping.jl:5-55)
# --------------------------------------------------------------------------------------
def column_to_multi_coords(col_indices, dims) -> np.ndarray:
    """pivot_mapping.jl:17-35.  1-based column index of the unfolding -> 1-based coordinates,
    first remaining mode fastest."""
    cols = np.asarray(col_indices, dtype=np.int64)
    out = np.empty((cols.shape[0], len(dims)), dtype=np.int64, order="F")
    rem = cols - 1
    for d, dim in enumerate(dims):
        out[:, d] = rem % dim + 1
        rem = rem // dim
    return out


def multi_coords_to_column(sizes, pivots) -> np.ndarray:
    """pivot_mapping.jl:42-47 (inverse of the above)."""
    piv = np.asarray(pivots, dtype=np.int64)
    col = np.zeros(piv.shape[0], dtype=np.int64)
    stride = 1
    for d, dim in enumerate(sizes):
        col += (piv[:, d] - 1) * stride
        stride *= int(dim)
    return col + 1


def column_to_bc_coords(col_indices, b):  # pivot_mapping.jl:5-14
    return [((c - 1) % b + 1, (c - 1) // b + 1) for c in col_indices]


def transform_alpha_to_vectorized_tensor_position(alpha: int, extent: int, stride: int) -> int:
    """pivot_mapping.jl:52-55 (1-based alpha -> 1-based position of the fibre start in vec(T))."""
    t = (alpha - 1) // stride
    return t * stride * extent + (alpha - t * stride)


# --------------------------------------------------------------------------------------
# gathers   (pivot_mapping.jl:59-85, had_contract.jl:254-329)
# --------------------------------------------------------------------------------------
def fused_flatten_sample(T: np.ndarray, k: int, pivots: np.ndarray) -> np.ndarray:
    """pivot_mapping.jl:59-85: column s = mode-k fibre of T at the other-mode coordinates
    pivots[s,:] (1-based).  `pivots` may also be a vector of 1-based unfolding columns."""
    piv = np.asarray(pivots, dtype=np.int64)
    others = [m for m in range(T.ndim) if m != k]
    if piv.ndim == 1:
        piv = column_to_multi_coords(piv, [T.shape[m] for m in others])
    idx = [None] * T.ndim
    for c, m in enumerate(others):
        idx[m] = piv[:, c] - 1
    idx[k] = slice(None)
    # advanced indexing puts the sample axis first unless k separates them; normalise:
    out = np.empty((T.shape[k], piv.shape[0]), order="F")
    for s in range(piv.shape[0]):
        sel = tuple(slice(None) if m == k else int(idx[m][s]) for m in range(T.ndim))
        out[:, s] = T[sel]
    return out


def pivot_hadamard(factors: Sequence[np.ndarray], pivots: np.ndarray) -> np.ndarray:
    """had_contract.jl:277-295: K[s,r] = prod_m A_m[piv[s,m], r] (1-based piv)."""
    piv = np.asarray(pivots, dtype=np.int64)
    R = factors[0].shape[1]
    prod = np.ones((piv.shape[0], R), order="F")
    for A, col in zip(factors, piv.T):
        prod *= A[col - 1, :]
    return prod


def omega_hadamard(factors: Sequence[np.ndarray], omega) -> np.ndarray:
    """had_contract.jl:300-329: Omega (l x n, sparse) times the KRP of `factors`, row by row."""
    dims = [f.shape[0] for f in factors]
    R = factors[0].shape[1]
    om = omega.tocsr()
    out = np.empty((om.shape[0], R), order="F")
    for j in range(om.shape[0]):
        sl = slice(om.indptr[j], om.indptr[j + 1])
        nz = om.indices[sl] + 1
        coords = column_to_multi_coords(nz, dims)
        kr = pivot_hadamard(factors, coords)
        out[j, :] = np.sum(kr * om.data[sl][:, None], axis=0)
    return out


# --------------------------------------------------------------------------------------
# sparse-sign embeddings   (SEQRCS.jl:29-60 + the two C files)
# --------------------------------------------------------------------------------------
_libs = {}


def _load(which: str):
    if which not in _libs:
        path = {
            "port": os.path.join(_HERE, "_build", "liboracle_sparse.so"),
            "ref": os.path.join(_HERE, "_ref", "libsparse_sign_ref.so"),
        }[which]
        lib = ctypes.CDLL(path)
        _libs[which] = lib
    return _libs[which]


def sparse_sign_call(l: int, n: int, s: int, injective=False, which="port", seed: Optional[int] = None):
    """Raw call (0-based rows as the C code writes them).  `which`: 'port' (oracle restatement) or
    'ref' (the reference's own C, oracle/_ref).  `seed` -> srand(seed) first (same libc stream)."""
    lib = _load(which)
    s_eff = min(s, l)
    vals = np.full(n * s_eff, np.nan)
    rows = np.zeros(n * s_eff, dtype=np.int32)
    colstarts = np.zeros(n + 1, dtype=np.int32)
    if seed is not None:
        ctypes.CDLL(None).srand(ctypes.c_uint(seed))
    name = {("port", False): "oracle_sparse_sign", ("port", True): "oracle_sparsestack",
            ("ref", False): "sparse_sign", ("ref", True): "sparsestack"}[(which, bool(injective))]
    fn = getattr(lib, name)
    fn.restype = None
    fn.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 3
    fn(l, n, s, vals.ctypes.data, rows.ctypes.data, colstarts.ctypes.data)
    return vals, rows, colstarts


def sparse_sign_matrix(l: int, n: int, s: int, omega=False, injective=False, which="port", seed=None):
    """SEQRCS.jl:29-39: returns (rows 1-based Int32, vals) and, if omega, the l x n sparse matrix."""
    vals, rows, _ = sparse_sign_call(l, n, s, injective, which, seed)
    rows1 = rows + np.int32(1)
    if omega:
        import scipy.sparse as sp

        s_eff = min(s, l)
        cols = np.repeat(np.arange(n), s_eff)
        return rows1, vals, sp.csc_matrix((vals, (rows, cols)), shape=(l, n))
    return rows1, vals, None


def sketched_matricization_omega(T: np.ndarray, k: int, omega) -> np.ndarray:
    """pivot_mapping.jl:90-104: A_sk = T_(k) * Omega^T (I_k x l), sparse-matrix variant."""
    return np.asfortranarray((omega @ unfold(T, k).T).T)


def sketched_matricization(T: np.ndarray, k: int, l: int, rows1: np.ndarray, vals: np.ndarray, s: int) -> np.ndarray:
    """pivot_mapping.jl:111-140 (matrix-free variant): for every sketch row j, the signed sum of
    the mode-k fibres hashed to j, visited in increasing non-zero order."""
    Tk = unfold(T, k)  # I_k x n
    Ik = Tk.shape[0]
    out = np.zeros((Ik, l), order="F")
    order = np.argsort(rows1, kind="stable")  # nz positions grouped by sketch row, increasing
    cols = order // s  # nz q (0-based) belongs to unfolding column q // s
    r_sorted = rows1[order] - 1
    starts = np.searchsorted(r_sorted, np.arange(l + 1))
    for j in range(l):
        sl = slice(starts[j], starts[j + 1])
        if sl.stop > sl.start:
            out[:, j] = Tk[:, cols[sl]] @ vals[order[sl]]
    return out


# --------------------------------------------------------------------------------------
# leverage scores + sampling   (src/math_tools/probability.jl)
# --------------------------------------------------------------------------------------
def compute_leverage_score_probability(A: np.ndarray) -> np.ndarray:
    """probability.jl:3-10: thin QR w.r.t. the row index; p_i = sum_r Q[i,r]^2 / min(dims)."""
    q, _ = np.linalg.qr(A, mode="reduced")
    return np.sum(q * q, axis=1) / min(A.shape)


def samples_from_probability_vector(pw: np.ndarray, nsamples: int, rng) -> np.ndarray:
    """probability.jl:12-14: i.i.d. weighted draws with replacement (1-based)."""
    w = np.abs(pw)
    return rng.choice(len(w), size=nsamples, replace=True, p=w / w.sum()).astype(np.int64) + 1


def sample_factor_matrices(nsamps: int, skip_factor: int, probs, rng) -> np.ndarray:
    """probability.jl:23-33 (skip_factor 0-based here)."""
    out = np.empty((nsamps, len(probs) - 1), dtype=np.int64, order="F")
    c = 0
    for m in range(len(probs)):
        if m == skip_factor:
            continue
        out[:, c] = samples_from_probability_vector(probs[m], nsamps, rng)
        c += 1
    return out


def sample_single_col_from_factors(skip_factor: int, probs, rng):
    return [int(samples_from_probability_vector(p, 1, rng)[0]) for m, p in enumerate(probs) if m != skip_factor]


def block_sample_factor_matrices(nsamps: int, probs, block_size: int, skip_fact: int, rng) -> np.ndarray:
    """probability.jl:61-108 (skip_fact 0-based here)."""
    nf = len(probs)
    out = np.empty((nsamps, nf - 1), dtype=np.int64, order="F")
    blocked = probs[1 if skip_fact == 0 else 0]
    size_fast = len(blocked)
    nblocks = size_fast // block_size
    resid = size_fast % block_size
    block_prob = np.empty(nblocks)
    block_sizes = [1]
    m = 1
    for i in range(1, nblocks + 1):
        w = block_size + (1 if i <= resid else 0)
        block_prob[i - 1] = np.sum(blocked[m - 1: m - 1 + w])
        m += w
        block_sizes.append(m)
    m = 1
    for _ in range(nsamps // block_size):
        other = sample_single_col_from_factors(skip_fact, probs, rng)
        bs = int(samples_from_probability_vector(block_prob, 1, rng)[0])
        start_block, end_block = block_sizes[bs - 1], block_sizes[bs]
        for j in range(1, end_block - start_block + 1):
            if m > nsamps:
                m += 1
                break
            out[m - 1, :] = [start_block + j - 1] + other[1:]
            m += 1
    if m < nsamps + 1:
        for i in range(m, nsamps + 1):
            out[i - 1, :] = sample_single_col_from_factors(skip_fact, probs, rng)
    return out


# --------------------------------------------------------------------------------------
# QRCP helpers and SE-QRCS   (src/algebra/SEQRCS.jl)
# --------------------------------------------------------------------------------------
def qrcp(A: np.ndarray, want_q=True):
    """`qr(A, ColumnNorm())` = LAPACK dgeqp3.  Returns Q (m x min(m,n)) or None, R (min x n), p 1-based."""
    if want_q:
        Q, R, p = _scipy_qr(asf(A), mode="economic", pivoting=True)
        return Q, R, p.astype(np.int64) + 1
    R, p = _scipy_qr(asf(A), mode="r", pivoting=True)
    return None, R, p.astype(np.int64) + 1


def seqrcs_tensor(T: np.ndarray, mode: int, l: int, s: int, t: int, use_omega=False, injective=False,
                  which="port", seed=None, info: Optional[dict] = None):
    """SEQRCS.jl:89-134 (use_omega) / :139-182 (matrix-free); compute_r=false shape.
    Returns (Q, R, p) with p the 1-based pivot order over the n = P / I_mode unfolding columns."""
    n = T.size // T.shape[mode]
    s_eff = min(s, l)
    rows1, vals, om = sparse_sign_matrix(l, n, s, omega=use_omega, injective=injective, which=which, seed=seed)
    if use_omega:
        A_sk = sketched_matricization_omega(T, mode, om)
    else:
        A_sk = sketched_matricization(T, mode, l, rows1, vals, s_eff)
    _, _, p_sk = qrcp(A_sk, want_q=False)
    p_sk = p_sk[:t]
    rows_mat = np.reshape(rows1, (s_eff, n), order="F")
    if use_omega:
        hit = np.zeros(n, dtype=bool)
        hit[np.unique(om[p_sk - 1, :].nonzero()[1])] = True
        indices = np.nonzero(hit)[0].astype(np.int64) + 1  # findall over columns: increasing order
    else:
        # SEQRCS.jl:159: for each selected sketch row (in pivot order) the columns hashed to it, unique'd
        seen, indices = set(), []
        for pr in p_sk:
            for c in np.nonzero(np.any(rows_mat == pr, axis=0))[0]:
                if int(c) not in seen:
                    seen.add(int(c))
                    indices.append(int(c) + 1)
        indices = np.asarray(indices, dtype=np.int64)
    A_sub = fused_flatten_sample(T, mode, indices)
    Q, R, p_sub = qrcp(A_sub)
    rem = np.setdiff1d(np.arange(1, n + 1, dtype=np.int64), indices)
    p = np.concatenate([indices[p_sub - 1], rem])
    if info is not None:
        info.update(sketch_shape=A_sk.shape, subset=len(indices))
    return Q, R, p


def seqrcs_krp(krp: Sequence[np.ndarray], l: int, s: int, t: int, injective=False, which="port", seed=None):
    """SEQRCS.jl:184-241 (KRP-structured; compute_r=false)."""
    dims = [f.shape[0] for f in krp]
    n = int(np.prod(dims))
    rows1, vals, om = sparse_sign_matrix(l, n, s, omega=True, injective=injective, which=which, seed=seed)
    A_sk = omega_hadamard(krp, om)  # l x R
    _, _, p_sk = qrcp(np.asfortranarray(A_sk.T), want_q=False)
    p_sk = p_sk[:t]
    hit = np.unique(om[p_sk - 1, :].nonzero()[1])
    indices = hit.astype(np.int64) + 1
    coords = column_to_multi_coords(indices, dims)
    ff = pivot_hadamard(krp, coords)  # |ind| x R
    Q, R, p_sub = qrcp(np.asfortranarray(ff.T))
    rem = np.setdiff1d(np.arange(1, n + 1, dtype=np.int64), indices)
    p = np.concatenate([indices[p_sub - 1], rem])
    return Q, R, p


# --------------------------------------------------------------------------------------
# ProjectionAlgorithm hooks   (randomized/ProjectionAlgorithm.jl + krp_/qr_lev_score_sampled.jl)
# --------------------------------------------------------------------------------------
class ProjectionAlgorithm:
    def compute_krp(self, als, factors, cp, fact):  # ProjectionAlgorithm.jl:7-10
        portion = [f for m, f in enumerate(factors) if m != fact]
        return self.project_krp(als, portion, cp, fact)

    def solve_ls_problem(self, als, K, Ts):
        """ProjectionAlgorithm.jl:57-68.  K: nsamp x R, Ts: I_n x nsamp."""
        if als.additional_items["normal"]:
            X = ldiv_solve(K.T @ K, np.asfortranarray((Ts @ K).T))
        else:
            X = ldiv_solve(K, np.asfortranarray(Ts.T))
        return np.asfortranarray(X.T)

    def check_converge(self, converge, als, mtkrp, factors, lam, verbose=False) -> bool:
        """ProjectionAlgorithm.jl:15-54: FitCheck is not supported (runs max_counter sweeps)."""
        if isinstance(als.check, cpals.FitCheck):
            als.check.iter += 1
            if als.check.iter >= als.check.max_counter:
                als.check.iter = 0
            return False
        return converge.check_converge(factors, lam, [], verbose=verbose)


def _pick(v, fact):
    if isinstance(v, (tuple, list)):
        return v[0] if len(v) == 1 else v[fact]
    return v


class LevScoreSampled(ProjectionAlgorithm):  # algorithms/.../krp_lev_score_sampled.jl:9-58
    def __init__(self, nsamples=1):
        self.NSamples = tuple(nsamples) if isinstance(nsamples, (tuple, list)) else (int(nsamples),)

    def project_krp(self, als, portion, cp, fact):
        ai = als.additional_items
        nsamps = _pick(self.NSamples, fact)
        stop = ai["stop_resample"]
        if stop < 0 or stop > als.check.iter:
            ai["projects_tensors"][fact][...] = sample_factor_matrices(nsamps, fact, ai["factor_weights"], ai["rng"])
        return pivot_hadamard(portion, ai["projects_tensors"][fact])

    def matricize_tensor(self, als, factors, cp, fact):
        ai = als.additional_items
        if not ai["cache_sampled_targets"]:
            return fused_flatten_sample(als.target, fact, ai["projects_tensors"][fact])
        if als.check.iter <= ai["stop_resample"]:
            ai["sampled_targets"][fact] = fused_flatten_sample(als.target, fact, ai["projects_tensors"][fact])
        return ai["sampled_targets"][fact]

    def post_solve(self, als, factors, lam, cp, fact):
        als.additional_items["factor_weights"][fact] = compute_leverage_score_probability(factors[fact])


class BlockLevScoreSampled(ProjectionAlgorithm):  # krp_lev_score_sampled.jl:64-108
    def __init__(self, nsamples=0, blocks=1):
        self.NSamples = tuple(nsamples) if isinstance(nsamples, (tuple, list)) else (int(nsamples),)
        self.Blocks = tuple(blocks) if isinstance(blocks, (tuple, list)) else (int(blocks),)

    def project_krp(self, als, portion, cp, fact):
        ai = als.additional_items
        nsamps, bs = _pick(self.NSamples, fact), _pick(self.Blocks, fact)
        stop = ai["stop_resample"]
        if stop < 0 or stop > als.check.iter:
            ai["projects_tensors"][fact][...] = block_sample_factor_matrices(nsamps, ai["factor_weights"], bs, fact, ai["rng"])
        return pivot_hadamard(portion, ai["projects_tensors"][fact])

    def matricize_tensor(self, als, factors, cp, fact):
        return fused_flatten_sample(als.target, fact, als.additional_items["projects_tensors"][fact])

    def post_solve(self, als, factors, lam, cp, fact):
        als.additional_items["factor_weights"][fact] = compute_leverage_score_probability(factors[fact])


class _PivotBased(ProjectionAlgorithm):  # qr_lev_score_sampled.jl:86-168
    def __init__(self, start=1, end=0, random_modes=None, rank_vect=None):
        self.Start, self.End = start, end
        self.random_modes = None if random_modes is None else tuple(random_modes)  # 1-based mode numbers
        if rank_vect is not None and not isinstance(rank_vect, dict):
            rv = rank_vect if isinstance(rank_vect, (tuple, list)) else (rank_vect,) * len(self.random_modes)
            rank_vect = dict(zip(self.random_modes, rv))
        self.rank_vect = rank_vect

    def project_krp(self, als, portion, cp, fact):
        return pivot_hadamard(portion, als.additional_items["projects_tensors"][fact])

    def matricize_tensor(self, als, factors, cp, fact):
        return als.additional_items["target_transform"][fact]

    def post_solve(self, als, factors, lam, cp, fact):
        return None


class QRPivProjected(_PivotBased):
    def __init__(self, start_or_n=None, end=None):
        if start_or_n is None:
            super().__init__(1, 0)
        elif end is None:
            n = start_or_n
            super().__init__(tuple([1] * len(n)) if isinstance(n, (tuple, list)) else 1, n)
        else:
            super().__init__(start_or_n, end)


class SEQRCSPivProjected(_PivotBased):
    pass


class KSEQRCSPivProjected(_PivotBased):
    pass


def _range_for(alg, n, dRis):
    int_end = _pick(alg.End, n)
    int_end = dRis if int_end == 0 else int_end
    int_end = min(dRis, int_end)
    int_start = _pick(alg.Start, n)
    assert 0 < int_start <= int_end
    return int_start, int_end


def _finish_pivots(p, dr, trunc_tol, shuffle_pivots, rng):
    meff = int(np.sum(np.abs(dr) / np.max(np.abs(dr)) > trunc_tol))
    p1, rest = p[:meff], p[meff:]
    p2 = rest[rng.permutation(len(rest))] if shuffle_pivots else rest
    return meff, np.concatenate([p1, p2])


def compute_als_projection(alg, target, cp: CPD, extra, check, normal=None, stop_resample=-1,
                           cache_sampled_targets=True, shuffle_pivots=True, trunc_tol=0.01,
                           injective=False, guess_num_levs=None, prelim_niter=10, rng=None,
                           sketch_lib="port", **_):
    """The `compute_als` setups of optimizers/.../randomized/{krp,qr}_lev_score_sampled.jl."""
    rng = np.random.default_rng(7) if rng is None else rng
    N = target.ndim
    dims = target.shape
    if isinstance(alg, (LevScoreSampled, BlockLevScoreSampled)):
        extra["factor_weights"] = [compute_leverage_score_probability(f) for f in cp.factors]
        extra["rng"] = rng
        pts = []
        for fact in range(N):
            nsamps = _pick(alg.NSamples, fact)
            if isinstance(alg, LevScoreSampled):
                sc = sample_factor_matrices(nsamps, fact, extra["factor_weights"], rng)
            else:
                sc = block_sample_factor_matrices(nsamps, extra["factor_weights"], _pick(alg.Blocks, fact), fact, rng)
            pts.append(sc)
        extra["projects_tensors"] = pts
        extra["normal"] = False if normal is None else normal
        extra["stop_resample"] = stop_resample
        extra["sampled_targets"] = [None] * N
        extra["cache_sampled_targets"] = False if stop_resample == -1 else cache_sampled_targets
        return ALS(target, alg, extra, check)

    # pivot-based solvers
    lst = () if alg.random_modes is None else alg.random_modes
    updated = None
    if isinstance(alg, KSEQRCSPivProjected):
        prelim = 10 * cp.rank
        start_cp = cp if guess_num_levs is None else cpals.random_CPD(target, guess_num_levs, rng)
        updated = cpals.als_optimize(target, start_cp, alg=LevScoreSampled(prelim), check=cpals.NoCheck(prelim_niter),
                                     normal=True, stop_resample=0, rng=rng)
    ref_pivs, pivots, projectors, targets, qr_factors, eff = [], [], [], [], [], []
    for n in range(N):
        rdims = [dims[m] for m in range(N) if m != n]
        dRis = int(np.prod(rdims))
        int_start, int_end = _range_for(alg, n, dRis)
        m = dims[n]
        if (n + 1) in lst and not isinstance(alg, QRPivProjected):
            k_sk = int_end if alg.rank_vect is None else alg.rank_vect[n + 1]
            l = int(round(3 * m * math.log(m)))
            s = int(round(math.log(m)))
            if isinstance(alg, KSEQRCSPivProjected):
                q, r, p = seqrcs_krp([f for i, f in enumerate(updated.factors) if i != n], l, s, k_sk,
                                     injective=injective, which=sketch_lib)
            else:
                q, r, p = seqrcs_tensor(target, n, l, s, k_sk, use_omega=False, injective=injective, which=sketch_lib)
        elif isinstance(alg, KSEQRCSPivProjected):
            K = cpals.khatri_rao([f for i, f in enumerate(updated.factors) if i != n])
            q, r, p = qrcp(np.asfortranarray(K.T))
        else:
            q, r, p = qrcp(unfold(target, n))
        dr = np.diag(r).copy()
        ref_pivs.append(p.copy())
        meff, pp = _finish_pivots(p, dr, trunc_tol, shuffle_pivots, rng)
        eff.append(meff)
        coords = column_to_multi_coords(pp, rdims)
        pivots.append(coords)
        qr_factors.append(q[:, : len(dr)] * dr[None, :] if q is not None else None)
        proj = np.asfortranarray(coords[int_start - 1: int_end, :])
        projectors.append(proj)
        targets.append(fused_flatten_sample(target, n, proj))
    extra.update(ref_projectors=ref_pivs, projects=pivots, projects_tensors=projectors, target_transform=targets,
                 qr_factors=qr_factors, effective_ranks=eff, normal=True if normal is None else normal)
    return ALS(np.zeros((0,) * 0), alg, extra, check)  # qr_lev...:77 drops the tensor


def update_samples(target, als: ALS, new_num_end, reshuffle=False, new_num_start=0, rng=None) -> ALS:
    """algorithms/.../qr_lev_score_sampled.jl:95-149."""
    rng = np.random.default_rng(11) if rng is None else rng
    old = als.mttkrp_alg
    alg = type(old).__new__(type(old))
    _PivotBased.__init__(alg, old.Start if new_num_start == 0 else new_num_start,
                         old.End if new_num_end == 0 else new_num_end, old.random_modes, old.rank_vect)
    ai = als.additional_items
    N = target.ndim
    pivots = [p.copy() for p in ai["projects"]]
    projectors, targets = [], []
    for pos in range(N):
        rdims = [target.shape[m] for m in range(N) if m != pos]
        if reshuffle:
            p, meff = ai["ref_projectors"][pos], ai["effective_ranks"][pos]
            rest = p[meff:]
            pivots[pos] = column_to_multi_coords(np.concatenate([p[:meff], rest[rng.permutation(len(rest))]]), rdims)
        int_start, int_end = _range_for(alg, pos, int(np.prod(rdims)))
        proj = np.asfortranarray(pivots[pos][int_start - 1: int_end, :])
        projectors.append(proj)
        targets.append(fused_flatten_sample(asf(target), pos, proj))
    extra = dict(ai)
    extra.update(projects=pivots, projects_tensors=projectors, target_transform=targets)
    return ALS(als.target, alg, extra, als.check)
