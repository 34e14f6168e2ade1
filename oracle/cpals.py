"""Dense CP-ALS oracle: numpy/scipy restatement of the reference's normal-equation path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parity status: **parity unpinned** at the
Julia third-party boundary (no golden vectors exist, Julia absent); LAPACK routines are the
ones Julia's LinearAlgebra calls.

Conventions: tensors are numpy arrays in Fortran (column-major) order,
first index fastest; factor A_n is (I_n, R); Gram G_n = A_n^T A_n; lambda is a length-R vector.
All citations are file:line into /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
from scipy.linalg import lapack as _lapack

CHOLESKY_EPSILON = 1e-6  # src/ITensorCPD.jl:2


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def asf(x) -> np.ndarray:
    """Float64, Fortran-ordered view/copy (Julia Array{Float64} layout)."""
    return np.asfortranarray(x, dtype=np.float64)


def unfold(T: np.ndarray, n: int) -> np.ndarray:
    """Mode-n unfolding T_(n): I_n x prod(others), other modes in original order, first fastest.

    This is `reshape(array(target, (i, Ris...)), (dim(i), dim(Ris)))`
    (src/optimizers/als_optimizers/randomized/qr_lev_score_sampled.jl:22).
    """
    N = T.ndim
    perm = [n] + [m for m in range(N) if m != n]
    return np.reshape(np.transpose(T, perm), (T.shape[n], -1), order="F")


def khatri_rao(mats: Sequence[np.ndarray]) -> np.ndarray:
    """KRP with the FIRST matrix's row index fastest: K[(i1,i2,..), r] = prod_m A_m[i_m, r].

    `had_contract(factors, rank)` of I_m x R matrices (src/algebra/had_contract.jl:72-124 with
    no rank-free tensor) yields the tensor (i_1, i_2, ..., r); flattening its leading indices
    column-major gives this matrix.
    """
    R = mats[0].shape[1]
    K = np.ones((1, R))
    for A in mats:
        # new index is slower than the existing ones
        K = (A[None, :, :] * K[:, None, :]).reshape(-1, R, order="F")
    return np.asfortranarray(K)


# --------------------------------------------------------------------------------------
# row_norm  (src/math_tools/row_norm.jl:4-24)
# --------------------------------------------------------------------------------------
def row_norm(X: np.ndarray):
    """Column-normalise X (I x R).  lambda_r = sqrt(sum_i X[i,r]^2); no zero guard (:19-21)."""
    X = asf(X)
    lam = np.sqrt(np.sum(X * X, axis=0))
    with np.errstate(divide="ignore", invalid="ignore"):
        A = X / lam[None, :]
    return np.asfortranarray(A), lam


# --------------------------------------------------------------------------------------
# CPD container  (src/cpd.jl:7-46)
# --------------------------------------------------------------------------------------
@dataclass
class CPD:
    factors: List[np.ndarray]
    lam: np.ndarray

    def __getitem__(self, i):  # cp[i] (0-based here); cp[()] is lambda   cpd.jl:28-29
        if i == () or i is None:
            return self.lam
        return self.factors[i]

    def __len__(self):
        return len(self.factors)

    @property
    def rank(self) -> int:  # cp_rank  cpd.jl:35
        return int(self.lam.shape[0])

    @property
    def dims(self):
        return tuple(int(f.shape[0]) for f in self.factors)

    def copy(self):
        return CPD([f.copy(order="F") for f in self.factors], self.lam.copy())


def random_factors(dims: Sequence[int], rank: int, rng=None):
    """cpd.jl:48-60.  Julia's MersenneTwister(3) stream cannot be reproduced without Julia;
    the default here is numpy `default_rng(3)` -- SAME SHAPE of computation (randn I x R per mode,
    drawn sequentially from one generator, column-normalised, lambda = norms of the LAST factor)."""
    rng = np.random.default_rng(3) if rng is None else rng
    facs, lam = [], None
    for I in dims:
        X = np.asfortranarray(rng.standard_normal((I, rank)))
        A, lam = row_norm(X)
        facs.append(A)
    return facs, lam


def random_CPD(target_or_dims, rank: int, rng=None) -> CPD:  # cpd.jl:63-82
    dims = target_or_dims.shape if hasattr(target_or_dims, "shape") else tuple(target_or_dims)
    f, l = random_factors(dims, rank, rng)
    return CPD(f, l)


def reconstruct(cp: CPD) -> np.ndarray:
    """src/algebra/reconstruct.jl:2-9: lambda * had_contract(factors, r)."""
    K = khatri_rao(cp.factors)  # (P, R)
    return np.reshape(K @ cp.lam, cp.dims, order="F")


# --------------------------------------------------------------------------------------
# MTTKRP  (src/algorithms/als_algorithms/standard/tensor.jl:10-49; had_contract.jl:72-124)
# --------------------------------------------------------------------------------------
def mttkrp_krp_normal(T: np.ndarray, factors: Sequence[np.ndarray], n: int) -> np.ndarray:
    """KRPNormal (tensor.jl:12-20): explicit KRP of the other factors then ONE contraction
    (GEMM I_n x prod(others) x R)."""
    others = [factors[m] for m in range(len(factors)) if m != n]
    K = khatri_rao(others)
    return np.asfortranarray(unfold(T, n) @ K)


def mttkrp_krp_free(T: np.ndarray, factors: Sequence[np.ndarray], n: int, ranks=None) -> np.ndarray:
    """KRPFreeNormal (tensor.jl:32-44 -> had_contract list version, had_contract.jl:72-124):
    for every r an independent tensor x (N-1 vectors) contraction; R passes over T.
    `ranks` restricts the r loop (used only to time a bounded sample of the work)."""
    N = T.ndim
    R = factors[0].shape[1]
    out = np.zeros((T.shape[n], R), order="F")
    rs = range(R) if ranks is None else ranks
    for r in rs:
        tmp = T
        # contract the trailing modes first (each contraction removes that axis)
        for m in range(N - 1, -1, -1):
            if m == n:
                continue
            tmp = np.tensordot(tmp, factors[m][:, r], axes=([m], [0]))
        out[:, r] = tmp
    return out


# --------------------------------------------------------------------------------------
# Gram-Hadamard  (MttkrpAlgorithm.jl:18-31) and Gram refresh (tensor.jl:46-49)
# --------------------------------------------------------------------------------------
def gram(A: np.ndarray) -> np.ndarray:
    return np.asfortranarray(A.T @ A)


def compute_krp_gram(grams: Sequence[np.ndarray], n: int) -> np.ndarray:
    R = grams[0].shape[0]
    G = np.ones((R, R), order="F")
    for i, g in enumerate(grams):
        if i == n:
            continue
        G = G * g
    return G


# --------------------------------------------------------------------------------------
# ldiv_solve  (src/algebra/ldiv_solve.jl:13-29)
# --------------------------------------------------------------------------------------
def _qrcp_minnorm_solve(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """`qr(A, ColumnNorm()) \\ B`: Julia's ldiv!(::QRPivoted, B, rcond=min(m,n)*eps) is the LAPACK
    xGELSY algorithm (geqp3 + laic1 rank estimate + tzrzf + ormrz), i.e. the min-norm solution."""
    m, n = A.shape
    rcond = min(m, n) * np.finfo(np.float64).eps
    nrhs = B.shape[1]
    Bp = np.zeros((max(m, n), nrhs), order="F")
    Bp[:m] = B
    lwork = max(1, 64 * (max(m, n) + nrhs + 3 * n + 1))
    v, x, j, rank, info = _lapack.dgelsy(asf(A).copy(order="F"), Bp, np.zeros(n, dtype=np.int32), rcond, lwork)
    if info != 0:
        raise np.linalg.LinAlgError(f"dgelsy info={info}")
    return np.asfortranarray(x[:n])


def ldiv_solve(A: np.ndarray, B: np.ndarray, info_out: Optional[dict] = None) -> np.ndarray:
    """Solve A X = B.  Square A: pivoted Cholesky (LAPACK dpstrf, upper, tol = 1e-6 absolute,
    `check=true` -> rank deficiency is an error) then permuted potrs; ANY failure silently falls
    back to column-pivoted QR least squares (ldiv_solve.jl:17-22).  Non-square: pivoted QR (:24)."""
    A = asf(A)
    B = asf(B)
    if A.shape[0] == A.shape[1]:
        n = A.shape[0]
        c, piv, rank, info = _lapack.dpstrf(A.copy(order="F"), tol=CHOLESKY_EPSILON, lower=0)
        if info == 0 and rank == n:
            p = piv - 1
            Bp = np.asfortranarray(B[p, :])
            x, info2 = _lapack.dpotrs(c, Bp, lower=0)
            if info2 == 0:
                X = np.empty_like(x, order="F")
                X[p, :] = x
                if info_out is not None:
                    info_out.update(path="cholesky", rank=int(rank))
                return X
        if info_out is not None:
            info_out.update(path="qrcp", rank=int(rank))
        return _qrcp_minnorm_solve(A, B)
    if info_out is not None:
        info_out.update(path="qrcp", rank=-1)
    return _qrcp_minnorm_solve(A, B)


def solve_ls_problem(Gamma: np.ndarray, M: np.ndarray, info_out=None) -> np.ndarray:
    """MttkrpAlgorithm.jl:34-41: X = (Gamma \\ M^T)^T, I x R column-major."""
    X = ldiv_solve(Gamma, np.asfortranarray(M.T), info_out)
    return np.asfortranarray(X.T)


# --------------------------------------------------------------------------------------
# Convergence checks  (src/converge_checks/*.jl)
# --------------------------------------------------------------------------------------
def norm_factors(grams: Sequence[np.ndarray], lam: np.ndarray) -> float:
    """converge_checks.jl:5-11: lambda^T (hadamard of all Grams) lambda."""
    had = grams[0].copy()
    for g in grams[1:]:
        had = had * g
    return float(lam @ had @ lam)


class ConvergeAlg:
    iter: int
    max_counter: int

    def save_mttkrp(self, M):  # converge_checks.jl:13
        return None


class NoCheck(ConvergeAlg):  # no_check.jl:1-20
    def __init__(self, maxiter: int):
        self.iter = 0
        self.max_counter = int(maxiter)
        self.lastfit = -1

    def check_converge(self, factors, lam, grams, verbose=False) -> bool:
        self.iter += 1
        if verbose:
            print(f"{lam.shape[0]}\t {self.iter}")
        if self.iter == self.max_counter:
            self.iter = 0
            return True
        return False


class FitCheck(ConvergeAlg):  # fit_check.jl:5-68
    def __init__(self, tol: float, maxiter: int, ref_norm: float):
        self.iter = 0
        self.counter = 0
        self.tolerance = tol
        self.max_counter = int(maxiter)
        self.ref_norm = float(ref_norm)
        self.MttKRP = None
        self.lastfit = 1.0
        self.final_fit = 0.0
        self.total_iter = 0
        self.history: List[float] = []  # oracle-side convenience, not in the reference

    def save_mttkrp(self, M):  # fit_check.jl:22
        self.MttKRP = M

    def fit_terms(self, factors, lam, grams):
        inner = float(np.sum(self.MttKRP * (factors[-1] * lam[None, :])))  # :28
        fact_square = norm_factors(grams, lam)  # :29
        return inner, fact_square

    def check_converge(self, factors, lam, grams, verbose=True) -> bool:
        self.iter += 1
        inner, fact_square = self.fit_terms(factors, lam, grams)
        return self.update(inner, fact_square, lam.shape[0], verbose)

    def update(self, inner: float, fact_square: float, R: int, verbose=False) -> bool:
        """State machine of fit_check.jl:30-65 given the two scalars."""
        norm_residual = math.sqrt(abs(self.ref_norm * self.ref_norm + fact_square - 2 * abs(inner)))
        curr_fit = 1.0 - norm_residual / self.ref_norm
        dfit = abs(self.lastfit - curr_fit)
        self.lastfit = curr_fit
        self.history.append(curr_fit)
        if verbose:
            print(f"{R}\t {self.iter} \t {curr_fit} \t {dfit}")
        if math.isnan(curr_fit):
            raise RuntimeError("Error NAN")  # :40-42
        if dfit < self.tolerance:
            self.counter += 1
            if self.counter >= 2:
                self.total_iter = self.iter
                self.iter = 0
                self.counter = 0
                self.final_fit = self.lastfit
                self.lastfit = 0
                return True
        else:
            self.counter = 0
        if self.iter >= self.max_counter:
            self.total_iter = self.iter
            self.iter = 0
            self.counter = 0
            self.final_fit = self.lastfit
            self.lastfit = 0
        return False


def CPDFit(check) -> float:  # fit_check.jl:68
    return check.final_fit


def cp_cp_inner(f1, f2) -> np.ndarray:
    """cp_cp_contract (src/algebra/cp_contract.jl:32-52) for two CPDs over the same indices:
    hadamard over modes of A_n^T B_n  (R1 x R2)."""
    inner = np.ones((f1[0].shape[1], f2[0].shape[1]))
    for a, b in zip(f1, f2):
        inner = inner * (a.T @ b)
    return inner


class _PrevCPCheck(ConvergeAlg):
    def _norm2(self, factors, lam):
        return norm_factors([gram(f) for f in factors], lam)

    def _finish(self):
        self.total_iter = self.iter
        self.iter = 0
        self.counter = 0
        self.final_fit = self._last()
        self._set_last(0)
        self.prev = None


class CPDiffCheck(_PrevCPCheck):  # cp_diff_check.jl:6-73
    def __init__(self, tol, maxiter):
        self.iter, self.counter, self.tolerance, self.max_counter = 0, 0, tol, int(maxiter)
        self.norm_prev_iter, self.prev, self.lastfit, self.final_fit, self.total_iter = 0.0, None, 1, 0, 0

    def _last(self):
        return self.lastfit

    def _set_last(self, v):
        self.lastfit = v

    def check_converge(self, factors, lam, grams, verbose=True) -> bool:
        self.iter += 1
        if self.prev is None:
            self.prev = ([f.copy() for f in factors], lam.copy())
            self.norm_prev_iter = self._norm2(factors, lam)
            return False
        pf, pl = self.prev
        inner = float(pl @ cp_cp_inner(pf, factors) @ lam)
        fact_square = self._norm2(factors, lam)
        resid = math.sqrt(abs(self.norm_prev_iter + fact_square - 2 * abs(inner)))
        curr_fit = 1.0 - resid / math.sqrt(abs(self.norm_prev_iter))
        dfit = abs(self.lastfit - curr_fit)
        self.lastfit = curr_fit
        self.prev = ([f.copy() for f in factors], lam.copy())
        self.norm_prev_iter = fact_square
        if verbose:
            print(f"{lam.shape[0]}\t {self.iter} \t {curr_fit} \t {dfit}")
        if dfit < self.tolerance:
            self.counter += 1
            if self.counter >= 2:
                self._finish()
                return True
        else:
            self.counter = 0
        if self.iter >= self.max_counter:
            self._finish()
        return False


class CPAngleCheck(_PrevCPCheck):  # cp_angle_check.jl:6-75
    def __init__(self, tol, maxiter):
        self.iter, self.counter, self.tolerance, self.max_counter = 0, 0, tol, int(maxiter)
        self.norm_prev_iter, self.prev, self.lastangle, self.final_fit, self.total_iter = 0.0, None, 1, 0, 0

    def _last(self):
        return self.lastangle

    def _set_last(self, v):
        self.lastangle = v

    def check_converge(self, factors, lam, grams, verbose=True) -> bool:
        self.iter += 1
        if self.prev is None:
            self.prev = ([f.copy() for f in factors], lam.copy())
            self.norm_prev_iter = math.sqrt(self._norm2(factors, lam))
            return False
        pf, pl = self.prev
        numer = float(pl @ cp_cp_inner(pf, factors) @ lam)
        norm_curr = math.sqrt(self._norm2(factors, lam))
        theta = numer / (norm_curr * self.norm_prev_iter)
        theta = 1.0 if theta > 1.0 else theta
        curr_angle = math.acos(theta)
        dfit = abs(self.lastangle - curr_angle)
        self.lastangle = curr_angle
        self.prev = ([f.copy() for f in factors], lam.copy())
        self.norm_prev_iter = norm_curr
        if verbose:
            print(f"{lam.shape[0]}\t {self.iter} \t {curr_angle} \t {dfit}")
        if dfit < self.tolerance:
            self.counter += 1
            if self.counter >= 2:
                self._finish()
                return True
        else:
            self.counter = 0
        if self.iter >= self.max_counter:
            self._finish()
        return False


# --------------------------------------------------------------------------------------
# Algorithm objects + ALS driver  (MttkrpAlgorithm.jl, tensor.jl, als_optimizer.jl, optimize.jl)
# --------------------------------------------------------------------------------------
class MttkrpAlgorithm:
    """The 5-hook contract of optimize.jl:19-30 for normal-equation solvers."""

    def compute_krp(self, als, factors, cp, fact):  # MttkrpAlgorithm.jl:18-31
        return compute_krp_gram(als.additional_items["part_grammian"], fact)

    def solve_ls_problem(self, als, krp, mtkrp):  # MttkrpAlgorithm.jl:34-41
        return solve_ls_problem(krp, mtkrp, als.additional_items.setdefault("solve_info", {}))

    def post_solve(self, als, factors, lam, cp, fact):  # tensor.jl:46-49 / :22-25
        als.additional_items["part_grammian"][fact] = gram(factors[fact])

    def check_converge(self, converge, als, mtkrp, factors, lam, verbose=False) -> bool:
        converge.save_mttkrp(mtkrp)  # MttkrpAlgorithm.jl:5-14
        return converge.check_converge(factors, lam, als.additional_items["part_grammian"], verbose=verbose)


class KRPNormal(MttkrpAlgorithm):  # tensor.jl:10-25
    def matricize_tensor(self, als, factors, cp, fact):
        return mttkrp_krp_normal(als.target, factors, fact)


class KRPFreeNormal(MttkrpAlgorithm):  # tensor.jl:30-49 (the default, als_optimizer.jl:45)
    def matricize_tensor(self, als, factors, cp, fact):
        return mttkrp_krp_free(als.target, factors, fact)


@dataclass
class ALS:  # als_optimizer.jl:5-10
    target: object
    mttkrp_alg: object
    additional_items: dict
    check: ConvergeAlg
    trace: list = field(default_factory=list)  # oracle-side convenience


def compute_als(target, cp: CPD, alg=None, check=None, maxiter=None, **kwargs) -> ALS:
    """als_optimizer.jl:37-55 + standard/tensor.jl:3-14 (+ dispatch to the sampled setups)."""
    alg = KRPFreeNormal() if alg is None else alg
    check = NoCheck(100 if maxiter is None else maxiter) if check is None else check
    extra = {"mttkrp_contract_sequences": [None] * target.ndim}
    if isinstance(alg, MttkrpAlgorithm):
        extra["part_grammian"] = [gram(f) for f in cp.factors]
        return ALS(asf(target), alg, extra, check)
    from . import sampled

    return sampled.compute_als_projection(alg, asf(target), cp, extra, check, **kwargs)


def optimize(cp: CPD, als: ALS, verbose=False, on_mode=None) -> CPD:
    """optimize.jl:6-35, line for line.  `on_mode(fact, mtkrp, factors, lam)` is an oracle-side
    observer used by the parity tests."""
    it = als.check.iter
    lam = cp.lam.copy()
    factors = [f.copy(order="F") for f in cp.factors]
    N = len(factors)
    converge = als.check
    alg = als.mttkrp_alg
    while it < converge.max_counter:
        mtkrp = None
        for fact in range(N):
            krp = alg.compute_krp(als, factors, cp, fact)
            mtkrp = alg.matricize_tensor(als, factors, cp, fact)
            solution = alg.solve_ls_problem(als, krp, mtkrp)
            factors[fact], lam = row_norm(solution)
            alg.post_solve(als, factors, lam, cp, fact)
            if on_mode is not None:
                on_mode(fact, mtkrp, factors, lam)
        if alg.check_converge(converge, als, mtkrp, factors, lam, verbose):
            break
        it += 1
    return CPD(factors, lam)


def als_optimize(target, cp: CPD, alg=None, check=None, maxiter=None, verbose=False, **kwargs) -> CPD:
    als = compute_als(target, cp, alg=alg, check=check, maxiter=maxiter, **kwargs)  # als_optimizer.jl:15-25
    return optimize(cp, als, verbose=verbose)


def decompose(A, rank: int, solver=None, rng=None, alg=None, check=None, maxiter=None, verbose=False, **kw) -> CPD:
    """decompose.jl:13-30."""
    cp = random_CPD(A, rank, rng)
    if solver is None:
        return als_optimize(A, cp, alg=alg, check=check, maxiter=maxiter, verbose=verbose, **kw)
    raise RuntimeError("OptimizerError")


def increase_cpd_rank(cp: CPD, new_rank: int, rng=None) -> CPD:  # decompose.jl:71-82
    rng = np.random.default_rng(3) if rng is None else rng
    assert new_rank >= cp.rank
    newf, lam = random_factors(cp.dims, new_rank, rng)
    for old, new in zip(cp.factors, newf):
        new[:, : cp.rank] = old
    return CPD(newf, lam)


def decompose_adaptive(A, epsilon, max_rank, rng=None, alg=None, check=None, maxiter=None, verbose=False,
                       start_rank=1, rank_step=1, **kw) -> CPD:
    """decompose.jl:32-69."""
    current = start_rank
    cp = random_CPD(A, start_rank, rng)
    check = FitCheck(1e-3, 100, float(np.linalg.norm(A))) if check is None else check
    while True:
        cp = als_optimize(A, cp, alg=alg, check=check, maxiter=maxiter, verbose=verbose, **kw)
        check.iter = 0
        if 1.0 - CPDFit(check) < epsilon:
            return cp
        current += rank_step
        if current > max_rank:
            print(f"Optimization Failed to converge within rank {max_rank}")
            return cp
        cp = increase_cpd_rank(cp, current, rng)
