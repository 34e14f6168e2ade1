"""Host-side mirror of ITensorCPD.jl's public API for the decomposition path, driving the C-ABI.

Same names, argument meaning and error behaviour as the reference (file:line into the reference):
`decompose` (src/decompose.jl:1-69), `als_optimize` / `ALS` / `compute_als`
(src/optimizers/als_optimizers/als_optimizer.jl:5-69), `optimize` (optimize.jl:6-35), `random_CPD`
(src/cpd.jl:63-82), `CPD` with `cp[n]` / `cp[()]` (cpd.jl:7-46), `reconstruct`
(src/algebra/reconstruct.jl:2-9), `FitCheck` / `NoCheck` / `CPDiffCheck` / `CPAngleCheck`
(src/converge_checks/*.jl), algorithm objects `KRPFreeNormal`, `KRPNormal`, `LevScoreSampled`.

This is what the Julia package extension does (julia/ext/ITCPDB200Ext.jl): Julia is not installed in
this image, so the same control flow is written in Python.  Everything numeric happens in
libitcpd_b200.so on the GPU; the only host arithmetic is the scalar convergence state machine,
exactly as in the reference.  Python indices are 0-based (modes 0..N-1); sample / pivot matrices
keep the reference's 1-based int64 convention.
"""
from __future__ import annotations

import math
from typing import List

import numpy as np

from .engine import Engine

CHOLESKY_EPSILON = 1e-6  # src/ITensorCPD.jl:2


# ------------------------------------------------------------------------------------------
# CPD container  (cpd.jl:7-46)
# ------------------------------------------------------------------------------------------
class CPD:
    def __init__(self, factors: List[np.ndarray], lam: np.ndarray):
        self.factors = [np.asfortranarray(f, dtype=np.float64) for f in factors]
        self.lam = np.ascontiguousarray(lam, dtype=np.float64)

    def __getitem__(self, i):
        if i == () or i is None:  # cp[] in Julia
            return self.lam
        return self.factors[i]

    def __len__(self):
        return len(self.factors)

    def __iter__(self):
        return iter(self.factors)

    @property
    def rank(self) -> int:
        return int(self.lam.shape[0])

    @property
    def dims(self):
        return tuple(int(f.shape[0]) for f in self.factors)

    def copy(self):
        return CPD([f.copy(order="F") for f in self.factors], self.lam.copy())

    def __eq__(self, other):
        return all(np.array_equal(a, b) for a, b in zip(self.factors, other.factors)) and np.array_equal(self.lam, other.lam)


def cp_rank(cp: CPD) -> int:
    return cp.rank


def random_factors(dims, rank: int, rng=None):
    """cpd.jl:48-60: randn(I_n, R) per mode from ONE generator, column-normalised; lambda = norms of
    the last factor.  (Default stream: numpy default_rng(3) stands in for MersenneTwister(3).)"""
    rng = np.random.default_rng(3) if rng is None else rng
    facs, lam = [], None
    for I in dims:
        X = np.asfortranarray(rng.standard_normal((int(I), int(rank))))
        lam = np.sqrt(np.sum(X * X, axis=0))
        facs.append(np.asfortranarray(X / lam[None, :]))
    return facs, lam


def random_CPD(target, rank: int, rng=None) -> CPD:
    dims = target.shape if hasattr(target, "shape") else tuple(target)
    f, l = random_factors(dims, rank, rng)
    return CPD(f, l)


# ------------------------------------------------------------------------------------------
# convergence checks (host state machines, src/converge_checks/*.jl)
# ------------------------------------------------------------------------------------------
class ConvergeAlg:
    iter = 0
    max_counter = 0


class NoCheck(ConvergeAlg):  # no_check.jl:1-20
    def __init__(self, maxiter: int):
        self.iter = 0
        self.max_counter = int(maxiter)
        self.lastfit = -1

    def check_converge(self, als, factors_lam, verbose=False) -> bool:
        self.iter += 1
        if verbose:
            print(f"{als.engine.rank}\t {self.iter}")
        if self.iter == self.max_counter:
            self.iter = 0
            return True
        return False


class FitCheck(ConvergeAlg):  # fit_check.jl:5-68
    def __init__(self, tol, maxiter, ref_norm):
        self.iter = 0
        self.counter = 0
        self.tolerance = tol
        self.max_counter = int(maxiter)
        self.ref_norm = float(ref_norm)
        self.lastfit = 1.0
        self.final_fit = 0.0
        self.total_iter = 0
        self.history: List[float] = []

    def update(self, inner: float, fact_square: float, R: int, verbose=False) -> bool:
        """fit_check.jl:25-65 given <T,That> and ||That||^2 from the device (itcpd_fit_terms)."""
        self.iter += 1
        resid = math.sqrt(abs(self.ref_norm * self.ref_norm + fact_square - 2 * abs(inner)))
        curr = 1.0 - resid / self.ref_norm
        dfit = abs(self.lastfit - curr)
        self.lastfit = curr
        self.history.append(curr)
        if verbose:
            print(f"{R}\t {self.iter} \t {curr} \t {dfit}")
        if math.isnan(curr):
            raise RuntimeError("Error NAN")
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

    def check_converge(self, als, factors_lam, verbose=False) -> bool:
        inner, norm2 = als.engine.fit_terms()
        return self.update(inner, norm2, als.engine.rank, verbose)


def CPDFit(check) -> float:
    return check.final_fit


class _PrevCheck(ConvergeAlg):
    """CPDiffCheck / CPAngleCheck compare the CPDs of consecutive sweeps through their factor matrices only
    (cp_diff_check.jl:20-71, cp_angle_check.jl:20-73); they are the stopping rules of the sampled solvers (FitCheck is
    disabled there).  The two scalars they need -- <That_prev, That_curr> and ||That_curr||^2 -- come from the device
    (itcpd_cpd_snapshot / itcpd_cpd_diff_terms); the state machine below is the reference's, verbatim."""

    def __init__(self, tol, maxiter):
        self.iter, self.counter, self.tolerance, self.max_counter = 0, 0, tol, int(maxiter)
        self.norm_prev_iter, self.prev, self.final_fit, self.total_iter = 0.0, None, 0, 0
        self._lastv = 1

    def _finish(self):
        self.total_iter = self.iter
        self.iter = 0
        self.counter = 0
        self.final_fit = self._lastv
        self._lastv = 0
        self.prev = None

    def check_converge(self, als, factors_lam, verbose=False) -> bool:
        eng = als.engine
        self.iter += 1
        if self.prev is None:
            eng.cpd_snapshot()                      # check.PrevCP = CPD(factors, lambda)
            _, sq = eng.cpd_diff_terms()
            self.prev = True
            self.norm_prev_iter = self._first_norm(sq)
            return False
        inner, sq = eng.cpd_diff_terms()
        val = self._measure(inner, sq)
        d = abs(self._lastv - val)
        self._lastv = val
        eng.cpd_snapshot()
        if verbose:
            print(f"{eng.rank}\t {self.iter} \t {val} \t {d}")
        if d < self.tolerance:
            self.counter += 1
            if self.counter >= 2:
                self._finish()
                return True
        else:
            self.counter = 0
        if self.iter >= self.max_counter:
            self._finish()
        return False


class CPDiffCheck(_PrevCheck):
    @property
    def lastfit(self):
        return self._lastv

    def _first_norm(self, sq):
        return sq

    def _measure(self, inner, sq):
        resid = math.sqrt(abs(self.norm_prev_iter + sq - 2 * abs(inner)))
        fit = 1.0 - resid / math.sqrt(abs(self.norm_prev_iter))
        self.norm_prev_iter = sq
        return fit


class CPAngleCheck(_PrevCheck):
    @property
    def lastangle(self):
        return self._lastv

    def _first_norm(self, sq):
        return math.sqrt(sq)

    def _measure(self, inner, sq):
        nc = math.sqrt(sq)
        theta = min(1.0, inner / (nc * self.norm_prev_iter))
        self.norm_prev_iter = nc
        return math.acos(theta)


# ------------------------------------------------------------------------------------------
# algorithm objects: the 5-hook contract of optimize.jl:19-30, each hook one C-ABI call
# ------------------------------------------------------------------------------------------
class MttkrpAlgorithm:
    """Normal-equation solvers (algorithms/.../standard/MttkrpAlgorithm.jl)."""

    mttkrp_alg = 0  # ITCPD_MTTKRP_TREE

    def compute_krp(self, als, fact):  # MttkrpAlgorithm.jl:18-31
        als.engine.gram_hadamard(fact, fetch=False)

    def matricize_tensor(self, als, fact):  # tensor.jl:12-44
        als.engine.mttkrp(fact, fetch=False)

    def solve_ls_problem(self, als, fact):  # MttkrpAlgorithm.jl:34-41 -> ldiv_solve.jl:13-29
        als.solve_paths.append(als.engine.solve(fact, CHOLESKY_EPSILON))

    def post_solve(self, als, fact):  # tensor.jl:46-49
        als.engine.post_solve(fact)

    def check_converge(self, als, verbose=False) -> bool:  # MttkrpAlgorithm.jl:5-14
        return als.check.check_converge(als, als.fetch_cpd_arrays, verbose=verbose)


class KRPFreeNormal(MttkrpAlgorithm):
    """The reference default (als_optimizer.jl:45).  On the device both reference MTTKRP
    formulations are served by the same dimension-tree GEMM: identical M, one tensor pass per half sweep."""


class KRPNormal(MttkrpAlgorithm):
    pass


class DirectNormal(MttkrpAlgorithm):
    """Cross-check variant: one plain-FMA pass over T per mode (ITCPD_MTTKRP_DIRECT)."""

    mttkrp_alg = 1


class ProjectionAlgorithm:
    pass


class LevScoreSampled(ProjectionAlgorithm):
    """algorithms/.../randomized/krp_lev_score_sampled.jl:9-58 with the device doing the leverage
    scores, the weighted sampling, both gathers and the sampled normal-equation solve."""

    def __init__(self, nsamples=1):
        self.NSamples = tuple(nsamples) if isinstance(nsamples, (tuple, list)) else (int(nsamples),)

    def nsamples(self, fact):
        return self.NSamples[0] if len(self.NSamples) == 1 else self.NSamples[fact]

    def compute_krp(self, als, fact):
        ai = als.additional_items
        stop = ai["stop_resample"]
        if stop < 0 or stop > als.check.iter or ai["projects_tensors"][fact] is None:
            ai["seed"] += 1
            ai["projects_tensors"][fact] = als.engine.sample_factor_matrices(fact, self.nsamples(fact), ai["seed"])

    def matricize_tensor(self, als, fact):
        pass  # gathered inside sampled_update together with the sampled KRP

    def solve_ls_problem(self, als, fact):
        als.engine.sampled_update(fact, als.additional_items["projects_tensors"][fact], CHOLESKY_EPSILON,
                                  normal=als.additional_items["normal"])

    def post_solve(self, als, fact):
        pass  # leverage refresh is part of sampled_update (krp_lev...:55-58)

    def check_converge(self, als, verbose=False) -> bool:  # ProjectionAlgorithm.jl:15-54
        if isinstance(als.check, FitCheck):
            if als.check.iter == 0:
                print(f"Warning: FitCheck is not enabled for {type(self).__name__} will run {als.check.max_counter} iterations.")
            als.check.iter += 1
            if als.check.iter >= als.check.max_counter:
                als.check.iter = 0
            return False
        return als.check.check_converge(als, als.fetch_cpd_arrays, verbose=verbose)


def block_sample_factor_matrices(nsamps, probs, block_size, skip_fact, rng):
    """math_tools/probability.jl:61-108 (host-side integer logic, exactly as in the reference): the first non-skipped
    mode is cut into near-equal blocks of `block_size` consecutive rows, a block is drawn with the summed leverage of its
    rows and expanded into all its rows, the other modes get one weighted draw per block.  1-based int64 output."""
    def draw(p):
        w = np.abs(np.asarray(p, dtype=np.float64))
        return int(rng.choice(len(w), p=w / w.sum())) + 1

    def one_per_mode():
        return [draw(p) for m, p in enumerate(probs) if m != skip_fact]

    nf = len(probs)
    out = np.empty((nsamps, nf - 1), dtype=np.int64, order="F")
    blocked = np.asarray(probs[1 if skip_fact == 0 else 0])
    nblocks, resid = len(blocked) // block_size, len(blocked) % block_size
    edges = [1]
    for i in range(1, nblocks + 1):
        edges.append(edges[-1] + block_size + (1 if i <= resid else 0))
    block_prob = [float(np.sum(blocked[edges[i] - 1: edges[i + 1] - 1])) for i in range(nblocks)]
    m = 1
    for _ in range(nsamps // block_size):
        other = one_per_mode()
        b = draw(block_prob)
        for j in range(1, edges[b] - edges[b - 1] + 1):
            if m > nsamps:
                m += 1
                break
            out[m - 1, :] = [edges[b - 1] + j - 1] + other[1:]
            m += 1
    for i in range(m, nsamps + 1):
        out[i - 1, :] = one_per_mode()
    return out


class BlockLevScoreSampled(LevScoreSampled):
    """algorithms/.../randomized/krp_lev_score_sampled.jl:64-108: leverage-score sampling in blocks of consecutive rows.
    The leverage scores come from the device; the block bookkeeping is host integer logic like in the reference; gathers,
    the sampled solve and the leverage refresh run on the device (itcpd_sampled_update)."""

    def __init__(self, nsamples=0, blocks=1):
        super().__init__(nsamples)
        self.Blocks = tuple(blocks) if isinstance(blocks, (tuple, list)) else (int(blocks),)

    def compute_krp(self, als, fact):
        ai = als.additional_items
        stop = ai["stop_resample"]
        if stop < 0 or stop > als.check.iter or ai["projects_tensors"][fact] is None:
            probs = [als.engine.leverage_scores(n) for n in range(len(als.engine.dims))]
            bs = self.Blocks[0] if len(self.Blocks) == 1 else self.Blocks[fact]
            ai["projects_tensors"][fact] = block_sample_factor_matrices(self.nsamples(fact), probs, bs, fact, ai["rng"])


def _pick(v, fact):
    if isinstance(v, (tuple, list)):
        return v[0] if len(v) == 1 else v[fact]
    return v


class _PivotBased(ProjectionAlgorithm):
    """QRPivProjected / SEQRCSPivProjected (algorithms/.../randomized/qr_lev_score_sampled.jl:10-168): the projector
    of every mode is fixed at setup (column-pivoted QR / SE-QRCS of the unfolding, run on the device), the sampled
    target T_s is gathered once and cached on the device; each sweep only forms the sampled KRP and solves."""

    def __init__(self, start=1, end=0, random_modes=None, rank_vect=None):
        self.Start, self.End = start, end
        self.random_modes = None if random_modes is None else tuple(random_modes)  # 1-based mode numbers, as in Julia
        if rank_vect is not None and not isinstance(rank_vect, dict):
            rv = rank_vect if isinstance(rank_vect, (tuple, list)) else (rank_vect,) * len(self.random_modes)
            rank_vect = dict(zip(self.random_modes, rv))
        self.rank_vect = rank_vect

    def compute_krp(self, als, fact):
        pass  # pivot_hadamard happens inside projected_update (qr_lev...:151-160)

    def matricize_tensor(self, als, fact):
        pass  # cached target_transform[fact] (qr_lev...:162-166)

    def solve_ls_problem(self, als, fact):
        als.engine.projected_update(fact, CHOLESKY_EPSILON, normal=als.additional_items["normal"])

    def post_solve(self, als, fact):
        pass  # qr_lev...:168

    check_converge = LevScoreSampled.check_converge


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
    """qr_lev_score_sampled.jl:61-83: the pivots come from an SE-QRCS of the Khatri-Rao product of preliminary factors
    (a short leverage-score sampled ALS), not of the target tensor."""


def start(alg):
    return alg.Start


def stop(alg):
    return alg.End


def _proj_range(alg, n, dRis):
    int_end = _pick(alg.End, n)
    int_end = dRis if int_end == 0 else int_end
    int_end = min(dRis, int_end)
    int_start = _pick(alg.Start, n)
    assert 0 < int_start <= int_end
    return int_start, int_end


def _setup_pivot_based(alg, eng: Engine, cp: CPD, check, normal, shuffle_pivots, trunc_tol, injective, rng, seed,
                       guess_num_levs=None, prelim_niter=10, owns_tensor=True) -> "ALS":
    """optimizers/.../randomized/qr_lev_score_sampled.jl:1-78 (QRPivProjected), :80-176 (SEQRCSPivProjected) and
    :178-282 (KSEQRCSPivProjected)."""
    from .engine import column_to_multi_coords

    rng = np.random.default_rng(7) if rng is None else rng
    dims = eng.dims
    N = len(dims)
    lst = () if alg.random_modes is None else alg.random_modes
    ref_pivs, pivots, projectors, eff = [], [], [], []
    krp_mode = isinstance(alg, KSEQRCSPivProjected)
    if krp_mode:
        # preliminary leverage-score sampled ALS (:193-202): its factors' KRP stands in for the tensor's unfoldings
        prelim = 10 * cp.rank
        start_cp = cp if guess_num_levs is None else random_CPD(dims, guess_num_levs, rng)
        updated = als_optimize(eng, start_cp, alg=LevScoreSampled(prelim), check=NoCheck(prelim_niter), normal=True,
                               stop_resample=0, seed=0 if seed is None else seed)
        eng.set_cpd(updated.factors, updated.lam)  # the device holds the preliminary factors during the pivot search
    def sketch_params(n):
        m = dims[n]
        int_end = _proj_range(alg, n, int(np.prod([dims[q] for q in range(N) if q != n])))[1]
        k_sk = int_end if alg.rank_vect is None else alg.rank_vect[n + 1]
        l = int(round(3 * m * math.log(m)))   # qr_lev...:122 / :234
        s = int(round(math.log(m)))           # :123 / :236
        return l, s, min(k_sk, l), (None if seed is None else seed + n)

    # SE-QRCS of the tensor's unfoldings: every random mode in ONE call, in mode order -- the reference's generators draw from one
    # global rand() stream and only these modes consume it, so the stream is the same as in the reference's loop, and the library
    # overlaps the host half of mode n+1 with the device half of mode n
    sketched = {}
    if not krp_mode and not isinstance(alg, QRPivProjected):
        ms = [n for n in range(N) if (n + 1) in lst]
        if ms:
            prm = [sketch_params(n) for n in ms]
            res = eng.seqrcs_modes(ms, [q[0] for q in prm], [q[1] for q in prm], [q[2] for q in prm], injective=injective,
                                   seeds=None if seed is None else [q[3] for q in prm])
            sketched = {n: (r[0], r[1]) for n, r in zip(ms, res)}
    for n in range(N):
        rdims = [dims[m] for m in range(N) if m != n]
        dRis = int(np.prod(rdims))
        int_start, int_end = _proj_range(alg, n, dRis)
        m = dims[n]
        if n in sketched:
            p, dr = sketched[n]
        elif (n + 1) in lst and not isinstance(alg, QRPivProjected):
            l, s, t, sd = sketch_params(n)
            p, dr, _ = eng.seqrcs_krp(n, l, s, t, injective=injective, seed=sd)
        elif krp_mode:
            # exact QRCP of the (cprank x dRis) Khatri-Rao matrix of the preliminary factors (:241-243)
            K = np.ones((1, updated.rank))
            for f in (f for i, f in enumerate(updated.factors) if i != n):
                K = (f[None, :, :] * K[:, None, :]).reshape(-1, updated.rank, order="F")
            p, dr = eng.qrcp_matrix(np.asfortranarray(K.T))
        else:
            p, dr = eng.qrcp_unfolding(n)
        ref_pivs.append(p.copy())
        meff = int(np.sum(np.abs(dr) / np.max(np.abs(dr)) > trunc_tol))   # :28 / :137
        eff.append(meff)
        rest = p[meff:]
        p = np.concatenate([p[:meff], rest[rng.permutation(len(rest))] if shuffle_pivots else rest])
        coords = column_to_multi_coords(p, rdims)
        pivots.append(coords)
        proj = np.asfortranarray(coords[int_start - 1: int_end, :])
        projectors.append(proj)
        eng.set_projector(n, proj)   # gathers + caches target_transform[n] on the device
    extra = dict(ref_projectors=ref_pivs, projects=pivots, projects_tensors=projectors, effective_ranks=eff,
                 normal=True if normal is None else normal, dims=tuple(dims), owns_tensor=owns_tensor)
    if owns_tensor:
        # ALS(ITensor(inds(target)), ...): only the samples are needed from here on (:77, :175).  The reference drops ITS reference to
        # the tensor, not the caller's: a tensor that lives in an Engine the caller passed in (rank-adaptive decompose keeps it
        # resident across rank steps, decompose.jl:51-66) stays where it is.
        eng.drop_tensor()
    return ALS(eng, alg, extra, check)


def update_samples(target, als: "ALS", new_num_end, reshuffle=False, new_num_start=0, rng=None) -> "ALS":
    """algorithms/.../qr_lev_score_sampled.jl:95-149: new sample range without redoing the QR (re-gathers T_s)."""
    from .engine import column_to_multi_coords

    rng = np.random.default_rng(11) if rng is None else rng
    old = als.mttkrp_alg
    assert isinstance(old, _PivotBased)
    alg = type(old).__new__(type(old))
    _PivotBased.__init__(alg, old.Start if new_num_start == 0 else new_num_start, old.End if new_num_end == 0 else new_num_end,
                         old.random_modes, old.rank_vect)
    eng = als.engine
    ai = als.additional_items
    owns = not isinstance(target, Engine)
    if owns:
        eng.set_tensor(target)  # the setup dropped the dense tensor; the reference also re-reads `target` here (:135)
    else:
        assert target is eng, "update_samples: pass the host tensor or the Engine the ALS object was set up on"
    dims = eng.dims
    N = len(dims)
    pivots = [p.copy() for p in ai["projects"]]
    projectors = []
    for pos in range(N):
        rdims = [dims[m] for m in range(N) if m != pos]
        if reshuffle:
            p, meff = ai["ref_projectors"][pos], ai["effective_ranks"][pos]
            rest = p[meff:]
            pivots[pos] = column_to_multi_coords(np.concatenate([p[:meff], rest[rng.permutation(len(rest))]]), rdims)
        int_start, int_end = _proj_range(alg, pos, int(np.prod(rdims)))
        proj = np.asfortranarray(pivots[pos][int_start - 1: int_end, :])
        projectors.append(proj)
        eng.set_projector(pos, proj)
    if owns:
        eng.drop_tensor()
    extra = dict(ai)
    extra.update(projects=pivots, projects_tensors=projectors)
    return ALS(eng, alg, extra, als.check)


# ------------------------------------------------------------------------------------------
# ALS driver
# ------------------------------------------------------------------------------------------
class ALS:  # als_optimizer.jl:5-10, with the device handle as `target`
    def __init__(self, engine: Engine, alg, additional_items: dict, check: ConvergeAlg):
        self.engine = engine
        self.target = engine
        self.mttkrp_alg = alg
        self.additional_items = additional_items
        self.check = check
        self.solve_paths = []

    def fetch_cpd_arrays(self):
        e = self.engine
        return [e.get_factor(n) for n in range(len(e.dims))], e.get_lambda()


_default_engines = {}


def _engine_for(target, device=0) -> Engine:
    """`target` is a host array (uploaded into the per-device default engine, which this module then owns) or an Engine
    that already holds the tensor (owned by the caller: never dropped or overwritten here)."""
    if isinstance(target, Engine):
        return target
    eng = _default_engines.get(device)
    if eng is None:
        eng = _default_engines[device] = Engine(device)
    eng.set_tensor(target)
    return eng


def compute_als(target, cp: CPD, alg=None, check=None, maxiter=None, normal=None, stop_resample=-1, device=0, seed=0,
                shuffle_pivots=True, trunc_tol=0.01, injective=False, rng=None, guess_num_levs=None, prelim_niter=10, **_) -> ALS:
    """als_optimizer.jl:37-69 + standard/tensor.jl:3-14 + randomized/krp_lev_score_sampled.jl:1-40."""
    alg = KRPFreeNormal() if alg is None else alg
    check = NoCheck(100 if maxiter is None else maxiter) if check is None else check
    eng = _engine_for(target, device)
    assert tuple(eng.dims) == tuple(cp.dims), (eng.dims, cp.dims)
    eng.set_cpd(cp.factors, cp.lam)
    extra = {}
    if isinstance(alg, MttkrpAlgorithm):
        eng.set_option("mttkrp_alg", alg.mttkrp_alg)
        eng.compute_grams()  # :part_grammian
    elif isinstance(alg, LevScoreSampled):
        extra.update(normal=False if normal is None else normal, stop_resample=stop_resample, seed=int(seed) * 1000003,
                     projects_tensors=[None] * len(cp), rng=np.random.default_rng(seed) if rng is None else rng)
        for n in range(len(cp)):
            eng.leverage_scores(n)  # :factor_weights
    elif isinstance(alg, _PivotBased):
        return _setup_pivot_based(alg, eng, cp, check, normal, shuffle_pivots, trunc_tol, injective, rng, seed,
                                  guess_num_levs=guess_num_levs, prelim_niter=prelim_niter, owns_tensor=not isinstance(target, Engine))
    else:
        raise TypeError(f"unsupported algorithm {type(alg).__name__}")
    return ALS(eng, alg, extra, check)


def optimize(cp: CPD, als: ALS, verbose=False) -> CPD:
    """optimize.jl:6-35 hook by hook; the state lives on the device between hooks."""
    it = als.check.iter
    alg = als.mttkrp_alg
    N = len(cp)
    eng = als.engine
    if isinstance(alg, _PivotBased):
        eng.set_cpd(cp.factors, cp.lam)  # an ALS object can be re-used with another starting CPD (update_samples)
    fused = isinstance(alg, MttkrpAlgorithm) and isinstance(als.check, (NoCheck, FitCheck)) and not als.additional_items.get("per_hook")
    # leverage-score sampling that redraws every iteration: the whole sweep (draws, gathers, sampled solves, leverage refresh) stays on
    # the device; same seeds and kernels as the hook-by-hook path below (bitwise equal), no host round trip per mode
    fused_sampled = (type(alg) is LevScoreSampled and als.additional_items.get("stop_resample", -1) < 0
                     and not als.additional_items.get("per_hook") and not getattr(eng, "sharded", False))
    while it < als.check.max_counter:
        if fused_sampled:
            ai = als.additional_items
            eng.sampled_sweep_async(1, [alg.nsamples(f) for f in range(N)], ai["seed"], CHOLESKY_EPSILON, normal=ai["normal"])
            ai["seed"] += N
            done = alg.check_converge(als, verbose)
        elif fused:
            # one device-resident sweep (itcpd_sweep == the five hooks for every mode + fit scalars)
            inner, norm2 = eng.sweep(1, CHOLESKY_EPSILON)
            if isinstance(als.check, FitCheck):
                done = als.check.update(float(inner[0]), float(norm2[0]), eng.rank, verbose)
            else:
                done = als.check.check_converge(als, None, verbose=verbose)
        else:
            for fact in range(N):
                alg.compute_krp(als, fact)
                alg.matricize_tensor(als, fact)
                alg.solve_ls_problem(als, fact)
                if isinstance(alg, MttkrpAlgorithm):
                    eng.normalize(fact)  # row_norm, optimize.jl:25
                alg.post_solve(als, fact)
            done = alg.check_converge(als, verbose)
        if done:
            break
        it += 1
    f, l = als.fetch_cpd_arrays()
    return CPD(f, l)


def als_optimize(target, cp: CPD, alg=None, check=None, maxiter=None, verbose=False, **kwargs) -> CPD:
    als = compute_als(target, cp, alg=alg, check=check, maxiter=maxiter, **kwargs)
    return optimize(cp, als, verbose=verbose)


def decompose(A, rank, *args, solver=None, rng=None, alg=None, check=None, maxiter=None, verbose=False, **kwargs) -> CPD:
    """decompose.jl:1-30 (fixed rank) and :32-69 (rank adaptive: decompose(A, epsilon, max_rank; ...))."""
    if args:
        return _decompose_adaptive(A, rank, args[0], rng=rng, alg=alg, check=check, maxiter=maxiter, verbose=verbose, **kwargs)
    if solver is not None:
        if not isinstance(solver, CPDOptimizer):
            raise TypeError("solver must be a CPDOptimizer or nothing")  # test/cp_als.jl:15
        raise RuntimeError("OptimizerError")  # decompose.jl:27-29
    dims = A.dims if isinstance(A, Engine) else A.shape
    cp = random_CPD(dims, rank, rng)
    return als_optimize(A, cp, alg=alg, check=check, maxiter=maxiter, verbose=verbose, **kwargs)


class CPDOptimizer:  # cpd_optimizers.jl:1
    pass


def increase_cpd_rank(cp: CPD, new_rank: int, rng=None) -> CPD:  # decompose.jl:71-82
    rng = np.random.default_rng(3) if rng is None else rng
    assert new_rank >= cp.rank
    newf, lam = random_factors(cp.dims, new_rank, rng)
    for old, new in zip(cp.factors, newf):
        new[:, : cp.rank] = old
    return CPD(newf, lam)


def _decompose_adaptive(A, epsilon, max_rank, rng=None, alg=None, check=None, maxiter=None, verbose=False,
                        start_rank=1, rank_step=1, device=0, **kw) -> CPD:
    eng = _engine_for(A, device)  # the tensor is uploaded once and stays resident across rank steps
    if verbose:
        print(f"Starting with rank: {start_rank}")
    current = start_rank
    cp = random_CPD(eng.dims, start_rank, rng)
    check = FitCheck(1e-3, 100, eng.tensor_norm()) if check is None else check
    while True:
        cp = als_optimize(eng, cp, alg=alg, check=check, maxiter=maxiter, verbose=verbose, **kw)
        check.iter = 0
        if 1.0 - CPDFit(check) < epsilon:
            return cp
        current += rank_step
        if verbose:
            print(f"\nIncreasing rank to: {current}")
        if current > max_rank:
            print(f"Optimization Failed to converge within rank {max_rank}")
            return cp
        cp = increase_cpd_rank(cp, current, rng)


_recon_engines = {}


def reconstruct(cp: CPD, device=0) -> np.ndarray:
    """reconstruct.jl:2-9 on the device (no P x R intermediate).  Runs on a handle of its own that only knows the SHAPE
    (itcpd_set_shape allocates no tensor), so a tensor resident in the decomposition engine is never disturbed."""
    eng = _recon_engines.get(device) or _recon_engines.setdefault(device, Engine(device))
    eng.set_shape(cp.dims)
    eng.set_cpd(cp.factors, cp.lam)
    return eng.reconstruct()
