"""Thin object wrapper over the C-ABI handle.  Arrays cross the boundary as numpy float64 /
int64 buffers in column-major (Fortran) order -- the layout Julia hands to `ccall`."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check


def _f64(a) -> np.ndarray:
    return np.asfortranarray(a, dtype=np.float64)


def _addr(a) -> int:
    """Raw host address of a numpy array or of anything exposing data_ptr() (pinned torch tensors)."""
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    return int(a.ctypes.data)


class Engine:
    """One handle = one GPU = one driver thread (include/itcpd_b200.h)."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        h = C.c_void_p()
        check(self._L.itcpd_create(C.byref(h), int(device)))
        self._h = h
        self.device = int(device)
        self.dims: tuple = ()
        self.rank = 0

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.itcpd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- info -------------------------------------------------------------------------------
    def device_info(self):
        sm, ma, mi, hb = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        check(self._L.itcpd_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(hb)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "hbm_bytes": hb.value}

    def synchronize(self):
        check(self._L.itcpd_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.itcpd_launch_count(self._h))

    def set_option(self, name: str, value: int):
        check(self._L.itcpd_set_option(self._h, name.encode(), int(value)))

    # -- tensor -----------------------------------------------------------------------------
    def _dims_arg(self, dims):
        arr = (C.c_int64 * len(dims))(*[int(d) for d in dims])
        return arr

    def set_tensor(self, T, dims: Optional[Sequence[int]] = None):
        """Upload a host tensor.  `T` is a numpy array (any layout; converted to column-major) or a
        pinned host buffer exposing data_ptr() together with `dims`."""
        if dims is None:
            T = _f64(T)
            dims = T.shape
            self._keep = T
        self.dims = tuple(int(d) for d in dims)
        check(self._L.itcpd_set_tensor(self._h, len(self.dims), self._dims_arg(self.dims), _addr(T)))

    def set_shape(self, dims):
        """extents only: no tensor storage is allocated or touched (reconstruct / CPD-only use of a handle)"""
        self.dims = tuple(int(d) for d in dims)
        check(self._L.itcpd_set_shape(self._h, len(self.dims), self._dims_arg(self.dims)))

    def generate_tensor(self, dims, seed: int = 0, elem_offset: int = 0):
        self.dims = tuple(int(d) for d in dims)
        check(self._L.itcpd_generate_tensor(self._h, len(self.dims), self._dims_arg(self.dims), int(seed), int(elem_offset)))

    def generate_lowrank_tensor(self, dims, rank: int, seed: int = 0, noise: float = 0.0):
        self.dims = tuple(int(d) for d in dims)
        self.rank = int(rank)
        check(self._L.itcpd_generate_lowrank_tensor(self._h, len(self.dims), self._dims_arg(self.dims), int(rank), int(seed), float(noise)))

    def get_tensor(self, out=None) -> np.ndarray:
        if out is None:
            out = np.empty(self.dims, dtype=np.float64, order="F")
        check(self._L.itcpd_get_tensor(self._h, _addr(out)))
        return out

    def tensor_norm(self) -> float:
        v = C.c_double()
        check(self._L.itcpd_tensor_norm(self._h, C.byref(v)))
        return v.value

    # -- CPD state --------------------------------------------------------------------------
    def set_rank(self, rank: int):
        self.rank = int(rank)
        check(self._L.itcpd_set_rank(self._h, self.rank))

    def set_factor(self, mode: int, A):
        A = _f64(A)
        assert A.shape == (self.dims[mode], self.rank), (A.shape, self.dims[mode], self.rank)
        check(self._L.itcpd_set_factor(self._h, mode, _addr(A)))

    def get_factor(self, mode: int) -> np.ndarray:
        out = np.empty((self.dims[mode], self.rank), dtype=np.float64, order="F")
        check(self._L.itcpd_get_factor(self._h, mode, _addr(out)))
        return out

    def set_lambda(self, lam):
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        check(self._L.itcpd_set_lambda(self._h, _addr(lam)))

    def get_lambda(self) -> np.ndarray:
        out = np.empty(self.rank, dtype=np.float64)
        check(self._L.itcpd_get_lambda(self._h, _addr(out)))
        return out

    def get_gram(self, mode: int) -> np.ndarray:
        out = np.empty((self.rank, self.rank), dtype=np.float64, order="F")
        check(self._L.itcpd_get_gram(self._h, mode, _addr(out)))
        return out

    def set_cpd(self, factors, lam=None):
        self.set_rank(factors[0].shape[1])
        for n, f in enumerate(factors):
            self.set_factor(n, f)
        if lam is not None:
            self.set_lambda(lam)

    def random_cpd(self, seed: int = 3):
        check(self._L.itcpd_random_cpd(self._h, int(seed)))

    # -- hooks ------------------------------------------------------------------------------
    def compute_grams(self):
        check(self._L.itcpd_compute_grams(self._h))

    def gram_hadamard(self, mode: int, fetch=True):
        out = np.empty((self.rank, self.rank), dtype=np.float64, order="F") if fetch else None
        check(self._L.itcpd_gram_hadamard(self._h, mode, _addr(out) if fetch else None))
        return out

    def mttkrp(self, mode: int, fetch=True):
        out = np.empty((self.dims[mode], self.rank), dtype=np.float64, order="F") if fetch else None
        check(self._L.itcpd_mttkrp(self._h, mode, _addr(out) if fetch else None))
        return out

    def solve(self, mode: int, chol_tol: float = 1e-6):
        path, rk = C.c_int(), C.c_int()
        check(self._L.itcpd_solve(self._h, mode, float(chol_tol), C.byref(path), C.byref(rk)))
        return path.value, rk.value

    def last_solve_status(self, slot: int = 0):
        """(path, rank) of the most recent R x R solve (slot 0; after whole sweeps slot = mode index)"""
        path, rank = C.c_int(0), C.c_int(0)
        check(self._L.itcpd_last_solve_status(self._h, int(slot), C.byref(path), C.byref(rank)))
        return path.value, rank.value

    def normalize(self, mode: int):
        check(self._L.itcpd_normalize(self._h, mode))

    def post_solve(self, mode: int):
        check(self._L.itcpd_post_solve(self._h, mode))

    def fit_terms(self):
        a, b = C.c_double(), C.c_double()
        check(self._L.itcpd_fit_terms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def cpd_snapshot(self):
        check(self._L.itcpd_cpd_snapshot(self._h))

    def cpd_diff_terms(self):
        a, b = C.c_double(), C.c_double()
        check(self._L.itcpd_cpd_diff_terms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def sweep(self, nsweeps: int = 1, chol_tol: float = 1e-6):
        inner = np.empty(nsweeps)
        norm2 = np.empty(nsweeps)
        check(self._L.itcpd_sweep(self._h, int(nsweeps), float(chol_tol), _addr(inner), _addr(norm2)))
        return inner, norm2

    def sweep_async(self, nsweeps: int = 1, chol_tol: float = 1e-6):
        check(self._L.itcpd_sweep_async(self._h, int(nsweeps), float(chol_tol)))

    def sweep_results(self, nsweeps: int):
        inner = np.empty(nsweeps)
        norm2 = np.empty(nsweeps)
        fb = C.c_int()
        check(self._L.itcpd_sweep_results(self._h, int(nsweeps), _addr(inner), _addr(norm2), C.byref(fb)))
        return inner, norm2, fb.value

    def als_from_host(self, T, factors, nsweeps: int, chol_tol: float = 1e-6, dims=None):
        """als_optimize from HOST buffers in one call (upload, sweeps, download)."""
        if dims is None:
            T = _f64(T)
            dims = T.shape
        self.dims = tuple(int(d) for d in dims)
        N = len(self.dims)
        fin = [_f64(f) for f in factors]
        R = fin[0].shape[1]
        self.rank = R
        fout = [np.empty((self.dims[n], R), order="F") for n in range(N)]
        lam = np.empty(R)
        inner, norm2 = np.empty(nsweeps), np.empty(nsweeps)
        pin = (C.c_void_p * N)(*[_addr(f) for f in fin])
        pout = (C.c_void_p * N)(*[_addr(f) for f in fout])
        check(self._L.itcpd_als_from_host(self._h, N, self._dims_arg(self.dims), _addr(T), R, pin, int(nsweeps), float(chol_tol),
                                          pout, _addr(lam), _addr(inner), _addr(norm2)))
        return fout, lam, inner, norm2

    # -- reconstruct ------------------------------------------------------------------------
    def reconstruct(self) -> np.ndarray:
        out = np.empty(self.dims, dtype=np.float64, order="F")
        check(self._L.itcpd_reconstruct(self._h, _addr(out)))
        return out

    def residual_norm(self) -> float:
        v = C.c_double()
        check(self._L.itcpd_residual_norm(self._h, C.byref(v)))
        return v.value

    # -- sampled path -----------------------------------------------------------------------
    def leverage_scores(self, mode: int) -> np.ndarray:
        out = np.empty(self.dims[mode])
        check(self._L.itcpd_leverage_scores(self._h, mode, _addr(out)))
        return out

    def sample_factor_matrices(self, skip_mode: int, nsamp: int, seed: int) -> np.ndarray:
        out = np.empty((nsamp, len(self.dims) - 1), dtype=np.int64, order="F")
        check(self._L.itcpd_sample_factor_matrices(self._h, skip_mode, int(nsamp), int(seed), _addr(out)))
        return out

    @staticmethod
    def _piv(p) -> np.ndarray:
        return np.asfortranarray(p, dtype=np.int64)

    def pivot_hadamard(self, mode: int, pivots) -> np.ndarray:
        p = self._piv(pivots)
        out = np.empty((p.shape[0], self.rank), order="F")
        check(self._L.itcpd_pivot_hadamard(self._h, mode, p.shape[0], _addr(p), _addr(out)))
        return out

    def gather_fibers(self, mode: int, pivots) -> np.ndarray:
        p = self._piv(pivots)
        out = np.empty((self.dims[mode], p.shape[0]), order="F")
        check(self._L.itcpd_gather_fibers(self._h, mode, p.shape[0], _addr(p), _addr(out)))
        return out

    def sketch_unfolding(self, mode: int, l: int, s: int, rows0, vals) -> np.ndarray:
        rows0 = np.ascontiguousarray(rows0, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        out = np.empty((self.dims[mode], l), order="F")
        check(self._L.itcpd_sketch_unfolding(self._h, mode, int(l), int(s), _addr(rows0), _addr(vals), _addr(out)))
        return out

    def sketch_unfolding_omega(self, mode: int, omega) -> np.ndarray:
        """sparse-matrix variant (pivot_mapping.jl:90-104): `omega` is a scipy.sparse matrix (l x ncols); handed over in Julia's
        SparseMatrixCSC convention (1-based colptr / rowval)"""
        om = omega.tocsc()
        om.sort_indices()
        colptr = np.ascontiguousarray(om.indptr, dtype=np.int64) + 1
        rowval = np.ascontiguousarray(om.indices, dtype=np.int64) + 1
        nzval = np.ascontiguousarray(om.data, dtype=np.float64)
        l, ncols = om.shape
        out = np.empty((self.dims[mode], l), order="F")
        check(self._L.itcpd_sketch_unfolding_csc(self._h, mode, int(l), int(ncols), _addr(colptr), _addr(rowval), _addr(nzval), _addr(out)))
        return out

    def sampled_update(self, mode: int, pivots, chol_tol: float = 1e-6, normal: bool = True):
        p = self._piv(pivots)
        check(self._L.itcpd_sampled_update(self._h, mode, p.shape[0], _addr(p), float(chol_tol), int(bool(normal))))

    # -- pivot-projected solvers ------------------------------------------------------------
    def sampled_sweep_async(self, nsweeps: int, nsamps, draw_counter: int, chol_tol: float = 1e-6, normal: bool = True):
        """device-resident sweeps of the leverage-score sampled solver; the k-th draw is seeded with draw_counter + k"""
        ns = (C.c_int64 * len(self.dims))(*[int(n) for n in nsamps])
        check(self._L.itcpd_sampled_sweep_async(self._h, int(nsweeps), ns, int(draw_counter), float(chol_tol), int(bool(normal))))

    def qrcp_unfolding(self, mode: int):
        """qr(T_(mode), ColumnNorm()): (p 1-based int64 over all unfolding columns, diag(R))."""
        m = self.dims[mode]
        n = int(np.prod(self.dims)) // m
        piv = np.empty(n, dtype=np.int64)
        rd = np.empty(min(m, n))
        check(self._L.itcpd_qrcp_unfolding(self._h, mode, _addr(piv), _addr(rd)))
        return piv, rd

    def qrcp_matrix(self, A, steps: Optional[int] = None):
        A = _f64(A)
        m, n = A.shape
        steps = min(m, n) if steps is None else int(steps)
        piv = np.empty(n, dtype=np.int64)
        rd = np.empty(min(m, n, steps))
        check(self._L.itcpd_qrcp_matrix(self._h, m, n, _addr(A), steps, _addr(piv), _addr(rd)))
        return piv, rd

    def seqrcs(self, mode: int, l: int, s: int, t: int, injective: bool = False, seed: Optional[int] = None, use_omega: bool = False):
        """SEQRCS (SEQRCS.jl:139-182, compute_r=false): (p 1-based, diag(R) of the candidate QR, #candidates).
        use_omega=True: the candidate order of the sparse-matrix variant (SEQRCS.jl:89-134)."""
        self.set_option("seqrcs_use_omega", int(bool(use_omega)))
        m = self.dims[mode]
        n = int(np.prod(self.dims)) // m
        piv = np.empty(n, dtype=np.int64)
        rd = np.empty(m)
        nr, nc = C.c_int64(), C.c_int64()
        if seed is not None:
            C.CDLL(None).srand(C.c_uint(seed))
        check(self._L.itcpd_seqrcs(self._h, mode, int(l), int(s), int(t), int(bool(injective)), _addr(piv), _addr(rd),
                                   C.byref(nr), C.byref(nc)))
        return piv, rd[: nr.value].copy(), nc.value

    def seqrcs_modes(self, modes, ls, ss, ts, injective: bool = False, seeds=None, use_omega: bool = False):
        """SE-QRCS of several modes in one call (the set-up loop of optimizers/.../qr_lev_score_sampled.jl:80-176): same results and same
        rand() stream as seqrcs() mode by mode, with the host half of the next mode overlapped with the device half of this one.
        Returns [(p, diag R, #candidates)] in the order of `modes`."""
        self.set_option("seqrcs_use_omega", int(bool(use_omega)))
        nm = len(modes)
        total = int(np.prod(self.dims))
        pivs = [np.empty(total // self.dims[m], dtype=np.int64) for m in modes]
        rds = [np.empty(self.dims[m]) for m in modes]
        as_i32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
        modes_a, l_a, s_a, t_a = as_i32(modes), as_i32(ls), as_i32(ss), as_i32(ts)
        seeds_a = None if seeds is None else np.ascontiguousarray([-1 if x is None else int(x) for x in seeds], dtype=np.int64)
        piv_ptrs = (C.c_void_p * nm)(*[_addr(p) for p in pivs])
        rd_ptrs = (C.c_void_p * nm)(*[_addr(r) for r in rds])
        nr, nc = np.zeros(nm, dtype=np.int64), np.zeros(nm, dtype=np.int64)
        check(self._L.itcpd_seqrcs_modes(self._h, nm, _addr(modes_a), _addr(l_a), _addr(s_a), _addr(t_a), int(bool(injective)),
                                         None if seeds_a is None else _addr(seeds_a), C.cast(piv_ptrs, C.c_void_p), C.cast(rd_ptrs, C.c_void_p),
                                         _addr(nr), _addr(nc)))
        return [(pivs[i], rds[i][: int(nr[i])].copy(), int(nc[i])) for i in range(nm)]

    def seqrcs_krp(self, mode: int, l: int, s: int, t: int, injective: bool = False, seed: Optional[int] = None):
        """KRP-structured SE-QRCS of the handle's current factors != mode (SEQRCS.jl:184-241, compute_r=false)."""
        n = int(np.prod([d for m, d in enumerate(self.dims) if m != mode]))
        piv = np.empty(n, dtype=np.int64)
        rd = np.empty(max(self.rank, 1))
        nr, nc = C.c_int64(), C.c_int64()
        if seed is not None:
            C.CDLL(None).srand(C.c_uint(seed))
        check(self._L.itcpd_seqrcs_krp(self._h, mode, int(l), int(s), int(t), int(bool(injective)), _addr(piv), _addr(rd),
                                       C.byref(nr), C.byref(nc)))
        return piv, rd[: nr.value].copy(), nc.value

    def set_projector(self, mode: int, pivots):
        p = self._piv(pivots)
        check(self._L.itcpd_set_projector(self._h, mode, p.shape[0], _addr(p)))

    def projected_update(self, mode: int, chol_tol: float = 1e-6, normal: bool = True):
        check(self._L.itcpd_projected_update(self._h, mode, float(chol_tol), int(bool(normal))))

    def drop_tensor(self):
        check(self._L.itcpd_drop_tensor(self._h))

    # -- multi-GPU --------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_lib.load().itcpd_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        check(self._L.itcpd_comm_init(self._h, int(nranks), int(rank), buf))

    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(self._L.itcpd_peer_export(self._h, buf))
        return buf.raw

    def peer_import(self, nranks: int, rank: int, handles: bytes):
        assert len(handles) == 64 * nranks
        buf = C.create_string_buffer(handles, len(handles))
        check(self._L.itcpd_peer_import(self._h, int(nranks), int(rank), buf))

    def peer_disable(self):
        check(self._L.itcpd_peer_disable(self._h))

    def allgather_factor(self, mode: int, rows_total: int) -> np.ndarray:
        out = np.empty((rows_total, self.rank), order="F")
        check(self._L.itcpd_allgather_factor(self._h, mode, int(rows_total), _addr(out)))
        return out

    # -- measurement ------------------------------------------------------------------------
    def gemm_timing(self, reset=True):
        ms, n = C.c_double(), C.c_int64()
        check(self._L.itcpd_gemm_timing(self._h, int(reset), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    PHASES = ("between_modes", "mttkrp", "peer_signal", "solve", "normalize", "gram", "fit")

    def phase_timing(self, reset=True) -> dict:
        """milliseconds per sweep phase since the last reset (option time_phases=1)"""
        ms = np.zeros(len(self.PHASES))
        n = C.c_int64()
        check(self._L.itcpd_phase_timing(self._h, int(reset), len(self.PHASES), _addr(ms), C.byref(n)))
        return dict(zip(self.PHASES, ms.tolist()), marks=n.value)

    def probe_dmma_peak(self) -> float:
        v = C.c_double()
        check(self._L.itcpd_probe_dmma_peak(self._h, C.byref(v)))
        return v.value

    def probe_dfma_peak(self) -> float:
        v = C.c_double()
        check(self._L.itcpd_probe_dfma_peak(self._h, C.byref(v)))
        return v.value


    def event_record(self, slot: int):
        check(self._L.itcpd_event_record(self._h, int(slot)))

    def event_elapsed_ms(self, s0: int, s1: int) -> float:
        v = C.c_double()
        check(self._L.itcpd_event_elapsed_ms(self._h, int(s0), int(s1), C.byref(v)))
        return v.value

    def flush_l2(self, nbytes: int = 256 << 20):
        check(self._L.itcpd_flush_l2(self._h, int(nbytes)))


class PinnedBuffer:
    """Page-locked host memory (cudaHostAlloc) viewed as a numpy float64 array."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)
        n = int(np.prod(self.shape))
        p = C.c_void_p()
        check(_lib.load().itcpd_host_alloc(n * 8, C.byref(p)))
        self._p = p
        buf = (C.c_double * n).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=np.float64).reshape(self.shape, order="F")

    def data_ptr(self) -> int:
        return int(self._p.value)

    def free(self):
        if self._p:
            self.array = None
            _lib.load().itcpd_host_free(self._p)
            self._p = None


# host-side integer maps and generators (no handle needed)
def column_to_multi_coords(cols, dims) -> np.ndarray:
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    d = (C.c_int64 * len(dims))(*[int(x) for x in dims])
    out = np.empty((cols.shape[0], len(dims)), dtype=np.int64, order="F")
    check(_lib.load().itcpd_column_to_multi_coords(cols.shape[0], _addr(cols), len(dims), d, _addr(out)))
    return out


def multi_coords_to_column(dims, coords) -> np.ndarray:
    coords = np.asfortranarray(coords, dtype=np.int64)
    d = (C.c_int64 * len(dims))(*[int(x) for x in dims])
    out = np.empty(coords.shape[0], dtype=np.int64)
    check(_lib.load().itcpd_multi_coords_to_column(coords.shape[0], _addr(coords), len(dims), d, _addr(out)))
    return out


def sparse_sign_matrix(l: int, n: int, s: int, injective: bool = False, seed: Optional[int] = None):
    """SEQRCS.jl:29-60 through the library's generators.  Returns (rows 0-based int32, vals, colstarts)."""
    L = _lib.load()
    s_eff = min(s, l)
    vals = np.full(n * s_eff, np.nan)
    rows = np.zeros(n * s_eff, dtype=np.int32)
    colstarts = np.zeros(n + 1, dtype=np.int32)
    if seed is not None:
        C.CDLL(None).srand(C.c_uint(seed))
    (L.itcpd_sparsestack if injective else L.itcpd_sparse_sign)(l, n, s, _addr(vals), _addr(rows), _addr(colstarts))
    return rows, vals, colstarts
