"""CPU emulation of the SE-QRCS set-up kernels (the kernel text is taken verbatim from csrc/qrcp_wide.cu and csrc/sampled.cu and
compiled with g++ behind tests/simt_emu.h: every CUDA thread is an OS thread, __syncthreads / warp shuffles are real rendezvous):
  * unfold_tiled_kernel  -- the explicit unfolding the sketch reads (32 x 32 tiles through shared memory, partial tiles, a padded
                            leading dimension, orders 2 to 4), against numpy's unfolding, bitwise;
  * sketch_offsets_kernel / sketch_offsets_unfolded_kernel / sketch_kernel -- the sparse-sign sketch from the tensor in place and
                            from the explicit unfolding: bitwise equal to each other, 1e-13 against the oracle;
  * qrw_norm_init / qrw_pivot_house / qrw_apply_kernel / qrw_apply_reg_kernel<NR> -- the wide column-pivoted QR with the host loop
                            of k_qrcp_wide restated in the driver: pivots and |diag R| against LAPACK dgeqp3 (scipy) for the two-pass
                            kernel and for every register-resident instantiation the row count selects.
The GPU suite runs the same comparisons on hardware (tests/test_gpu_sampled.py); this file keeps them in the CPU suite."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import cpals, sampled

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "itensorcpd.jl_b200", "csrc")

PRE = r"""
#include "simt_emu.h"
#include <algorithm>
#include <cstdlib>
#define ITCPD_MAX_ORDER 8
#define ITCPD_OK 0
using std::min;
using std::max;
static double g_dyn_smem[1 << 15];   // the one `extern __shared__` array of qrw_apply_kernel
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
"""

DRIVER = r"""
template <typename F> static void launch1d(int64_t grid, int threads, F f) {
    for (int64_t b = 0; b < grid; ++b) emu_launch(threads, 0, f, (unsigned)b, 0, (unsigned)grid, 1);
}

extern "C" void run_unfold(int n, const int64_t *dims, int64_t ld0, const double *T, int mode, double *out) {
    UDims d; d.n = n;
    int64_t nelem = 1;
    for (int i = 0; i < n; ++i) { d.ext[i] = i == 0 ? ld0 : dims[i]; d.dim[i] = dims[i]; nelem *= dims[i]; }
    const int64_t I = dims[mode], ncols = nelem / I;
    if (mode == 0) { launch1d(3, 64, [&] { unfold_kernel(T, d, mode, ncols, out); }); return; }
    int64_t A = 1;
    for (int q = 0; q < mode; ++q) A *= dims[q];
    const int64_t B = ncols / A, tiles_a = ceil_div(A, 32), tiles_i = ceil_div(I, 32);
    launch1d(tiles_a * tiles_i * B, 256, [&] { unfold_tiled_kernel(T, d, mode, A, I, tiles_a, tiles_i, out); });
}

// k_sketch_csr, restated: decode the column numbers once (in place), then gather
extern "C" void run_sketch(int n, const int64_t *dims, int64_t ld0, const double *T, int mode, int l, int64_t nnz, const int64_t *row_ptr,
                           int64_t *col, const double *val, double *out, const double *unfolded) {
    SDims d; d.n = n;
    for (int i = 0; i < n; ++i) { d.ext[i] = i == 0 ? ld0 : dims[i]; d.dim[i] = dims[i]; }
    int64_t stride_mode = 1;
    for (int m = 0; m < mode; ++m) stride_mode *= d.ext[m];
    if (unfolded) launch1d(2, 256, [&] { sketch_offsets_unfolded_kernel(dims[mode], nnz, col); });
    else launch1d(2, 256, [&] { sketch_offsets_kernel(d, mode, nnz, col); });
    const unsigned gy = (unsigned)std::min<int64_t>(ceil_div(dims[mode], 128), 8);
    for (int j = 0; j < l; ++j)
        for (unsigned y = 0; y < gy; ++y)
            emu_launch(128, 0, [&] { sketch_kernel(unfolded ? unfolded : T, dims[mode], unfolded ? 1 : stride_mode, row_ptr, col, val, out); },
                       (unsigned)j, y, (unsigned)l, gy);
}

// the host loop of k_qrcp_wide (qrcp_wide.cu), launch for launch; two_pass = the older apply kernel for every step
extern "C" int run_qrcp(double *A, int64_t m, int64_t n, int64_t steps, int64_t *jpvt, double *rdiag, int two_pass, int *nr_used) {
    const int64_t kmax = std::min<int64_t>(std::min(m, n), steps);
    const int64_t nb0 = ceil_div(n, QW_WARPS);
    std::vector<double> vn(n), bmax_v(nb0 + 1), vbuf(m), scal(2);
    std::vector<int64_t> bmax_i(nb0 + 1);
    launch1d(ceil_div(n, 256), 256, [&] { iota_kernel(jpvt, n); });
    launch1d(nb0, 256, [&] { qrw_norm_init_kernel(A, m, n, vn.data(), bmax_v.data(), bmax_i.data()); });
    int64_t nblocks = nb0;
    for (int64_t j = 0; j < kmax; ++j) {
        const int64_t nbl = nblocks;
        launch1d(1, 1024, [&] { qrw_pivot_house_kernel(A, m, n, j, vn.data(), bmax_v.data(), bmax_i.data(), nbl, jpvt, vbuf.data(), scal.data(), rdiag); });
        const int64_t trailing = n - j - 1;
        nblocks = ceil_div(std::max<int64_t>(trailing, 0), QW_WARPS);
        if (trailing <= 0) continue;
        const int64_t span = m - (j & ~(int64_t)3);
        double *a = A; double *v = vn.data(), *bv = bmax_v.data(); int64_t *bi = bmax_i.data();
        const double *vb = vbuf.data(), *sc = scal.data();
        int nr = 0;
        if (two_pass || span > 2048) launch1d(nblocks, 256, [&] { qrw_apply_kernel(a, m, n, j, vb, sc, v, bv, bi); });
        else if (span > 1024) { nr = 64; launch1d(nblocks, 256, [&] { qrw_apply_reg_kernel<64, 1>(a, m, n, j, vb, sc, v, bv, bi); }); }
        else if (span > 512) { nr = 32; launch1d(nblocks, 256, [&] { qrw_apply_reg_kernel<32, 2>(a, m, n, j, vb, sc, v, bv, bi); }); }
        else if (span > 256) { nr = 16; launch1d(nblocks, 256, [&] { qrw_apply_reg_kernel<16, 2>(a, m, n, j, vb, sc, v, bv, bi); }); }
        else if (span > 128) { nr = 8; launch1d(nblocks, 256, [&] { qrw_apply_reg_kernel<8, 2>(a, m, n, j, vb, sc, v, bv, bi); }); }
        else if (span > 64) { nr = 4; launch1d(nblocks, 256, [&] { qrw_apply_reg_kernel<4, 2>(a, m, n, j, vb, sc, v, bv, bi); }); }
        else { nr = 2; launch1d(nblocks, 256, [&] { qrw_apply_reg_kernel<2, 2>(a, m, n, j, vb, sc, v, bv, bi); }); }
        if (nr_used) nr_used[j] = nr;
    }
    return 0;
}
"""


def _between(text, a, b):
    i = text.index(a)
    return text[i:text.index(b, i)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    qw = open(os.path.join(CSRC, "qrcp_wide.cu")).read()
    sm = open(os.path.join(CSRC, "sampled.cu")).read()
    kernels = _between(qw, "constexpr int QW_WARPS", "template <int NR, int MINB = 2>\nstatic void launch_apply_reg")
    kernels += _between(qw, "__global__ void iota_kernel", "// In-place QRCP of A")
    kernels += _between(qw, "struct UDims", "int k_unfold(")
    kernels += _between(sm, "struct SDims {", "static SDims sdims(")
    kernels += _between(sm, "__global__ void __launch_bounds__(256) sketch_offsets_kernel", "// col -> col * I")
    kernels += _between(sm, "__global__ void __launch_bounds__(256) sketch_offsets_unfolded_kernel", "// unfolded_dev: nullptr")
    # the dynamic shared array of the two-pass kernel becomes a static one; the launch helper's <<< >>> never reaches the compiler
    kernels, nsub = re.subn(r"extern __shared__ double sh_v\[\];[^\n]*", "double *sh_v = g_dyn_smem;", kernels)
    assert nsub == 1 and "<<<" not in kernels
    # the driver above restates k_qrcp_wide's dispatch: fail loudly when the product's changes
    dispatch = _between(qw, "            const int64_t span = m - (j & ~(int64_t)3);", "            c->launches++;")
    for frag in ("span > 2048", "launch_apply_reg<64, 1>", "span > 512) launch_apply_reg<32, 2>", "span > 256) launch_apply_reg<16>",
                 "span > 128) launch_apply_reg<8>", "span > 64) launch_apply_reg<4>", "else launch_apply_reg<2>"):
        assert frag in dispatch, frag
    td = tmp_path_factory.mktemp("setup_emu")
    cpp, so = str(td / "emu.cpp"), str(td / "emu.so")
    open(cpp, "w").write(PRE + kernels + DRIVER)
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-I", os.path.join(ROOT, "tests"), "-o", so, cpp],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _padded(T):
    """device storage: the leading dimension rounded up to even (common.cuh: ld0)"""
    ld0 = T.shape[0] + (T.shape[0] & 1)
    S = np.full((ld0,) + T.shape[1:], np.nan, order="F")
    S[: T.shape[0]] = T
    return S, ld0


@pytest.mark.parametrize("dims", [(5, 7, 6), (34, 33, 3), (3, 70), (4, 3, 5, 6)])
def test_tiled_unfolding_is_the_unfolding(emu, dims):
    rng = np.random.default_rng(len(dims))
    T = np.asfortranarray(rng.standard_normal(dims))
    S, ld0 = _padded(T)
    d = np.array(dims, dtype=np.int64)
    for mode in range(len(dims)):
        out = np.full((dims[mode], T.size // dims[mode]), np.nan, order="F")
        emu.run_unfold(len(dims), _p(d), C.c_int64(ld0), _p(S), mode, _p(out))
        assert np.array_equal(out, cpals.unfold(T, mode)), mode


@pytest.mark.parametrize("dims", [(6, 5, 4), (3, 9, 2, 4)])
def test_sketch_in_place_and_from_unfolding(emu, dims):
    rng = np.random.default_rng(7)
    T = np.asfortranarray(rng.standard_normal(dims))
    S, ld0 = _padded(T)
    d = np.array(dims, dtype=np.int64)
    for mode in range(len(dims)):
        I, n = dims[mode], T.size // dims[mode]
        l, s = 2 * I + 1, 2
        vals, rows0, _ = sampled.sparse_sign_call(l, n, s, False, "port", seed=5 + mode)
        order = np.argsort(rows0, kind="stable")                      # CSR by sketch row, entries in increasing non-zero order
        row_ptr = np.concatenate([[0], np.cumsum(np.bincount(rows0, minlength=l))]).astype(np.int64)
        col = (order // s).astype(np.int64)
        val = np.ascontiguousarray(vals[order])
        want = sampled.sketched_matricization(T, mode, l, rows0 + 1, vals, s)
        a = np.full((I, l), np.nan, order="F")
        emu.run_sketch(len(dims), _p(d), C.c_int64(ld0), _p(S), mode, l, C.c_int64(len(col)), _p(row_ptr), _p(col.copy()), _p(val), _p(a), None)
        U = np.asfortranarray(cpals.unfold(T, mode))
        b = np.full((I, l), np.nan, order="F")
        emu.run_sketch(len(dims), _p(d), C.c_int64(ld0), _p(S), mode, l, C.c_int64(len(col)), _p(row_ptr), _p(col.copy()), _p(val), _p(b), _p(U))
        assert np.array_equal(a, b), mode
        assert np.linalg.norm(a - want) <= 1e-13 * np.linalg.norm(want), mode


@pytest.mark.parametrize("m,n,two_pass", [(9, 20, 0), (9, 20, 1), (70, 24, 0), (131, 17, 0), (12, 5, 0)])
def test_wide_qrcp_matches_lapack(emu, m, n, two_pass):
    rng = np.random.default_rng(m * n)
    A = np.asfortranarray(rng.standard_normal((m, n)) * np.exp(rng.standard_normal(n))[None, :])
    k = min(m, n)
    W = A.copy(order="F")
    piv = np.zeros(n, dtype=np.int64)
    rd = np.zeros(k)
    used = np.zeros(k, dtype=np.int32)
    emu.run_qrcp(_p(W), C.c_int64(m), C.c_int64(n), C.c_int64(k), _p(piv), _p(rd), two_pass, _p(used))
    _, R, p = sampled.qrcp(A)
    assert sorted(piv.tolist()) == list(range(n))
    assert np.array_equal(piv[:k] + 1, p[:k])
    assert np.max(np.abs(np.abs(rd) - np.abs(np.diag(R)))) < 1e-12 * abs(R[0, 0])
    # the upper triangle left in place is R itself (up to the sign convention of each row)
    Rk = np.triu(W[:k, :])
    assert np.max(np.abs(np.abs(Rk) - np.abs(R[:k, :]))) < 1e-11 * abs(R[0, 0])
    if not two_pass:
        expect = {9: {2}, 70: {4, 2}, 131: {8, 4}, 12: {2}}[m]
        assert set(used[: k - 1 if n <= m else k].tolist()) - {0} == expect, set(used.tolist())
