"""CPU emulation of barrier-free CUDA kernels: the kernel text is taken verbatim from csrc/sampled_sharded.cu, compiled with
g++ behind a shim that turns blockIdx/threadIdx into loop variables, and compared with the oracle.  It covers the index
arithmetic of the owner-rank gather kernels of the sharded sampled path, which have not run on hardware yet."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import sampled

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "itensorcpd.jl_b200", "csrc", "sampled_sharded.cu")

SHIM = r"""
#include <cstdint>
#include <cstring>
#define ITCPD_MAX_ORDER 8
#define __global__
#define __restrict__
#define __launch_bounds__(x)
struct dim3_ { unsigned x, y, z; };
static thread_local dim3_ blockIdx, threadIdx, blockDim, gridDim;
"""

DRIVER = r"""
template <typename F> static void launch(unsigned grid, unsigned block, F f) {
    gridDim = {grid, 1, 1}; blockDim = {block, 1, 1};
    for (unsigned b = 0; b < grid; ++b) for (unsigned t = 0; t < block; ++t) { blockIdx = {b, 0, 0}; threadIdx = {t, 0, 0}; f(); }
}
extern "C" void run_pivot_hadamard_owned(int n, const int64_t *dims, const double *const *fac, int mode, int R, int64_t nsamp,
                                         const int64_t *piv, int64_t off, double *K) {
    ShFac fp; ShDims d; d.n = n;
    for (int i = 0; i < n; ++i) { fp.a[i] = fac[i]; d.ext[i] = dims[i]; d.dim[i] = dims[i]; }
    launch(7, 64, [&] { pivot_hadamard_owned_kernel(fp, d, mode, R, nsamp, piv, off, K); });
}
extern "C" void run_gather_fibers_owned(int n, const int64_t *dims, const double *T, int mode, int64_t nsamp, const int64_t *piv,
                                        int64_t off, double *out) {
    ShDims d; d.n = n;
    for (int i = 0; i < n; ++i) { d.ext[i] = dims[i]; d.dim[i] = dims[i]; }
    launch((unsigned)nsamp, 128, [&] { gather_fibers_owned_kernel(T, d, mode, nsamp, piv, off, out); });
}
"""


@pytest.fixture(scope="module")
def emu():
    text = open(SRC).read()
    start = text.index("struct ShDims")
    end = text.index("int64_t sharded_last_rows")
    body = re.sub(r"static ShDims shdims\(const itcpd_ctx \*c\) \{.*?\n\}\n", "", text[start:end], flags=re.S)
    with tempfile.TemporaryDirectory() as td:
        cpp = os.path.join(td, "emu.cpp")
        so = os.path.join(td, "emu.so")
        open(cpp, "w").write(SHIM + body + DRIVER)
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-o", so, cpp], check=True, capture_output=True)
        lib = C.CDLL(so)
        yield lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("dims,world", [((6, 5, 8), 2), ((4, 3, 5, 6), 3)])
def test_owned_gather_kernels_match_oracle(emu, dims, world):
    rng = np.random.default_rng(0)
    N, R, nsamp = len(dims), 3, 50
    T = np.asfortranarray(rng.standard_normal(dims))
    factors = [np.asfortranarray(rng.standard_normal((d, R))) for d in dims]
    slab = dims[-1] // world
    for rank in range(world):
        off = rank * slab
        Tl = np.asfortranarray(T[..., off:off + slab])
        ldims = np.array(dims[:-1] + (slab,), dtype=np.int64)
        lf = factors[:-1] + [np.asfortranarray(factors[-1][off:off + slab])]
        fac_ptrs = (C.c_void_p * N)(*[f.ctypes.data for f in lf])
        for mode in range(N - 1):
            others = [m for m in range(N) if m != mode]
            piv = np.asfortranarray(np.stack([rng.integers(1, dims[m] + 1, size=nsamp) for m in others], axis=1).astype(np.int64))
            g = piv[:, -1] - 1 - off
            owned = (g >= 0) & (g < slab)
            lp = piv.copy()
            lp[:, -1] = np.where(owned, g + 1, 1)
            K = np.zeros((nsamp, R), order="F")
            emu.run_pivot_hadamard_owned(N, _ptr(ldims), fac_ptrs, mode, R, C.c_int64(nsamp), _ptr(piv), C.c_int64(off), _ptr(K))
            want = sampled.pivot_hadamard([lf[m] for m in others], lp) * owned[:, None]
            assert np.array_equal(K, want), (rank, mode)
            out = np.full((dims[mode], nsamp), np.nan, order="F")
            emu.run_gather_fibers_owned(N, _ptr(ldims), _ptr(Tl), mode, C.c_int64(nsamp), _ptr(piv), C.c_int64(off), _ptr(out))
            wantT = sampled.fused_flatten_sample(Tl, mode, lp) * owned[None, :]
            assert np.array_equal(out, wantT), (rank, mode)
            assert owned.any() and (~owned).any()
