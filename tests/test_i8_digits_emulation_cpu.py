"""The numerical half of the experimental INT8 tensor-core contraction (csrc/gemm_i8.cu), executed on the host: exponent
kernels, digit extraction, the packed Khatri-Rao digit planes, the converter threads and the final combination are compiled
VERBATIM behind tests/simt_emu.h; the tensor-core step itself (tcgen05.mma kind::i8: exact int8 x int8 -> int32) is replaced
by exact integer matrix products on digit planes DECODED from the canonical UMMA shared-memory layouts as documented in
cute/atom/mma_traits_sm100.hpp.  What this pins: the digit arithmetic, the layouts the kernel writes, the exponent
bookkeeping (the 2^(ea+eb-14) of the epilogue) and the accuracy (< 1e-12 relative Frobenius, BASELINE.json north_star).
What it cannot pin: descriptors, barriers and TMEM addressing -- the kernel has not run on hardware yet."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "itensorcpd.jl_b200", "csrc", "gemm_i8.cu")
BUILD = os.path.join(ROOT, "oracle", "_build")

HARNESS = r"""
#include "simt_emu.h"
#include <algorithm>
#define ITCPD_MAX_ORDER 8
using std::min;
struct uint4 { unsigned x, y, z, w; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline long long __double2ll_rn(double x) { return llrint(x); }
inline double __hiloint2double(int hi, int lo) { unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo; double d; memcpy(&d, &b, 8); return d; }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const unsigned long long xy = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((xy >> (8 * ((s >> (4 * i)) & 7))) & 0xffu) << (8 * i);
    return r;
}
inline int atomicMax(int *p, int v) { static std::atomic_flag l = ATOMIC_FLAG_INIT; while (l.test_and_set()) {} int o = *p; if (v > o) *p = v; l.clear(); return o; }
namespace itcpd_emu {
%(section)s
}
using namespace itcpd_emu;

extern "C" void emu_row_exponents(int kind, const double *base, long nrows, long nred, long stride, int *E) {
    for (long i = 0; i < nrows; ++i) E[i] = I8_EXP_ZERO;
    if (kind == 0) {
        const unsigned gx = (unsigned)((nrows + 63) / 64), gy = 3;
        for (unsigned bx = 0; bx < gx; ++bx)
            for (unsigned by = 0; by < gy; ++by)
                emu_launch(64, 0, [&] { i8_row_exponent_strided_kernel(base, nrows, nred, stride, E); }, bx, by, gx, gy);
    } else {
        const unsigned gx = (unsigned)((nrows * 32 + 63) / 64);
        for (unsigned bx = 0; bx < gx; ++bx) emu_launch(64, 0, [&] { i8_row_exponent_contig_kernel(base, nrows, nred, stride, E); }, bx, 0, gx, 1);
    }
}
extern "C" void emu_krp_pack(int nf, const double *const *fac, const long *ext, long kext, int R, int *E, unsigned char *out) {
    I8Krp a;
    memset(&a, 0, sizeof(a));
    a.nf = nf; a.kext = kext; a.R = R;
    for (int f = 0; f < nf; ++f) { a.fac[f] = fac[f]; a.ext[f] = ext[f]; a.dim[f] = ext[f]; }
    for (int i = 0; i < I8_BN; ++i) E[i] = I8_EXP_ZERO;
    const long ktiles = kext / I8_BK;
    unsigned gx = (unsigned)((((kext + 255) / 256) * I8_BN + 63) / 64);
    for (unsigned bx = 0; bx < gx; ++bx) emu_launch(64, 0, [&] { i8_krp_exponent_kernel(a, E); }, bx, 0, gx, 1);
    gx = (unsigned)((ktiles * I8_BN * 2 + 63) / 64);
    for (unsigned bx = 0; bx < gx; ++bx) emu_launch(64, 0, [&] { i8_krp_pack_kernel(a, E, ktiles, out); }, bx, 0, gx, 1);
}
extern "C" void emu_convert_tile(int kind, const double *F, const int *ea_tile, unsigned char *A) {
    emu_launch(256, 0, [&] { if (kind == 0) i8_convert_thread<0>(F, ea_tile, A, (int)threadIdx.x); else i8_convert_thread<1>(F, ea_tile, A, (int)threadIdx.x); });
}
extern "C" double emu_combine(const long long *acc7, int em, int er) {
    double v = 0.0;
    for (int t = I8_NACC - 1; t >= 0; --t) v = fma(i8_i2d((int)acc7[t]), i8_weight(t), v);
    return i8_finish(v, em, er);
}
extern "C" int emu_consts(int which) {
    return which == 0 ? I8_B_BYTES : which == 1 ? I8_A_BYTES : which == 2 ? I8_EXP_ZERO : which == 3 ? I8_A_PLANE : which == 4 ? I8_NDIG : which == 5 ? I8_NACC : which == 6 ? I8_FRAC : I8_NDIG_B;
}
"""


@pytest.fixture(scope="module")
def emu():
    text = open(SRC).read()
    start = text.index("constexpr int I8_NDIG = 6;")
    end = text.index("#ifndef ITCPD_I8_HOST_EMULATION")
    os.makedirs(BUILD, exist_ok=True)
    cpp, so = os.path.join(BUILD, "i8_emu.cpp"), os.path.join(BUILD, "i8_emu.so")
    open(cpp, "w").write(HARNESS % {"section": text[start:end]})
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-ffp-contract=off", "-Wl,-Bsymbolic", "-I", os.path.join(ROOT, "tests"), "-o", so, cpp,
                    "-lpthread"], check=True, capture_output=True)
    lib = C.CDLL(so)
    lib.emu_combine.restype = C.c_double
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def decode_kmajor(plane_bytes, rows):
    """canonical K-major INTERLEAVE layout, 32 int8 per row: byte(r, kk) = (r%8)*16 + (r/8)*256 + (kk/16)*128 + kk%16"""
    out = np.zeros((rows, 32), dtype=np.int64)
    b = plane_bytes.view(np.int8)
    for r in range(rows):
        for kk in range(32):
            out[r, kk] = b[(r % 8) * 16 + (r // 8) * 256 + (kk // 16) * 128 + kk % 16]
    return out


def decode_mnmajor(plane_bytes, rows):
    """canonical MN-major INTERLEAVE layout: byte(n, kk) = (kk%8)*16 + (kk/8)*128 + (n/16)*512 + n%16"""
    out = np.zeros((rows, 32), dtype=np.int64)
    b = plane_bytes.view(np.int8)
    for n in range(rows):
        for kk in range(32):
            out[n, kk] = b[(kk % 8) * 16 + (kk // 8) * 128 + (n // 16) * 512 + n % 16]
    return out


@pytest.mark.parametrize("kind", [0, 1])
def test_i8_digit_pipeline_reproduces_the_fp64_contraction(emu, kind):
    rng = np.random.default_rng(3 + kind)
    M, K, R = 256, 64, 48                       # two 128-row tiles, two k-tiles, 48 of the 64 columns used
    B_BYTES, A_BYTES, EXP_ZERO, A_PLANE, NDIG, NACC, FRAC, NDIG_B = (emu.emu_consts(i) for i in range(8))
    assert (NDIG, NDIG_B, NACC, FRAC) == (6, 7, 7, 48)          # 6 balanced base-256 digits, accumulators t = p + q = 0 .. 6, 48-bit fixed point
    f1 = rng.standard_normal((8, R)); f1 /= np.linalg.norm(f1, axis=0)
    f2 = rng.standard_normal((K // 8, R)); f2 /= np.linalg.norm(f2, axis=0)
    f1, f2 = np.asfortranarray(f1), np.asfortranarray(f2)
    Kr = (f2[:, None, :] * f1[None, :, :]).reshape(K, R)              # k = i1 + 8 i2  (first factor fastest)
    A = rng.standard_normal((M, K)) * np.exp2(rng.integers(-6, 7, size=(M, 1)))   # rows of very different scale
    A[5, :] = 0.0                                                        # an all-zero row
    A[9, 3] = np.nextafter(4.0, 0.0)                                     # the top of a binade: takes the next exponent (top digit stays < 128)
    A[9, 4] = -np.nextafter(4.0, 0.0)
    A[11, :] *= 2.0 ** -6                                                # small entries beside a maximum JUST below the next-exponent threshold:
    A[11, 2] = np.nextafter(2.0 * (1.0 - 2.0 ** -7), 0.0)                # the largest top digit (+127) that can occur, with carries from below
    A[11, 7] = -np.nextafter(2.0 * (1.0 - 2.0 ** -7), 0.0)
    A[11, 8] = 2.0 * (1.0 - 2.0 ** -7) - 2.0 ** -40
    # memory image of the tensor view: kind 0 has the output rows contiguous (T[m + M k]), kind 1 the contraction index
    if kind == 0:
        img = np.asfortranarray(A)
        stride = M
    else:
        img = np.ascontiguousarray(A)
        stride = K
    ea = np.zeros(M, dtype=np.int32)
    emu.emu_row_exponents(kind, _p(img), M, K, stride, _p(ea))
    amax = np.max(np.abs(A), axis=1)
    for m in range(M):
        if amax[m] == 0:
            assert ea[m] == EXP_ZERO
        else:
            z = amax[m] * 2.0 ** (-float(ea[m]))
            assert z < 0.5 - 2.0 ** -8 and z >= 0.25 - 2.0 ** -9     # below 1/2 - 2^-8, with at most one binade of slack
    eb = np.zeros(64, dtype=np.int32)
    ktiles = K // 32
    Bdig = np.zeros(ktiles * B_BYTES, dtype=np.uint8)
    fac = (C.c_void_p * 2)(f1.ctypes.data, f2.ctypes.data)
    ext = np.array([8, K // 8], dtype=np.int64)
    emu.emu_krp_pack(2, fac, _p(ext), K, R, _p(eb), _p(Bdig))
    assert np.all(eb[R:] == EXP_ZERO)
    acc = np.zeros((M, 64, NACC), dtype=np.int64)
    XA = np.zeros((M, K), dtype=object)
    for tile in range(M // 128):
        rows = slice(128 * tile, 128 * tile + 128)
        for kt in range(ktiles):
            ks = slice(32 * kt, 32 * kt + 32)
            F = np.ascontiguousarray(A[rows, ks].T if kind == 0 else A[rows, ks])      # [k][m] or [n][k] as TMA would land it
            Adig = np.zeros(A_BYTES, dtype=np.uint8)
            emu.emu_convert_tile(kind, _p(F), _p(ea[rows]), _p(Adig))
            dA = [(decode_kmajor if kind == 0 else decode_mnmajor)(Adig[p * A_PLANE:(p + 1) * A_PLANE], 128) for p in range(NDIG)]
            blk = Bdig[kt * B_BYTES:(kt + 1) * B_BYTES]
            dBall = decode_kmajor(blk, NDIG_B * 64)                                       # planes stacked along N: row = q*64 + n
            for p in range(NDIG):
                XA[rows, ks] += dA[p].astype(object) * (256 ** (NDIG - 1 - p))
                for q in range(NACC - p):                                                # the 27 pairs with p + q <= 6 (7 digits on the Khatri-Rao side)
                    acc[rows, :, p + q] += dA[p] @ dBall[q * 64:(q + 1) * 64].T
                assert -128 <= np.min(dA[p]) and np.max(dA[p]) <= 127
    # digits reconstruct the rounded fixed-point value exactly
    for m in (0, 5, 9, 11, 17, 200):
        sc = 0.0 if ea[m] == EXP_ZERO else 2.0 ** (FRAC - float(ea[m]))
        assert all(int(XA[m, k]) == int(np.rint(A[m, k] * sc)) for k in range(K))
    out = np.zeros((M, R))
    for m in range(M):
        for r in range(R):
            a7 = np.ascontiguousarray(acc[m, r, :])
            assert np.max(np.abs(a7)) < 2 ** 31
            out[m, r] = emu.emu_combine(_p(a7), int(ea[m]), int(eb[r]))
    ref = (A.astype(np.longdouble) @ Kr.astype(np.longdouble)).astype(np.float64)
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert err < 1e-12, err
    rowerr = np.linalg.norm(out - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-300)
    assert np.max(rowerr[amax > 0]) < 1e-11, np.max(rowerr)      # per-row scaling: small rows are as accurate as large ones
    assert np.all(out[5] == 0.0)
