"""The WHOLE experimental INT8 kernel (csrc/gemm_i8.cu: partial_gemm_i8_kernel, verbatim text) executed on the host against a
small functional model of the hardware it programs:
  * 320 OS threads per CTA (tests/simt_emu.h), CTAs run one after the other;
  * shared memory = one byte array, "shared addresses" = offsets into it; mbarriers = (pending arrivals, pending transaction
    bytes, phase) records keyed by address with init / arrive / expect_tx / try_wait.parity / complete_tx;
  * TMA 2-D tile loads (zero fill outside the tensor) and 1-D bulk copies complete synchronously and post their bytes;
  * TMEM = 128 lanes x 512 int32 columns; tcgen05.alloc hands out column 0; tcgen05.ld 32x32b checks the warp's lane quarter;
  * tcgen05.mma kind::i8 DECODES the 64-bit shared-memory descriptors and the instruction descriptor the kernel built
    (start address, LBO, SBO, major-ness, N) with the canonical-layout formulas of cute/atom/mma_traits_sm100.hpp and
    multiplies exactly; tcgen05.commit arrives on its mbarrier.
The model is synchronous, so it says nothing about speed or about proxy-fence placement; it does execute the kernel's real
control flow: barrier counts and parities, stage indices, descriptor arithmetic, accumulator columns, the epilogue's lane and
column addressing, ragged tiles.  The result must match the FP64 contraction to 1e-12."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "itensorcpd.jl_b200", "csrc", "gemm_i8.cu")
BUILD = os.path.join(ROOT, "oracle", "_build")

MODEL = r"""
#include "simt_emu.h"
#include <algorithm>
#include <map>
#include <mutex>
#define ITCPD_MAX_ORDER 8
#define __grid_constant__
#define __host__
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
inline int atomicMax(int *p, int v) { static std::mutex m; std::lock_guard<std::mutex> g(m); int o = *p; if (v > o) *p = v; return o; }
inline void __trap() { abort(); }
inline void __nanosleep(unsigned) { sched_yield(); }

// ---------------- the modelled machine ----------------
alignas(1024) static uint8_t i8_smem_raw[232448 + 2048];
static const uint32_t SMEM_BASE_ADDR = 0x1230;       // deliberately not 1024-aligned: the kernel must align itself
static int32_t TMEM[128][512];
struct CUtensorMap { const double *base; long d0, d1; int box0, box1; };
struct MBarModel { int count = 0, pending = 0; long tx = 0; unsigned phase = 0; };
static std::map<uint32_t, MBarModel> g_bars;
static std::mutex g_bar_mutex;
static long g_mma_count = 0;

inline uint32_t i8_smem_u32(const void *p) { return SMEM_BASE_ADDR + (uint32_t)((const uint8_t *)p - i8_smem_raw); }
inline uint8_t *smem_ptr(uint32_t addr) { return i8_smem_raw + (addr - SMEM_BASE_ADDR); }
static void bar_flip(MBarModel &b) { if (b.pending == 0 && b.tx == 0) { b.phase++; b.pending = b.count; } }
inline void i8_mbar_init(uint32_t bar, uint32_t count) { std::lock_guard<std::mutex> g(g_bar_mutex); MBarModel b; b.count = b.pending = (int)count; g_bars[bar] = b; }
inline void bar_arrive(uint32_t bar, long tx) {
    std::lock_guard<std::mutex> g(g_bar_mutex);
    auto it = g_bars.find(bar);
    if (it == g_bars.end()) { fprintf(stderr, "arrive on an uninitialised mbarrier 0x%x\n", bar); abort(); }
    it->second.tx += tx; it->second.pending -= 1;
    if (it->second.pending < 0) { fprintf(stderr, "too many arrivals on mbarrier 0x%x\n", bar); abort(); }
    bar_flip(it->second);
}
inline void bar_complete_tx(uint32_t bar, long bytes) { std::lock_guard<std::mutex> g(g_bar_mutex); auto &b = g_bars.at(bar); b.tx -= bytes; bar_flip(b); }
inline void i8_mbar_expect_tx(uint32_t bar, uint32_t bytes) { bar_arrive(bar, (long)bytes); }
inline void i8_mbar_arrive(uint32_t bar) { bar_arrive(bar, 0); }
inline bool i8_mbar_test(uint32_t bar, uint32_t parity) {
    std::lock_guard<std::mutex> g(g_bar_mutex);
    const bool done = (g_bars.at(bar).phase & 1u) != parity;
    if (!done) sched_yield();
    return done;
}
inline void i8_mbar_wait(uint32_t bar, uint32_t parity) {
    for (long spins = 0;; ++spins) {
        { std::lock_guard<std::mutex> g(g_bar_mutex); if ((g_bars.at(bar).phase & 1u) != parity) return; }
        if (spins > 40000000) { fprintf(stderr, "dead-lock: mbarrier 0x%x parity %u (thread %u)\n", bar, parity, threadIdx.x); abort(); }
        sched_yield();
    }
}
inline void i8_tma_2d(uint32_t dst, const CUtensorMap *m, int c0, int c1, uint32_t bar) {
    double *d = (double *)smem_ptr(dst);
    for (int j = 0; j < m->box1; ++j)
        for (int i = 0; i < m->box0; ++i) {
            const long g0 = c0 + i, g1 = c1 + j;
            d[j * m->box0 + i] = (g0 < m->d0 && g1 < m->d1) ? m->base[g0 + m->d0 * g1] : 0.0;
        }
    bar_complete_tx(bar, (long)m->box0 * m->box1 * 8);
}
inline void i8_bulk_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) { memcpy(smem_ptr(dst), src, bytes); bar_complete_tx(bar, bytes); }
inline void i8_fence_mbar_init() {}
inline void i8_fence_proxy_async() {}
inline void i8_tc_fence_before() {}
inline void i8_tc_fence_after() {}
inline void i8_tmem_alloc(uint32_t smem_dst, uint32_t cols) { if (cols != 512) abort(); if ((threadIdx.x & 31) == 0) *(uint32_t *)smem_ptr(smem_dst) = 0u; __syncwarp(); }
inline void i8_tmem_dealloc(uint32_t, uint32_t) {}
inline void i8_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t a0 = (uint32_t)(adesc & 0x3fff) << 4, albo = (uint32_t)((adesc >> 16) & 0x3fff) << 4, asbo = (uint32_t)((adesc >> 32) & 0x3fff) << 4;
    const uint32_t b0 = (uint32_t)(bdesc & 0x3fff) << 4, blbo = (uint32_t)((bdesc >> 16) & 0x3fff) << 4, bsbo = (uint32_t)((bdesc >> 32) & 0x3fff) << 4;
    if (((adesc >> 46) & 3) != 1 || ((bdesc >> 46) & 3) != 1 || (adesc >> 61) != 0 || (bdesc >> 61) != 0) { fprintf(stderr, "bad descriptor version / layout type\n"); abort(); }
    const int N = (int)((idesc >> 17) & 0x3f) << 3, M = (int)((idesc >> 24) & 0x1f) << 4, a_mn = (int)(idesc >> 15) & 1, b_mn = (int)(idesc >> 16) & 1;
    if (M != 128 || b_mn != 0 || ((idesc >> 4) & 3) != 2 || ((idesc >> 7) & 7) != 1 || ((idesc >> 10) & 7) != 1 || N % 16 != 0 || N < 16 || N > 256) { fprintf(stderr, "bad instruction descriptor 0x%x\n", idesc); abort(); }
    // the descriptor holds a 14-bit (>> 4) address: compare modulo 2^18 like the hardware
    auto A = [&](int m, int k) -> int { const uint32_t off = a_mn ? (k % 8) * 16 + (k / 8) * albo + (m / 16) * asbo + m % 16 : (m % 8) * 16 + (m / 8) * asbo + (k / 16) * albo + k % 16; return (int8_t)*smem_ptr(a0 + off); };
    auto B = [&](int n, int k) -> int { const uint32_t off = (n % 8) * 16 + (n / 8) * bsbo + (k / 16) * blbo + k % 16; return (int8_t)*smem_ptr(b0 + off); };
    const int col0 = (int)(d_tmem & 0xffff);
    if ((d_tmem >> 16) != 0 || col0 + N > 512) { fprintf(stderr, "MMA writes outside the TMEM allocation\n"); abort(); }
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            int s = 0;
            for (int k = 0; k < 32; ++k) s += A(m, k) * B(n, k);
            TMEM[m][col0 + n] = (accumulate ? TMEM[m][col0 + n] : 0) + s;
        }
    ++g_mma_count;
}
inline void i8_commit(uint32_t bar) { bar_arrive(bar, 0); }
inline void i8_tmem_ld32_issue(uint32_t taddr, int (&v)[32]);
inline void i8_tmem_ld_wait(int (&)[32]) {}
inline void i8_tmem_ld32(uint32_t taddr, int (&v)[32]) { i8_tmem_ld32_issue(taddr, v); }
inline void i8_tmem_ld32_issue(uint32_t taddr, int (&v)[32]) {
    const int lane0 = (int)(taddr >> 16), col = (int)(taddr & 0xffff), warp = (int)(threadIdx.x >> 5);
    if (lane0 != 32 * (warp % 4) || col + 32 > 512) { fprintf(stderr, "tcgen05.ld outside the warp's lane quarter / allocation\n"); abort(); }
    for (int c = 0; c < 32; ++c) v[c] = TMEM[lane0 + (threadIdx.x & 31)][col + c];
}

namespace itcpd_emu {
@@NUMERICS@@
@@DESCS@@
@@KERNEL@@
}
using namespace itcpd_emu;

static void run_fixup(std::vector<double> &part, int ksplit, int rblocks, long rows_out, int R, double *out) {
    if (ksplit <= 1) return;
    const long n = rows_out * R;
    const unsigned gx = (unsigned)((n + 63) / 64);
    for (unsigned bx = 0; bx < gx; ++bx)
        emu_launch(64, 0, [&] { i8_splitk_fixup_kernel(part.data(), ksplit, rblocks, rows_out, n, out); }, bx, 0, gx, 1);
}
// column exponents and packed digit planes of every 64-column rank block (what launch_partial_gemm_i8 does per block)
static void pack_krp_blocks(I8Krp a, int rblocks, long kext, long ktiles, std::vector<int> &eb, std::vector<uint8_t> &bdig) {
    for (int rb = 0; rb < rblocks; ++rb) {
        a.r0 = rb * I8_BN;
        int *e = eb.data() + rb * I8_BN;
        uint8_t *b = bdig.data() + (size_t)rb * ktiles * I8_B_BYTES;
        unsigned gx = (unsigned)((((kext + 255) / 256) * I8_BN + 63) / 64);
        for (unsigned bx = 0; bx < gx; ++bx) emu_launch(64, 0, [&] { i8_krp_exponent_kernel(a, e); }, bx, 0, gx, 1);
        gx = (unsigned)((ktiles * I8_BN * 2 + 63) / 64);
        for (unsigned bx = 0; bx < gx; ++bx) emu_launch(64, 0, [&] { i8_krp_pack_kernel(a, e, ktiles, b); }, bx, 0, gx, 1);
    }
}

// out (rows_out x R) = unfolding(T) * K  through the emulated kernel; exponents and digit planes by the emulated helpers
extern "C" long emu_gemm_i8(int kind, const double *T, long Mrows, long Ncols, int nf, const double *const *fac, const long *ext, int R, int grid, int ksplit, double *out) {
    const long kext = kind == 0 ? Ncols : Mrows, rows_out = kind == 0 ? Mrows : Ncols;
    const long ktiles = (kext + I8_BK - 1) / I8_BK, row_tiles = (rows_out + I8_BM - 1) / I8_BM;
    const int rblocks = (R + I8_BN - 1) / I8_BN;
    std::vector<int> ea(row_tiles * I8_BM, I8_EXP_ZERO), eb(rblocks * I8_BN, I8_EXP_ZERO);
    if (kind == 0) {
        const unsigned gx = (unsigned)((rows_out + 63) / 64), gy = 2;
        for (unsigned bx = 0; bx < gx; ++bx) for (unsigned by = 0; by < gy; ++by)
            emu_launch(64, 0, [&] { i8_row_exponent_strided_kernel(T, rows_out, kext, Mrows, ea.data()); }, bx, by, gx, gy);
    } else {
        const unsigned gx = (unsigned)((rows_out * 32 + 63) / 64);
        for (unsigned bx = 0; bx < gx; ++bx) emu_launch(64, 0, [&] { i8_row_exponent_contig_kernel(T, rows_out, kext, Mrows, ea.data()); }, bx, 0, gx, 1);
    }
    I8Krp a;
    memset(&a, 0, sizeof(a));
    a.nf = nf; a.kext = kext; a.R = R;
    for (int f = 0; f < nf; ++f) { a.fac[f] = fac[f]; a.ext[f] = ext[f]; a.dim[f] = ext[f]; }
    std::vector<uint8_t> bdig((size_t)rblocks * ktiles * I8_B_BYTES);
    unsigned gx = 0;
    pack_krp_blocks(a, rblocks, kext, ktiles, eb, bdig);
    CUtensorMap map;
    map.base = T; map.d0 = Mrows; map.d1 = Ncols;
    if (kind == 0) { map.box0 = I8_BM; map.box1 = I8_BK; } else { map.box0 = I8_BK; map.box1 = I8_BM; }
    g_mma_count = 0;
    const int kchunk = (int)((ktiles + ksplit - 1) / ksplit);
    if ((ktiles + kchunk - 1) / kchunk != ksplit) return -1;          // the host never launches an empty chunk
    std::vector<double> part(ksplit > 1 ? (size_t)ksplit * rblocks * rows_out * I8_BN : 1, NAN);
    I8Sched sc;
    sc.num_row_tiles = (int)row_tiles; sc.kt_count = (int)ktiles; sc.ksplit = ksplit; sc.kchunk = kchunk; sc.rblocks = rblocks;
    sc.bdig_rb_stride = (long long)ktiles * I8_B_BYTES;
    for (int cta = 0; cta < grid; ++cta) {
        g_bars.clear();
        memset(TMEM, 0x5a, sizeof(TMEM));          // stale accumulator contents must not leak into results
        if (kind == 0) emu_launch(320, 0, [&] { partial_gemm_i8_kernel<0>(map, bdig.data(), ea.data(), eb.data(), out, rows_out, R, sc, part.data()); }, cta, 0, grid, 1);
        else emu_launch(320, 0, [&] { partial_gemm_i8_kernel<1>(map, bdig.data(), ea.data(), eb.data(), out, rows_out, R, sc, part.data()); }, cta, 0, grid, 1);
    }
    run_fixup(part, ksplit, rblocks, rows_out, R, out);
    return g_mma_count;
}

// the pre-packed variant: digit planes of T built once (i8_pack_tensor_kernel), then partial_gemm_i8p_kernel
extern "C" long emu_gemm_i8p(int kind, const double *T, long Mrows, long Ncols, int nf, const double *const *fac, const long *ext, int R, int grid, int ksplit, double *out) {
    const long kext = kind == 0 ? Ncols : Mrows, rows_out = kind == 0 ? Mrows : Ncols;
    const long ktiles = (kext + I8_BK - 1) / I8_BK, row_tiles = (rows_out + I8_BM - 1) / I8_BM;
    const int rblocks = (R + I8_BN - 1) / I8_BN;
    std::vector<int> ea(row_tiles * I8_BM, I8_EXP_ZERO), eb(rblocks * I8_BN, I8_EXP_ZERO);
    if (kind == 0) {
        const unsigned gx = (unsigned)((rows_out + 63) / 64);
        for (unsigned bx = 0; bx < gx; ++bx) emu_launch(64, 0, [&] { i8_row_exponent_strided_kernel(T, rows_out, kext, Mrows, ea.data()); }, bx, 0, gx, 1);
    } else {
        const unsigned gx = (unsigned)((rows_out * 32 + 63) / 64);
        for (unsigned bx = 0; bx < gx; ++bx) emu_launch(64, 0, [&] { i8_row_exponent_contig_kernel(T, rows_out, kext, Mrows, ea.data()); }, bx, 0, gx, 1);
    }
    I8Krp a;
    memset(&a, 0, sizeof(a));
    a.nf = nf; a.kext = kext; a.R = R;
    for (int f = 0; f < nf; ++f) { a.fac[f] = fac[f]; a.ext[f] = ext[f]; a.dim[f] = ext[f]; }
    std::vector<uint8_t> bdig((size_t)rblocks * ktiles * I8_B_BYTES), adig((size_t)row_tiles * ktiles * I8_A_BYTES);
    unsigned gx = 0;
    pack_krp_blocks(a, rblocks, kext, ktiles, eb, bdig);
    gx = (unsigned)(row_tiles * ktiles);
    for (unsigned bx = 0; bx < gx; ++bx) {
        if (kind == 0) emu_launch(256, 0, [&] { i8_pack_tensor_kernel<0>(T, rows_out, kext, 1, Mrows, ea.data(), ktiles, adig.data()); }, bx, 0, gx, 1);
        else emu_launch(256, 0, [&] { i8_pack_tensor_kernel<1>(T, rows_out, kext, Mrows, 1, ea.data(), ktiles, adig.data()); }, bx, 0, gx, 1);
    }
    g_mma_count = 0;
    const int kchunk = (int)((ktiles + ksplit - 1) / ksplit);
    if ((ktiles + kchunk - 1) / kchunk != ksplit) return -1;          // the host never launches an empty chunk
    std::vector<double> part(ksplit > 1 ? (size_t)ksplit * rblocks * rows_out * I8_BN : 1, NAN);
    I8Sched sc;
    sc.num_row_tiles = (int)row_tiles; sc.kt_count = (int)ktiles; sc.ksplit = ksplit; sc.kchunk = kchunk; sc.rblocks = rblocks;
    sc.bdig_rb_stride = (long long)ktiles * I8_B_BYTES;
    for (int cta = 0; cta < grid; ++cta) {
        g_bars.clear();
        memset(TMEM, 0x5a, sizeof(TMEM));
        if (kind == 0) emu_launch(192, 0, [&] { partial_gemm_i8p_kernel<0>(adig.data(), bdig.data(), ea.data(), eb.data(), out, rows_out, R, sc, part.data()); }, cta, 0, grid, 1);
        else emu_launch(192, 0, [&] { partial_gemm_i8p_kernel<1>(adig.data(), bdig.data(), ea.data(), eb.data(), out, rows_out, R, sc, part.data()); }, cta, 0, grid, 1);
    }
    run_fixup(part, ksplit, rblocks, rows_out, R, out);
    return g_mma_count;
}
"""


def model_source():
    """the functional model with the VERBATIM device text of csrc/gemm_i8.cu spliced in (also used by test_i8_host_device_cpu.py)"""
    text = open(SRC).read()
    numerics = text[text.index("constexpr int I8_NDIG = 6;"):text.index("#ifndef ITCPD_I8_HOST_EMULATION")]
    descs = text[text.index("// shared-memory matrix descriptor, SWIZZLE_NONE"):text.index("// ------------------------------------------------------------------------------------------------------------------\n// the kernel")]
    k0 = text.index("template <int KIND>\n__global__ void __launch_bounds__(320, 1)")
    k1 = text.index("// ------------------------------------------------------------------------------------------------------------------\n// host side")
    kernel = text[k0:k1].replace("extern __shared__ uint8_t i8_smem_raw[];", "")
    fill = text[text.index("__global__ void i8_fill_int_kernel"):text.index("// Same contract as launch_partial_gemm")]
    return MODEL.replace("@@NUMERICS@@", numerics).replace("@@DESCS@@", descs).replace("@@KERNEL@@", kernel + "\n" + fill)


@pytest.fixture(scope="module")
def sim():
    os.makedirs(BUILD, exist_ok=True)
    cpp, so = os.path.join(BUILD, "i8_sim.cpp"), os.path.join(BUILD, "i8_sim.so")
    open(cpp, "w").write(model_source())
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-Wl,-Bsymbolic", "-I", os.path.join(ROOT, "tests"), "-o", so, cpp,
                    "-lpthread"], check=True, capture_output=True)
    lib = C.CDLL(so)
    lib.emu_gemm_i8.restype = C.c_long
    lib.emu_gemm_i8p.restype = C.c_long
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("variant", ["on_the_fly", "prepacked"])
@pytest.mark.parametrize("kind,Mrows,Ncols,R,grid,ksplit", [(0, 256, 64, 48, 2, 1), (1, 64, 256, 64, 2, 1), (0, 200, 40, 20, 1, 1), (1, 40, 330, 33, 3, 1),
                                                          (0, 384, 160, 64, 2, 1),
                                                          # split-K (short-and-wide contractions): units = row tiles x k-chunks, FP64 partial tiles + fix-up
                                                          (0, 128, 320, 64, 3, 5), (1, 352, 100, 24, 2, 4), (0, 200, 200, 40, 5, 3),
                                                          # several rank blocks of 64 columns in ONE launch (rank block fastest in the unit index)
                                                          (0, 256, 64, 100, 2, 1), (1, 96, 200, 130, 3, 1), (0, 128, 320, 70, 4, 2)])
def test_whole_i8_kernel_on_the_functional_model(sim, kind, Mrows, Ncols, R, grid, ksplit, variant):
    rng = np.random.default_rng(100 * kind + Mrows)
    T = np.asfortranarray(rng.standard_normal((Mrows, Ncols)) * np.exp2(rng.integers(-5, 6, size=(Mrows, 1))))   # memory image T[m + Mrows n]
    kext = Ncols if kind == 0 else Mrows
    e1 = 8 if kext % 8 == 0 else 5 if kext % 5 == 0 else 1
    f1 = np.asfortranarray(rng.standard_normal((e1, R)))
    f2 = np.asfortranarray(rng.standard_normal((kext // e1, R)))
    Kr = (f2[:, None, :] * f1[None, :, :]).reshape(kext, R)
    rows_out = Mrows if kind == 0 else Ncols
    out = np.full((rows_out, R), np.nan, order="F")
    fac = (C.c_void_p * 2)(f1.ctypes.data, f2.ctypes.data)
    ext = np.array([e1, kext // e1], dtype=np.int64)
    run = sim.emu_gemm_i8 if variant == "on_the_fly" else sim.emu_gemm_i8p
    nmma = run(kind, _p(T), Mrows, Ncols, 2, fac, _p(ext), R, grid, ksplit, _p(out))
    ref = (T.astype(np.longdouble) @ Kr.astype(np.longdouble)) if kind == 0 else (T.T.astype(np.longdouble) @ Kr.astype(np.longdouble))
    ref = ref.astype(np.float64)
    assert np.all(np.isfinite(out)), "some outputs were never written"
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert err < 1e-12, err
    ktiles, row_tiles = -(-kext // 32), -(-rows_out // 128)
    assert nmma == 9 * ktiles * row_tiles * -(-R // 64)       # 27 digit products per k-step and rank block as 9 instructions


def test_split_k_schedule_policy():
    """i8_choose_ksplit (host, verbatim text): no empty chunk, data parallel when there are many row tiles, whole waves otherwise"""
    text = open(SRC).read()
    a = text.index("void i8_choose_ksplit(")
    fn = text[a:text.index("__global__ void i8_fill_int_kernel")]
    os.makedirs(BUILD, exist_ok=True)
    cpp, so = os.path.join(BUILD, "i8_ksplit.cpp"), os.path.join(BUILD, "i8_ksplit.so")
    maxk = re.search(r"constexpr int I8_MAX_KCHUNK = (\d+);", text).group(1)
    open(cpp, "w").write("#include <algorithm>\n#include <cstdint>\nstatic inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }\n"
                         f"constexpr int I8_MAX_KCHUNK = {maxk};\n"
                         + fn + '\nextern "C" void choose(long rt, long kt, int sms, int *ks, int *kc) { i8_choose_ksplit(rt, kt, sms, ks, kc); }\n')
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-o", so, cpp], check=True, capture_output=True)
    lib = C.CDLL(so)

    def choose(rt, kt, sms=148):
        ks, kc = C.c_int(), C.c_int()
        lib.choose(C.c_long(rt), C.c_long(kt), sms, C.byref(ks), C.byref(kc))
        return ks.value, kc.value

    for rt, kt in [(8192, 32), (8, 4096), (16, 16384), (313, 7), (1, 10), (1, 100000), (147, 64), (149, 64), (600, 9), (3, 24), (512, 2048), (4096, 3000)]:
        ks, kc = choose(rt, kt)
        assert ks >= 1 and kc >= 1 and -(-kt // kc) == ks, (rt, kt, ks, kc)        # every chunk holds at least one k-tile
        assert kc <= int(maxk), (rt, kt, ks, kc)        # 6 pairs x kc x 32 x 2^14 < 2^31: the int32 accumulators stay exact
        assert 6 * int(maxk) * 32 * 2 ** 14 < 2 ** 31
        assert ks == -(-kt // int(maxk)) or kc >= 8, (rt, kt, ks, kc)
        if rt >= 4 * 148:
            assert ks == -(-kt // int(maxk))
    # the per-rank slabs of configs B and D at 8 GPUs (pass A of the (1,1) tree): the units fill the SMs to >= 95 % in a few whole waves
    for rt, kt in [(8, 4096), (16, 16384)]:
        ks, kc = choose(rt, kt)
        units = rt * ks
        waves = -(-units // 148)
        assert waves <= 4 and units / (148 * waves) >= 0.95, (rt, kt, ks, kc)


@pytest.mark.parametrize("variant", ["on_the_fly", "prepacked"])
@pytest.mark.parametrize("kind", [0, 1])
def test_non_finite_inputs_become_nan_rows_and_columns(sim, kind, variant):
    """A NaN / Inf in T poisons its output row, one in a factor its rank column -- what the FP64 contraction does, so that the
    host still raises the reference's "Error NAN" (fit_check.jl:40-42); every other entry keeps the 1e-12 bar."""
    rng = np.random.default_rng(5)
    Mrows, Ncols, R = (256, 96, 40) if kind == 0 else (96, 256, 40)
    T = np.asfortranarray(rng.standard_normal((Mrows, Ncols)))
    kext = Ncols if kind == 0 else Mrows
    rows_out = Mrows if kind == 0 else Ncols
    bad_rows = [3, 130]
    if kind == 0:
        T[3, 7], T[130, 50] = np.nan, -np.inf
    else:
        T[7, 3], T[50, 130] = np.inf, np.nan
    f1 = np.asfortranarray(rng.standard_normal((8, R)))
    f2 = np.asfortranarray(rng.standard_normal((kext // 8, R)))
    f2[5, 11] = np.nan
    Kr = (f2[:, None, :] * f1[None, :, :]).reshape(kext, R)
    out = np.full((rows_out, R), 777.0, order="F")
    fac = (C.c_void_p * 2)(f1.ctypes.data, f2.ctypes.data)
    ext = np.array([8, kext // 8], dtype=np.int64)
    run = sim.emu_gemm_i8 if variant == "on_the_fly" else sim.emu_gemm_i8p
    run(kind, _p(T), Mrows, Ncols, 2, fac, _p(ext), R, 2, 1, _p(out))
    assert np.all(np.isnan(out[bad_rows, :])) and np.all(np.isnan(out[:, 11]))
    good_r = [r for r in range(R) if r != 11]
    good_m = [m for m in range(rows_out) if m not in bad_rows]
    Tc = np.where(np.isfinite(T), T, 0.0)
    ref = (Tc @ Kr[:, good_r]) if kind == 0 else (Tc.T @ Kr[:, good_r])
    got = out[np.ix_(good_m, good_r)]
    assert np.all(np.isfinite(got))
    assert np.linalg.norm(got - ref[good_m]) / np.linalg.norm(ref[good_m]) < 1e-12
