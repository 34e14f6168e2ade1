// EXPERIMENTAL (option "gemm_i8", off by default; written without GPU access, not yet run on hardware):
// the partial contraction of the dimension tree on the INT8 tensor cores (tcgen05.mma kind::i8, TMEM accumulators)
// with FP64-equivalent accuracy, by splitting the operands into balanced base-256 digits (an Ozaki-style scheme): 6 digits for
// the tensor, 7 for the Khatri-Rao operand.
//
//   out[m, r] = sum_k T[m, k] K[k, r]                                  (kind 0; kind 1 is the transposed view of T)
//   T[m, k] ~= XA[m, k] 2^(ea[m] - 48),  K[k, r] ~= XB[k, r] 2^(eb[r] - 56),   signed fixed point relative to the row / column maximum
//   XA = sum_{p=0..5} dA_p 256^(5-p),  XB = sum_{q=0..6} dB_q 256^(6-q),  d in [-128, 127]: the bytes of X + 0x80..80 with bit 7 of
//   the low bytes flipped ARE the digits as int8 -- one 64-bit add and one xor per element
//   sum_k XA XB = sum_{p,q} 256^(11-p-q) S_pq,   S_pq = sum_k dA_p dB_q   exact in int32 while 6 K 2^14 < 2^31 (K <= 21845: the
//   split-K schedule keeps a chunk below that)
// The 27 digit pairs with p + q <= 6 are formed; all pairs of equal weight p + q = t share ONE int32 accumulator in TMEM, so a
// 128 x 64 tile needs 7 x 64 = 448 of the 512 TMEM columns.  Stacking the B digit planes along N turns the 27 products of a k-step
// into 9 instructions:  A_p (128 x 32)  x  [B_0; ..; B_{6-p}]  ->  accumulators t = p .. 6.
// Accuracy (relative Frobenius, profiles/r1_ozaki_int8_numerics.txt): 2e-14 .. 3e-14 for Gaussian T, set by the 48 bits of T relative
// to its row maxima; the Khatri-Rao operand keeps 56 bits because products of factor entries are heavy-tailed (column maximum 20 x the
// typical entry for two factors, more for higher orders) and its planes cost next to nothing.  The first draft (7 base-128 digits
// for both, 28 products, 7 bytes per element of T) gave 5e-14 .. 1e-13.
// T costs 6 bytes per element as pre-packed digit planes (variant 2), so the pass is bound by that HBM stream (0.98 ms at 1024^3)
// instead of the FP64 pipe (3.9 ms); at 4.5 POPS the 27 products take 0.82 ms.
//
// Warp roles (10 warps, 1 CTA per SM, persistent over 128-row tiles):
//   warps 0-7  converters: FP64 tile (TMA) -> 6 int8 digit planes in the canonical UMMA shared-memory layout
//   warps 0-3  also the epilogue: TMEM -> registers -> sum_t acc_t 2^(-8t) in FP64 -> scale -> global
//   warp 8     TMA producer (FP64 tiles of T, packed digit planes of the Khatri-Rao operand)
//   warp 9     TMEM allocation + the single-thread MMA issuer
#include "common.cuh"
#include <cfloat>

namespace itcpd {

constexpr int I8_NDIG = 6;              // base-256 digits of the tensor operand (6 bytes per element of T)
constexpr int I8_NDIG_B = 7;            // base-256 digits of the Khatri-Rao operand: it is small, and products of factor entries are
                                        // heavy-tailed (column maximum >> typical entry), so it gets a digit more than T
constexpr int I8_NACC = 7;              // int32 accumulators per output element: digit pairs with p + q = t, t = 0 .. 6
constexpr int I8_FRAC = 8 * I8_NDIG;    // fixed-point bits (48 / 56); i8_exponent keeps |x| 2^-E below 1/2 - 2^-8 so the top digit stays < 128
constexpr int I8_FRAC_B = 8 * I8_NDIG_B;
constexpr int I8_MAX_KCHUNK = 680;      // k-tiles one accumulator may sum: 6 pairs x 680 x 32 x 2^14 < 2^31
constexpr int I8_BM = 128, I8_BN = 64, I8_BK = 32;
constexpr int I8_FSTAGES = 3, I8_DSTAGES = 2;
constexpr int I8_F_BYTES = I8_BM * I8_BK * 8;                  // 32768: one FP64 tile
constexpr int I8_A_PLANE = I8_BM * I8_BK;                      // 4096
constexpr int I8_A_BYTES = I8_NDIG * I8_A_PLANE;               // 24576
constexpr int I8_B_BYTES = I8_NDIG_B * I8_BN * I8_BK;          // 14336
constexpr int I8_SMEM = I8_FSTAGES * I8_F_BYTES + I8_DSTAGES * (I8_A_BYTES + I8_B_BYTES) + 256 + 1024;
constexpr int I8_EXP_ZERO = -100000;                           // exponent of an all-zero row / column
constexpr int I8_EXP_NONFINITE = 100000;                       // a NaN / Inf was seen: the whole output row / column becomes NaN (as in FP64)

// ------------------------------------------------------------------------------------------------------------------
// digit arithmetic (barrier-free device code: exercised on the CPU by tests/test_i8_digits_emulation_cpu.py)
// ------------------------------------------------------------------------------------------------------------------
// smallest E with |x| 2^-E < 1/2 - 2^-8 (up to one binade of slack)
// running maximum of |x| that a NaN or an infinity turns into +inf for good (fmax alone would drop a NaN)
__device__ __forceinline__ double i8_amax(double amax, double x) {
    const double a = fabs(x);
    return (a <= DBL_MAX) ? fmax(amax, a) : __longlong_as_double(0x7ff0000000000000ll);
}
__device__ __forceinline__ int i8_exponent(double amax) {
    if (!(amax > 0.0)) return I8_EXP_ZERO;
    const int biased = (int)((unsigned long long)__double_as_longlong(amax) >> 52) & 0x7ff;
    if (biased == 0x7ff) return I8_EXP_NONFINITE;
    if (biased < 128) return I8_EXP_ZERO;   // < 2^-895: treated as an all-zero row (keeps 2^(48-E) representable)
    // amax in [2^(b-1023), 2^(b-1022))  ->  amax 2^-(b-1021) < 1/2.  The digits of X + 0x8080808080 need X < 2^47 - 2^39.01 (the
    // signed top byte must not reach +128): a maximum whose six leading mantissa bits are ones (amax 2^-E >= 1/2 - 2^-8) takes the
    // next exponent, so that always  X < 2^47 - 2^40  and  X + 0x8080808080 < 2^47
    const int near_top = (((unsigned long long)__double_as_longlong(amax) >> 46) & 0x3f) == 0x3f;
    return biased - 1022 + 1 + near_top;
}
// 2^(48 - E) as a double (0 for an all-zero row: every digit is then 0)
__device__ __forceinline__ double i8_scale(int E, int frac = I8_FRAC) {
    if (E == I8_EXP_ZERO || E == I8_EXP_NONFINITE) return 0.0;
    return __longlong_as_double((long long)(1023 + frac - E) << 52);
}
// X = rint(x scale), |X| < 2^(8 ND - 1) - 2^(8 ND - 8).  The bytes of Z = (X + C) ^ C, C = 0x80 in each of the ND - 1 low bytes, are
// the balanced base-256 digits as int8: low byte j of the sum is d + 128 for digit plane ND - 1 - j, flipping its bit 7 gives
// d; the top byte is the signed top digit (plane 0).  `lo` = the four lowest planes (bytes 0..3), `hi` = the remaining ones.
template <int ND>
__device__ __forceinline__ void i8_fields(double x, double scale, unsigned &lo, unsigned &hi) {
    const long long X = __double2ll_rn(x * scale);
    constexpr long long C = (ND == 6) ? 0x8080808080ll : 0x808080808080ll;
    const unsigned long long Z = (unsigned long long)(X + C) ^ (unsigned long long)C;
    lo = (unsigned)Z;
    hi = (unsigned)(Z >> 32) & ((ND == 6) ? 0xffffu : 0xffffffu);
}
// bytes of four consecutive elements -> one word per plane (byte i = element i): a 4 x 4 byte transpose with PRMT
template <int ND>
__device__ __forceinline__ void i8_pack4(const unsigned (&lo)[4], const unsigned (&hi)[4], unsigned (&out)[ND]) {
    const unsigned a = __byte_perm(lo[0], lo[1], 0x5140), b = __byte_perm(lo[2], lo[3], 0x5140);   // bytes 0 and 1 of the four lo words
    const unsigned c = __byte_perm(lo[0], lo[1], 0x7362), d = __byte_perm(lo[2], lo[3], 0x7362);   // bytes 2 and 3
    const unsigned e = __byte_perm(hi[0], hi[1], 0x5140), f = __byte_perm(hi[2], hi[3], 0x5140);   // bytes 0 and 1 of the four hi words
    out[ND - 1] = __byte_perm(a, b, 0x5410);
    out[ND - 2] = __byte_perm(a, b, 0x7632);
    out[ND - 3] = __byte_perm(c, d, 0x5410);
    out[ND - 4] = __byte_perm(c, d, 0x7632);
    out[ND - 5] = __byte_perm(e, f, 0x5410);
    out[ND - 6] = __byte_perm(e, f, 0x7632);
    if (ND == 7) {
        const unsigned g = __byte_perm(hi[0], hi[1], 0x7362), h = __byte_perm(hi[2], hi[3], 0x7362);   // byte 2 of the hi words
        out[0] = __byte_perm(g, h, 0x5410);
    }
}

// exponents of the rows of a strided matrix view: row r, reduction index j at  base[r * sr + j * sj]
// (a) sr == 1 (rows contiguous): one thread per row, the reduction range is split over gridDim.y (atomicMax on the exponent)
__global__ void i8_row_exponent_strided_kernel(const double *__restrict__ base, int64_t nrows, int64_t nred, int64_t sj, int *__restrict__ E) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int64_t chunk = (nred + gridDim.y - 1) / gridDim.y;
    const int64_t j0 = blockIdx.y * chunk, j1 = min(nred, j0 + chunk);
    double amax = 0.0;
    for (int64_t j = j0; j < j1; ++j) amax = i8_amax(amax, base[r + j * sj]);
    atomicMax(&E[r], i8_exponent(amax));
}
// (b) sj == 1 (reduction index contiguous): one warp per row
__global__ void i8_row_exponent_contig_kernel(const double *__restrict__ base, int64_t nrows, int64_t nred, int64_t sr, int *__restrict__ E) {
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= nrows) return;
    double amax = 0.0;
    for (int64_t j = lane; j < nred; j += 32) amax = i8_amax(amax, base[r * sr + j]);
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) E[r] = i8_exponent(amax);
}

// Khatri-Rao operand: K[k, r] = prod_f A_f[i_f(k), r]
struct I8Krp {
    const double *fac[ITCPD_MAX_ORDER];
    int64_t ext[ITCPD_MAX_ORDER], dim[ITCPD_MAX_ORDER];
    int nf;
    int64_t kext;
    int R;
    int r0;     // first rank column of this 64-column block
};
__device__ __forceinline__ double i8_krp_value(const I8Krp &a, int64_t k, int r) {
    r += a.r0;
    if (k >= a.kext || r >= a.R) return 0.0;
    double v = 1.0;
    int64_t rem = k;
    for (int f = 0; f < a.nf; ++f) {
        const int64_t i = rem % a.ext[f];
        rem /= a.ext[f];
        if (i >= a.dim[f]) return 0.0;     // padded row of a leading mode
        v *= a.fac[f][i + a.dim[f] * (int64_t)r];
    }
    return v;
}
// The same values for CONSECUTIVE k without a 64-bit division per element: the multi-index is decomposed once and then
// incremented with carries (same factors multiplied in the same order: bitwise the values of i8_krp_value).
struct I8KrpIter {
    int64_t i[ITCPD_MAX_ORDER];
    int64_t k;
    __device__ __forceinline__ void init(const I8Krp &a, int64_t k0) {
        k = k0;
        int64_t rem = k0;
#pragma unroll
        for (int f = 0; f < ITCPD_MAX_ORDER; ++f) {
            i[f] = 0;
            if (f < a.nf) { i[f] = rem % a.ext[f]; rem /= a.ext[f]; }
        }
    }
    __device__ __forceinline__ double value(const I8Krp &a, int r) const {
        r += a.r0;
        if (k >= a.kext || r >= a.R) return 0.0;
        double v = 1.0;
        bool pad = false;
#pragma unroll
        for (int f = 0; f < ITCPD_MAX_ORDER; ++f)
            if (f < a.nf) {
                if (i[f] >= a.dim[f]) pad = true;                           // padded row of a leading mode
                else v *= a.fac[f][i[f] + a.dim[f] * (int64_t)r];
            }
        return pad ? 0.0 : v;
    }
    __device__ __forceinline__ void next(const I8Krp &a) {
        ++k;
        bool carry = true;
#pragma unroll
        for (int f = 0; f < ITCPD_MAX_ORDER; ++f)
            if (f < a.nf && carry) {
                if (++i[f] < a.ext[f]) carry = false;
                else i[f] = 0;
            }
    }
};
// column exponents eb[r] (E must be pre-set to I8_EXP_ZERO); one thread per (chunk of 256 k, r), r fastest
__global__ void i8_krp_exponent_kernel(I8Krp a, int *__restrict__ E) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int r = (int)(idx % I8_BN);
    const int64_t k0 = (idx / I8_BN) * 256;
    if (k0 >= a.kext) return;
    double amax = 0.0;
    I8KrpIter it;
    it.init(a, k0);
    for (int64_t k = k0; k < min(a.kext, k0 + 256); ++k, it.next(a)) amax = i8_amax(amax, it.value(a, r));
    atomicMax(&E[r], i8_exponent(amax));
}
// digit planes of the Khatri-Rao operand in the canonical K-major UMMA layout, one 14336-byte block per k-tile of 32:
//   byte(q, n, kk) = ((q*64 + n) % 8) * 16 + ((q*64 + n) / 8) * 256 + (kk / 16) * 128 + (kk % 16)
// one thread per (k-tile, n, half): 16 consecutive k of one column -> one 16-byte store per plane
__global__ void i8_krp_pack_kernel(I8Krp a, const int *__restrict__ E, int64_t ktiles, uint8_t *__restrict__ out) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= ktiles * I8_BN * 2) return;
    const int half = (int)(idx & 1);
    const int n = (int)((idx >> 1) % I8_BN);
    const int64_t kt = (idx >> 1) / I8_BN;
    const double scale = i8_scale(E[n], I8_FRAC_B);
    unsigned plane[I8_NDIG_B][4];
    I8KrpIter it;
    it.init(a, kt * I8_BK + 16 * half);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        unsigned lo[4], hi[4], o[I8_NDIG_B];
#pragma unroll
        for (int j = 0; j < 4; ++j, it.next(a)) i8_fields<I8_NDIG_B>(it.value(a, n), scale, lo[j], hi[j]);
        i8_pack4<I8_NDIG_B>(lo, hi, o);
#pragma unroll
        for (int p = 0; p < I8_NDIG_B; ++p) plane[p][w] = o[p];
    }
    uint8_t *blk = out + kt * (int64_t)I8_B_BYTES;
#pragma unroll
    for (int q = 0; q < I8_NDIG_B; ++q) {
        const int row = q * I8_BN + n;
        uint4 v = make_uint4(plane[q][0], plane[q][1], plane[q][2], plane[q][3]);
        *reinterpret_cast<uint4 *>(blk + (row & 7) * 16 + (row >> 3) * 256 + half * 128) = v;
    }
}

// One converter thread's share of a 128 x 32 FP64 tile -> 6 digit planes (16 elements, one 16-byte store per plane).
//   KIND 0: F = [k (32)][m (128)] doubles; thread tid (0..255) owns row m = tid % 128 and k = 16 (tid / 128) .. +15;
//           K-major planes   byte(m, kk) = (m % 8) * 16 + (m / 8) * 256 + (kk / 16) * 128 + (kk % 16)
//   KIND 1: F = [n (128)][k (32)] doubles; thread (warp = tid / 32, lane) owns k = lane and rows n = 16 warp .. +15;
//           MN-major planes  byte(n, kk) = (kk % 8) * 16 + (kk / 8) * 128 + (n / 16) * 512 + (n % 16)
// ea_tile points at the 128 row exponents of the tile.
template <int KIND>
__device__ __forceinline__ void i8_convert_thread(const double *__restrict__ F, const int *__restrict__ ea_tile, uint8_t *__restrict__ A, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    unsigned plane[I8_NDIG][4];
    const double scale0 = (KIND == 0) ? i8_scale(ea_tile[tid & 127]) : 0.0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        unsigned lo[4], hi[4], o[I8_NDIG];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = 4 * w + jj;
            double x, sc;
            if (KIND == 0) { x = F[(16 * (tid >> 7) + j) * I8_BM + (tid & 127)]; sc = scale0; }
            else { x = F[(16 * warp + j) * I8_BK + lane]; sc = i8_scale(ea_tile[16 * warp + j]); }
            i8_fields<I8_NDIG>(x, sc, lo[jj], hi[jj]);
        }
        i8_pack4<I8_NDIG>(lo, hi, o);
#pragma unroll
        for (int p = 0; p < I8_NDIG; ++p) plane[p][w] = o[p];
    }
    int off;
    if (KIND == 0) { const int m = tid & 127; off = (m & 7) * 16 + (m >> 3) * 256 + (tid >> 7) * 128; }
    else { off = (lane & 7) * 16 + (lane >> 3) * 128 + warp * 512; }
#pragma unroll
    for (int p = 0; p < I8_NDIG; ++p)
        *reinterpret_cast<uint4 *>(A + p * I8_A_PLANE + off) = make_uint4(plane[p][0], plane[p][1], plane[p][2], plane[p][3]);
}

// C = 2^(ea + eb - 104 + 88) sum_t acc_t 2^(-8 t):  v = sum_t acc_t 2^(-8 t) is formed by the caller, smallest weights first
// exact int32 -> double without the (quarter-rate) I2F.F64 conversion: the double whose high word is 0x43300000 and whose low
// word is a + 2^31 equals 2^52 + 2^31 + a; one DADD removes the offset exactly
__device__ __forceinline__ double i8_i2d(int a) {
    return __hiloint2double(0x43300000, (int)((unsigned)a ^ 0x80000000u)) - 4503601774854144.0;
}
__device__ __forceinline__ double i8_weight(int t) { return __longlong_as_double((long long)(1023 - 8 * t) << 52); }
__device__ __forceinline__ double i8_finish(double v, int em, int er) {
    if (em == I8_EXP_NONFINITE || er == I8_EXP_NONFINITE) return __longlong_as_double(0x7ff8000000000000ll);   // NaN, like the FP64 contraction
    if (em == I8_EXP_ZERO || er == I8_EXP_ZERO) return 0.0;
    // v 2^e as two exact power-of-two multiplications (e = em + er - 16 lies in [-1802, 2036]: each half is a normal double);
    // ldexp() is a ~20-instruction library routine, and this runs once per output element
    const int e = em + er - I8_FRAC - I8_FRAC_B + 8 * (I8_NDIG - 1 + I8_NDIG_B - 1), h = e >> 1;   // = em + er - 16
    return v * __longlong_as_double((long long)(1023 + h) << 52) * __longlong_as_double((long long)(1023 + e - h) << 52);
}

#ifndef ITCPD_I8_HOST_EMULATION
// ------------------------------------------------------------------------------------------------------------------
// PTX wrappers (tcgen05 / TMEM / mbarrier / TMA)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t i8_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void i8_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void i8_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void i8_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded while the kernel is a draft: a protocol error must surface as a trapped launch, not as a hung GPU
__device__ __forceinline__ void i8_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    unsigned long long spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1ull << 26)) {
            // bar & 1023 names the barrier (both kernels put the barrier block at a multiple of 1024 bytes):
            //   partial_gemm_i8_kernel : full_f 0/8/16, empty_f 24/32/40, full_d 48/56, empty_d 64/72, acc_full 80, acc_empty 88
            //   partial_gemm_i8p_kernel: full 0/8/16/24, empty 32/40/48/56, acc_full 64, acc_empty 72
            printf("itcpd gemm_i8: mbarrier +%u parity %u never completed (block %d thread %d)\n", bar & 1023u, parity, (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    } while (!done);
}
// one non-blocking poll of a phase (no hardware suspend): for a thread that serves two rings at once
__device__ __forceinline__ bool i8_mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void i8_tma_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void i8_bulk_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void i8_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void i8_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void i8_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void i8_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void i8_tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void i8_tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void i8_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void i8_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 consecutive int32 columns of this thread's TMEM lane.  The load is asynchronous: the registers are defined only after
// tcgen05.wait::ld, so the wait takes them as in/out operands -- the compiler then cannot schedule a consumer above it -- and
// issue / wait are separate calls so that the next load can be in flight while the previous buffer is being consumed.
#define I8_R32(v) "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]),      \
                  "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]),      \
                  "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
__device__ __forceinline__ void i8_tmem_ld32_issue(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void i8_tmem_ld_wait(int (&v)[32]) { asm volatile("tcgen05.wait::ld.sync.aligned;" : I8_R32(v)::"memory"); }
__device__ __forceinline__ void i8_tmem_ld32(uint32_t taddr, int (&v)[32]) {
    i8_tmem_ld32_issue(taddr, v);
    i8_tmem_ld_wait(v);
}

// shared-memory matrix descriptor, SWIZZLE_NONE ("interleave"): start address, leading / stride byte offsets in 16-byte
// units, descriptor version 1 (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__host__ __device__ __forceinline__ uint64_t i8_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) |
           (1ull << 46);
}
// instruction descriptor (UMMA::InstrDescriptor): D = S32, A = B = signed int8, M = 128, N = n, B K-major, A K- or MN-major
__host__ __device__ constexpr uint32_t i8_idesc(int n, int a_mn_major) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// Work unit of the persistent kernels: (row tile, k-chunk).  ksplit = 1 is the plain data-parallel schedule; a short-and-wide
// contraction (fewer row tiles than SMs) is cut along k into ksplit chunks of kchunk k-tiles, every unit writes its scaled
// FP64 partial tile to a slot buffer and i8_splitk_fixup_kernel adds the chunks in ascending-k order (deterministic).
// All rank blocks of 64 columns ride in ONE launch, the rank block fastest in the unit index: neighbouring CTAs then stream the
// same digit planes of T for different rank blocks at the same time, so the second reader is served by the L2 and a rank above
// 64 does not multiply the HBM traffic of the pass.
struct I8Sched {
    int num_row_tiles, kt_count, ksplit, kchunk, rblocks;
    long long bdig_rb_stride;     // bytes between the packed Khatri-Rao digit planes of consecutive rank blocks
};
struct I8Unit { int tile, chunk, rb, kb, ke; };
__device__ __forceinline__ I8Unit i8_unit(int u, const I8Sched &sc) {
    I8Unit x;
    const int v = u / sc.rblocks;
    x.rb = u - v * sc.rblocks;
    x.tile = v / sc.ksplit;
    x.chunk = v - x.tile * sc.ksplit;
    x.kb = x.chunk * sc.kchunk;
    x.ke = min(sc.kt_count, x.kb + sc.kchunk);
    return x;
}
__device__ __forceinline__ int i8_units(const I8Sched &sc) { return sc.num_row_tiles * sc.ksplit * sc.rblocks; }
// position of a pipeline role inside the CTA's sequence of (unit, k-tile) steps
struct I8Cursor {
    int w, kt;
    I8Unit u;
    __device__ __forceinline__ void start(int cta, const I8Sched &sc) {
        w = 0;
        u = i8_unit(cta, sc);
        kt = u.kb;
    }
    __device__ __forceinline__ void next(int cta, int G, int my_units, const I8Sched &sc) {
        if (++kt >= u.ke) {
            ++w;
            if (w < my_units) { u = i8_unit(cta + w * G, sc); kt = u.kb; }
        }
    }
};
// part = [chunk][rank block][rows_out x 64]; out = rows_out x R (column-major)
__global__ void i8_splitk_fixup_kernel(const double *__restrict__ part, int ksplit, int rblocks, int64_t rows_out, int64_t n, double *__restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t r = i / rows_out, m = i - r * rows_out;
    const int64_t blk = rows_out * (int64_t)I8_BN;
    const double *p = part + (r / I8_BN) * blk + m + rows_out * (r % I8_BN);
    double v = p[0];
    for (int c = 1; c < ksplit; ++c) v += p[(int64_t)c * rblocks * blk];
    out[i] = v;
}

// Epilogue of one 128 x 64 tile by one warp: TMEM lane quarter q (= warp index mod 4, the hardware's rule for tcgen05.ld)
// holds rows row0 + 32 q + lane.  v = sum_t acc_t 2^(-8 t) is formed smallest weights first, scaled, stored column-major.
__device__ __forceinline__ void i8_epilogue_warp(uint32_t tmem, uint32_t acc_full, uint32_t acc_empty, int tile_seq, int64_t row0, int q, int lane,
                                                 const int *__restrict__ ea, const int *__restrict__ eb, double *__restrict__ out,
                                                 int64_t rows_out, int R) {
    i8_mbar_wait(acc_full, (uint32_t)(tile_seq & 1));
    i8_tc_fence_after();
    const int64_t m = row0 + 32 * q + lane;
    const int em = (m < rows_out) ? ea[m] : I8_EXP_ZERO;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        double v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = 0.0;
        // software pipeline over the 7 accumulators (smallest weight first): the load of accumulator t - 1 is in flight while
        // accumulator t is folded in; fully unrolled so that both buffers stay in registers
        const uint32_t tbase = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * half);
        int a[2][32];
        i8_tmem_ld32(tbase + (uint32_t)(I8_BN * (I8_NACC - 1)), a[0]);
#pragma unroll
        for (int i = 0; i < I8_NACC; ++i) {
            const int t = I8_NACC - 1 - i;
            if (t > 0) i8_tmem_ld32_issue(tbase + (uint32_t)(I8_BN * (t - 1)), a[(i + 1) & 1]);
            const double wt = i8_weight(t);
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = fma(i8_i2d(a[i & 1][c]), wt, v[c]);
            if (t > 0) i8_tmem_ld_wait(a[(i + 1) & 1]);
        }
        if (m < rows_out) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int r = 32 * half + c;
                if (r < R) out[m + rows_out * (int64_t)r] = i8_finish(v[c], em, eb[r]);
            }
        }
    }
    i8_tc_fence_before();
    i8_mbar_arrive(acc_empty);
}

// one A digit plane  x  the stacked B planes q0 .. q0 + nq - 1  ->  accumulators t0 .. t0 + nq - 1 (split at 256 columns)
__device__ __forceinline__ void i8_mma_cols(uint32_t tmem, uint64_t adesc, uint32_t b0, int a_mn_major, int t0, int q0, int nq, uint32_t acc) {
    const int ncols = I8_BN * nq;
    const uint32_t d = tmem + (uint32_t)(I8_BN * t0);
    const uint32_t brow = (uint32_t)(I8_BN * q0);                                    // first B row: 8-row groups are 256 bytes apart
    const uint64_t bdesc = i8_smem_desc(b0 + brow / 8 * 256, 128, 256);
    if (ncols > 256) {
        i8_mma(d, adesc, bdesc, i8_idesc(256, a_mn_major), acc);
        i8_mma(d + 256, adesc, i8_smem_desc(b0 + (brow + 256) / 8 * 256, 128, 256), i8_idesc(ncols - 256, a_mn_major), acc);
    } else {
        i8_mma(d, adesc, bdesc, i8_idesc(ncols, a_mn_major), acc);
    }
}
// the 9 tcgen05.mma of one k-step: A digit plane p (128 x 32)  x  B planes 0 .. 6 - p  ->  accumulators p .. 6 (the 27 digit pairs
// with p + q <= 6).  Plane 0 touches every accumulator first, so it alone overwrites on the first k-step of a unit.
template <int KIND>
__device__ __forceinline__ void i8_issue_kstep(uint32_t tmem, uint32_t a0, uint32_t b0, bool first_kstep) {
#pragma unroll
    for (int p = 0; p < I8_NDIG; ++p) {
        const uint64_t adesc = (KIND == 0) ? i8_smem_desc(a0 + p * I8_A_PLANE, 128, 256)    // K-major: LBO = k chunk, SBO = 8-row group
                                           : i8_smem_desc(a0 + p * I8_A_PLANE, 128, 512);   // MN-major: LBO = k group of 8, SBO = 16-row block
        i8_mma_cols(tmem, adesc, b0, KIND, p, 0, I8_NACC - p, (first_kstep && p == 0) ? 0u : 1u);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------------
// KIND 0: T tile lands as [k (32)][m (128)] doubles (m contiguous in memory); converter thread = one row m, 16 k:
//         K-major digit planes  byte(m, kk) = (m % 8) * 16 + (m / 8) * 256 + (kk / 16) * 128 + (kk % 16)
// KIND 1: T tile lands as [n (128)][k (32)] doubles (k contiguous in memory); converter thread = one k, 16 rows n:
//         MN-major digit planes byte(n, kk) = (kk % 8) * 16 + (kk / 8) * 128 + (n / 16) * 512 + (n % 16)
template <int KIND>
__global__ void __launch_bounds__(320, 1)
partial_gemm_i8_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t *__restrict__ Bdig, const int *__restrict__ ea,
                       const int *__restrict__ eb, double *__restrict__ out, int64_t rows_out, int R, const I8Sched sc, double *__restrict__ part) {
    extern __shared__ uint8_t i8_smem_raw[];
    const uint32_t base = (i8_smem_u32(i8_smem_raw) + 1023u) & ~1023u;
    const uint32_t sF = base;
    const uint32_t sA = sF + I8_FSTAGES * I8_F_BYTES;
    const uint32_t sB = sA + I8_DSTAGES * I8_A_BYTES;
    const uint32_t bars = sB + I8_DSTAGES * I8_B_BYTES;
    const uint32_t full_f = bars, empty_f = bars + 8 * I8_FSTAGES;                           // FP64 ring
    const uint32_t full_d = empty_f + 8 * I8_FSTAGES, empty_d = full_d + 8 * I8_DSTAGES;    // digit ring
    const uint32_t acc_full = empty_d + 8 * I8_DSTAGES, acc_empty = acc_full + 8;
    const uint32_t tmem_slot = acc_empty + 8;
    uint8_t *gen_base = i8_smem_raw + (base - i8_smem_u32(i8_smem_raw));                    // generic pointer to `base`

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < I8_FSTAGES; ++s) { i8_mbar_init(full_f + 8 * s, 1); i8_mbar_init(empty_f + 8 * s, 256); }
        for (int s = 0; s < I8_DSTAGES; ++s) { i8_mbar_init(full_d + 8 * s, 256 + 1); i8_mbar_init(empty_d + 8 * s, 1); }
        i8_mbar_init(acc_full, 1);
        i8_mbar_init(acc_empty, 128);
        i8_fence_mbar_init();
    }
    if (warp == 9) i8_tmem_alloc(tmem_slot, 512);
    i8_tc_fence_before();
    __syncthreads();
    i8_tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(gen_base + (tmem_slot - base));

    const int G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int my_tiles = (i8_units(sc) - cta + G - 1) / G;   // work units of this CTA (see I8Unit)
    const int64_t part_stride = rows_out * (int64_t)I8_BN;

    if (warp == 8) {
        // ===== TMA producer (one thread serving two independent rings with non-blocking polls: the 3-deep FP64 prefetch of T
        // must not wait for the 2-deep digit ring, and two spinning lanes of one warp would lean on intra-warp fairness) =====
        if (lane == 0 && my_tiles > 0) {
            I8Cursor cf, cd;
            cf.start(cta, sc);
            cd.start(cta, sc);
            int itf = 0, itd = 0;
            unsigned long long idle = 0;
            while (cf.w < my_tiles || cd.w < my_tiles) {
                bool progressed = false;
                if (cf.w < my_tiles) {
                    const int sf = itf % I8_FSTAGES;
                    if (itf < I8_FSTAGES || i8_mbar_test(empty_f + 8 * sf, (uint32_t)((itf / I8_FSTAGES - 1) & 1))) {
                        const int row0 = cf.u.tile * I8_BM;
                        i8_mbar_expect_tx(full_f + 8 * sf, I8_F_BYTES);
                        if (KIND == 0) i8_tma_2d(sF + sf * I8_F_BYTES, &tmap, row0, cf.kt * I8_BK, full_f + 8 * sf);
                        else i8_tma_2d(sF + sf * I8_F_BYTES, &tmap, cf.kt * I8_BK, row0, full_f + 8 * sf);
                        ++itf;
                        cf.next(cta, G, my_tiles, sc);
                        progressed = true;
                    }
                }
                if (cd.w < my_tiles) {
                    const int sd = itd % I8_DSTAGES;
                    if (itd < I8_DSTAGES || i8_mbar_test(empty_d + 8 * sd, (uint32_t)((itd / I8_DSTAGES - 1) & 1))) {
                        i8_mbar_expect_tx(full_d + 8 * sd, I8_B_BYTES);
                        i8_bulk_1d(sB + sd * I8_B_BYTES, Bdig + (size_t)cd.u.rb * sc.bdig_rb_stride + (size_t)cd.kt * I8_B_BYTES, I8_B_BYTES, full_d + 8 * sd);
                        ++itd;
                        cd.next(cta, G, my_tiles, sc);
                        progressed = true;
                    }
                }
                if (progressed) idle = 0;
                else if (++idle > (1ull << 24)) {
                    printf("itcpd gemm_i8: producer stalled (block %d, F step %d, digit step %d)\n", (int)blockIdx.x, itf, itd);
                    __trap();
                } else {
                    __nanosleep(64);   // both rings are full: do not spin hot on the scheduler this warp shares with two converter warps
                }
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            int it = 0;
            for (int w = 0; w < my_tiles; ++w) {
                const I8Unit u = i8_unit(cta + w * G, sc);
                if (w > 0) i8_mbar_wait(acc_empty, (uint32_t)((w - 1) & 1));   // the epilogue has drained the accumulators
                i8_tc_fence_after();
                for (int kt = u.kb; kt < u.ke; ++kt, ++it) {
                    const int sd = it % I8_DSTAGES;
                    i8_mbar_wait(full_d + 8 * sd, (uint32_t)((it / I8_DSTAGES) & 1));
                    i8_tc_fence_after();
                    i8_issue_kstep<KIND>(tmem, sA + sd * I8_A_BYTES, sB + sd * I8_B_BYTES, kt == u.kb);
                    i8_commit(empty_d + 8 * sd);                                // frees the digit slot when these MMAs are done
                }
                i8_commit(acc_full);                                            // accumulators of this tile are complete
            }
        }
    } else {
        // ===== converters (warps 0-7), epilogue (warps 0-3) =====
        const int tid = threadIdx.x;   // 0 .. 255
        int it = 0;
        for (int w = 0; w < my_tiles; ++w) {
            const I8Unit u = i8_unit(cta + w * G, sc);
            const int64_t row0 = (int64_t)u.tile * I8_BM;
            for (int kt = u.kb; kt < u.ke; ++kt, ++it) {
                const int sf = it % I8_FSTAGES, sd = it % I8_DSTAGES;
                i8_mbar_wait(full_f + 8 * sf, (uint32_t)((it / I8_FSTAGES) & 1));
                if (it >= I8_DSTAGES) i8_mbar_wait(empty_d + 8 * sd, (uint32_t)((it / I8_DSTAGES - 1) & 1));
                const double *F = reinterpret_cast<const double *>(gen_base + (sF - base) + sf * I8_F_BYTES);
                uint8_t *A = gen_base + (sA - base) + sd * I8_A_BYTES;
                i8_convert_thread<KIND>(F, ea + row0, A, tid);
                i8_fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core
                i8_mbar_arrive(full_d + 8 * sd);
                i8_mbar_arrive(empty_f + 8 * sf);
            }
            if (warp < 4) {
                i8_epilogue_warp(tmem, acc_full, acc_empty, w, row0, warp, lane, ea, eb + u.rb * I8_BN,
                                 sc.ksplit > 1 ? part + (int64_t)(u.chunk * sc.rblocks + u.rb) * part_stride : out + (int64_t)u.rb * part_stride, rows_out,
                                 min(I8_BN, R - u.rb * I8_BN));
            }
        }
    }
    i8_tc_fence_before();
    __syncthreads();
    if (warp == 9) i8_tmem_dealloc(tmem, 512);
}


// ------------------------------------------------------------------------------------------------------------------
// Pre-packed variant (gemm_i8 = 2): T never changes during a decomposition, so its digit planes are computed ONCE per
// unfolding and kept in HBM as one 24576-byte block per (row tile, k-tile), already in the canonical UMMA layout.
// A pass then streams 6 bytes per tensor element instead of 8, needs no conversion in the loop, and the kernel is a
// plain TMA -> tcgen05.mma -> TMEM pipeline: warp 0 producer, warp 1 MMA issuer (+ TMEM allocation), warps 2-5 epilogue.
// ------------------------------------------------------------------------------------------------------------------
constexpr int I8P_STAGES = 4;
constexpr int I8P_STAGE_BYTES = I8_A_BYTES + I8_B_BYTES;      // 38912
constexpr int I8P_SMEM = I8P_STAGES * I8P_STAGE_BYTES + 256 + 1024;

// one CTA per (row tile, k-tile): stage the 128 x 32 FP64 tile in shared memory (zero fill outside the tensor), then the
// converter threads write the digit block.  view: element (row r, contraction index j) at T[r * sr + j * sj]
template <int KIND>
__global__ void __launch_bounds__(256) i8_pack_tensor_kernel(const double *__restrict__ T, int64_t nrows, int64_t nred, int64_t sr, int64_t sj,
                                                             const int *__restrict__ ea, int64_t ktiles, uint8_t *__restrict__ Adig) {
    __shared__ double F[I8_BM * I8_BK];
    const int64_t blk = blockIdx.x;
    const int64_t tile = blk / ktiles, kt = blk - tile * ktiles;
    const int64_t row0 = tile * I8_BM, k0 = kt * I8_BK;
    for (int e = threadIdx.x; e < I8_BM * I8_BK; e += 256) {
        // KIND 0 staging order [k][m] (m contiguous in memory), KIND 1 [n][k] (k contiguous in memory)
        const int rr = (KIND == 0) ? e % I8_BM : e / I8_BK, kk = (KIND == 0) ? e / I8_BM : e % I8_BK;
        const int64_t r = row0 + rr, j = k0 + kk;
        F[e] = (r < nrows && j < nred) ? T[r * sr + j * sj] : 0.0;
    }
    __syncthreads();
    i8_convert_thread<KIND>(F, ea + row0, Adig + blk * (int64_t)I8_A_BYTES, (int)threadIdx.x);
}

template <int KIND>
__global__ void __launch_bounds__(192, 1)
partial_gemm_i8p_kernel(const uint8_t *__restrict__ Adig, const uint8_t *__restrict__ Bdig, const int *__restrict__ ea, const int *__restrict__ eb,
                        double *__restrict__ out, int64_t rows_out, int R, const I8Sched sc, double *__restrict__ part) {
    extern __shared__ uint8_t i8_smem_raw[];
    const uint32_t base = (i8_smem_u32(i8_smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + I8P_STAGES * I8P_STAGE_BYTES;
    const uint32_t full = bars, empty = bars + 8 * I8P_STAGES, acc_full = empty + 8 * I8P_STAGES, acc_empty = acc_full + 8, tmem_slot = acc_empty + 8;
    uint8_t *gen_base = i8_smem_raw + (base - i8_smem_u32(i8_smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < I8P_STAGES; ++s) { i8_mbar_init(full + 8 * s, 1); i8_mbar_init(empty + 8 * s, 1); }
        i8_mbar_init(acc_full, 1);
        i8_mbar_init(acc_empty, 128);
        i8_fence_mbar_init();
    }
    if (warp == 1) i8_tmem_alloc(tmem_slot, 512);
    i8_tc_fence_before();
    __syncthreads();
    i8_tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(gen_base + (tmem_slot - base));
    const int G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int my_tiles = (i8_units(sc) - cta + G - 1) / G;   // work units of this CTA (see I8Unit)
    const int64_t part_stride = rows_out * (int64_t)I8_BN;
    const int kt_count = sc.kt_count;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int w = 0; w < my_tiles; ++w) {
                const I8Unit u = i8_unit(cta + w * G, sc);
                const int64_t tile = u.tile;
                for (int kt = u.kb; kt < u.ke; ++kt, ++it) {
                    const int st = it % I8P_STAGES;
                    if (it >= I8P_STAGES) i8_mbar_wait(empty + 8 * st, (uint32_t)((it / I8P_STAGES - 1) & 1));
                    i8_mbar_expect_tx(full + 8 * st, I8P_STAGE_BYTES);
                    const uint32_t dst = base + st * I8P_STAGE_BYTES;
                    i8_bulk_1d(dst, Adig + (size_t)(tile * kt_count + kt) * I8_A_BYTES, I8_A_BYTES, full + 8 * st);
                    i8_bulk_1d(dst + I8_A_BYTES, Bdig + (size_t)u.rb * sc.bdig_rb_stride + (size_t)kt * I8_B_BYTES, I8_B_BYTES, full + 8 * st);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int it = 0;
            for (int w = 0; w < my_tiles; ++w) {
                const I8Unit u = i8_unit(cta + w * G, sc);
                if (w > 0) i8_mbar_wait(acc_empty, (uint32_t)((w - 1) & 1));
                i8_tc_fence_after();
                for (int kt = u.kb; kt < u.ke; ++kt, ++it) {
                    const int st = it % I8P_STAGES;
                    i8_mbar_wait(full + 8 * st, (uint32_t)((it / I8P_STAGES) & 1));
                    i8_tc_fence_after();
                    const uint32_t a0 = base + st * I8P_STAGE_BYTES;
                    i8_issue_kstep<KIND>(tmem, a0, a0 + I8_A_BYTES, kt == u.kb);
                    i8_commit(empty + 8 * st);
                }
                i8_commit(acc_full);
            }
        }
    } else {
        for (int w = 0; w < my_tiles; ++w) {
            const I8Unit u = i8_unit(cta + w * G, sc);
            i8_epilogue_warp(tmem, acc_full, acc_empty, w, (int64_t)u.tile * I8_BM, warp & 3, lane, ea, eb + u.rb * I8_BN,
                             sc.ksplit > 1 ? part + (int64_t)(u.chunk * sc.rblocks + u.rb) * part_stride : out + (int64_t)u.rb * part_stride, rows_out,
                             min(I8_BN, R - u.rb * I8_BN));
        }
    }
    i8_tc_fence_before();
    __syncthreads();
    if (warp == 1) i8_tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
#ifndef ITCPD_I8_PROBE   // tools/i8_probe.cu includes this file for the device code only
typedef CUresult (*I8EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                               CUtensorMapFloatOOBfill);

static int i8_make_tmap(CUtensorMap *map, const double *base, uint64_t d0, uint64_t d1, uint32_t box0, uint32_t box1) {
    static I8EncodeFn enc = nullptr;
    if (!enc) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return ITCPD_ERR_CUDA;
        enc = reinterpret_cast<I8EncodeFn>(p);
    }
    cuuint64_t gdim[2] = {d0, d1};
    cuuint64_t gstr[1] = {d0 * 8};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ITCPD_OK : ITCPD_ERR_CUDA;
}

// picks (ksplit, kchunk): a chunk never exceeds I8_MAX_KCHUNK k-tiles (the int32 accumulators stay exact); beyond that,
// maximise the SM occupancy of the last wave and prefer the smallest split among equals
void i8_choose_ksplit(int64_t row_tiles, int64_t ktiles, int sms, int *ksplit_out, int *kchunk_out) {
    const int64_t min_split = ceil_div(ktiles, (int64_t)I8_MAX_KCHUNK);
    int64_t best = min_split;
    double best_eff = 0.0;
    const int64_t max_split = std::max<int64_t>(min_split, std::min<int64_t>(ktiles / 8, 4 * (int64_t)sms));
    if (row_tiles < 4 * (int64_t)sms) {
        for (int64_t ks = min_split; ks <= max_split; ++ks) {
            const int64_t chunk = ceil_div(ktiles, ks), real = ceil_div(ktiles, chunk);   // no empty chunk
            if (real != ks) continue;
            const int64_t units = row_tiles * ks, waves = ceil_div(units, (int64_t)sms);
            // time ~ waves * (chunk + fill); efficiency relative to the ideal  row_tiles * ktiles / sms
            const double eff = (double)(row_tiles * ktiles) / ((double)sms * (double)waves * (double)(chunk + 2));
            if (eff > best_eff * 1.02) { best_eff = eff; best = ks; }
        }
    }
    const int64_t chunk = ceil_div(ktiles, best);
    *kchunk_out = (int)chunk;
    *ksplit_out = (int)ceil_div(ktiles, chunk);     // == best unless min_split itself left an empty chunk
}

__global__ void i8_fill_int_kernel(int *x, int64_t n, int v) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}

// Same contract as launch_partial_gemm (gemm_dmma.cu).  Ragged shapes ride on TMA's zero fill and zero digit planes; rank
// columns are processed in blocks of 64 (one pass over T per block: 7 x 64 accumulator columns fill the TMEM).
// Returns ITCPD_ERR_UNSUPPORTED outside the envelope (the caller then uses the DMMA kernel).
int launch_partial_gemm_i8(itcpd_ctx *c, int kind, int split, double *out) {
    const int N = c->order, R = c->rank;
    int64_t Mrows = c->ld0, Ncols = 1;
    for (int n = 1; n < split; ++n) Mrows *= c->dims[n];
    for (int n = split; n < N; ++n) Ncols *= c->dims[n];
    const int64_t kext = (kind == 0) ? Ncols : Mrows;
    const int64_t rows_out = (kind == 0) ? Mrows : Ncols;
    const int64_t ktiles = ceil_div(kext, I8_BK), row_tiles = ceil_div(rows_out, I8_BM);
    if (row_tiles > INT32_MAX / I8_BM || ktiles > INT32_MAX / I8_BK || Mrows > INT32_MAX || Ncols > INT32_MAX) return ITCPD_ERR_UNSUPPORTED;
    const int rblocks = (int)ceil_div(R, I8_BN);
    // split-K schedule (I8Unit): data parallel when there are enough row tiles; otherwise cut k so that the units fill the
    // SMs in whole waves (a chunk keeps at least 8 k-tiles so the pipeline fill and the partial-tile traffic stay small)
    int ksplit = 1, kchunk = (int)ktiles;
    const int sms = std::max(1, c->sm_count - c->i8_spare_sms);   // SMs this persistent kernel may occupy
    i8_choose_ksplit(row_tiles * rblocks, ktiles, sms, &ksplit, &kchunk);   // rank blocks are independent units too

    // ---- row exponents of this unfolding of T: computed once per tensor and split, cached in the handle ----
    I8ExpCache &ec = c->i8_exp[kind];
    if (!ec.valid || ec.split != split || ec.tensor_epoch != c->i8_tensor_epoch) {
        const int64_t padded = row_tiles * I8_BM;
        TRY(ec.buf.reserve((size_t)padded * 4));
        i8_fill_int_kernel<<<(unsigned)ceil_div(padded, 256), 256, 0, c->stream>>>(ec.buf.as<int>(), padded, I8_EXP_ZERO);
        if (kind == 0) {
            const int ysplit = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(kext, 65535),   // gridDim.y limit
                                                                           (int64_t)c->sm_count * 8 * 256 / std::max<int64_t>(rows_out, 1)));
            i8_row_exponent_strided_kernel<<<dim3((unsigned)ceil_div(rows_out, 256), (unsigned)ysplit), 256, 0, c->stream>>>(
                c->T.as<double>(), rows_out, kext, Mrows, ec.buf.as<int>());
        } else {
            i8_row_exponent_contig_kernel<<<(unsigned)ceil_div(rows_out * 32, 256), 256, 0, c->stream>>>(c->T.as<double>(), rows_out, kext, Mrows,
                                                                                                         ec.buf.as<int>());
        }
        c->launches += 2;
        CUDA_TRY(cudaGetLastError());
        ec.valid = true;
        ec.split = split;
        ec.tensor_epoch = c->i8_tensor_epoch;
    }

    CUtensorMap map;
    if (kind == 0) TRY(i8_make_tmap(&map, c->T.as<double>(), (uint64_t)Mrows, (uint64_t)Ncols, I8_BM, I8_BK));
    else TRY(i8_make_tmap(&map, c->T.as<double>(), (uint64_t)Mrows, (uint64_t)Ncols, I8_BK, I8_BM));
    static bool attr[2][64] = {{false}};
    if (!attr[kind][c->device & 63]) {
        if (kind == 0) CUDA_TRY(cudaFuncSetAttribute(partial_gemm_i8_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
        else CUDA_TRY(cudaFuncSetAttribute(partial_gemm_i8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
        attr[kind][c->device & 63] = true;
    }

    // ---- pre-packed digit planes of T (gemm_i8 = 2): built once per tensor / unfolding, 6 bytes per element in HBM ----
    bool prepacked = c->gemm_i8 == 2;
    I8ExpCache &pk = c->i8_apack[kind];
    if (prepacked && pk.nofit_epoch == c->i8_tensor_epoch && pk.nofit_split == split) prepacked = false;
    if (prepacked && (!pk.valid || pk.split != split || pk.tensor_epoch != c->i8_tensor_epoch)) {
        if (pk.buf.reserve((size_t)row_tiles * ktiles * I8_A_BYTES) != ITCPD_OK) {
            cudaGetLastError();                 // not enough HBM for the digit planes (e.g. 2048^3 on one GPU): convert on the fly instead
            pk.valid = false;
            pk.nofit_epoch = c->i8_tensor_epoch;
            pk.nofit_split = split;
            prepacked = false;
        }
    }
    if (prepacked && (!pk.valid || pk.split != split || pk.tensor_epoch != c->i8_tensor_epoch)) {
        if (kind == 0)
            i8_pack_tensor_kernel<0><<<(unsigned)(row_tiles * ktiles), 256, 0, c->stream>>>(c->T.as<double>(), rows_out, kext, 1, Mrows, ec.buf.as<int>(), ktiles,
                                                                                             pk.buf.as<uint8_t>());
        else
            i8_pack_tensor_kernel<1><<<(unsigned)(row_tiles * ktiles), 256, 0, c->stream>>>(c->T.as<double>(), rows_out, kext, Mrows, 1, ec.buf.as<int>(), ktiles,
                                                                                             pk.buf.as<uint8_t>());
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        pk.valid = true;
        pk.split = split;
        pk.tensor_epoch = c->i8_tensor_epoch;
        static bool attr_p[2][64] = {{false}};
        if (!attr_p[kind][c->device & 63]) {
            if (kind == 0) CUDA_TRY(cudaFuncSetAttribute(partial_gemm_i8p_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8P_SMEM));
            else CUDA_TRY(cudaFuncSetAttribute(partial_gemm_i8p_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8P_SMEM));
            attr_p[kind][c->device & 63] = true;
        }
    }

    // ---- Khatri-Rao operand: column exponents and packed digit planes of every 64-column block ----
    I8Krp pa;
    memset(&pa, 0, sizeof(pa));
    pa.R = R;
    pa.kext = kext;
    if (kind == 0) {
        for (int n = split; n < N; ++n) { pa.fac[pa.nf] = c->A[n].as<double>(); pa.ext[pa.nf] = c->dims[n]; pa.dim[pa.nf] = c->dims[n]; pa.nf++; }
    } else {
        for (int n = 0; n < split; ++n) {
            pa.fac[pa.nf] = c->A[n].as<double>(); pa.ext[pa.nf] = (n == 0) ? c->ld0 : c->dims[n]; pa.dim[pa.nf] = c->dims[n]; pa.nf++;
        }
    }
    TRY(c->i8_eb.reserve((size_t)rblocks * I8_BN * 4));
    TRY(c->i8_bdig.reserve((size_t)rblocks * ktiles * I8_B_BYTES));
    i8_fill_int_kernel<<<(unsigned)ceil_div(rblocks * I8_BN, 64), 64, 0, c->stream>>>(c->i8_eb.as<int>(), rblocks * I8_BN, I8_EXP_ZERO);
    c->launches++;
    // one launch covers every rank block (I8Sched): the grid is a multiple of rblocks so that CTA c always serves rank block c % rblocks
    const int64_t units = row_tiles * ksplit * rblocks;
    const int grid = (int)std::min<int64_t>(units, std::max(rblocks, sms / rblocks * rblocks));
    const int64_t part_stride = rows_out * (int64_t)I8_BN;
    if (ksplit > 1) TRY(c->i8_part.reserve((size_t)ksplit * rblocks * part_stride * 8));
    double *part = ksplit > 1 ? c->i8_part.as<double>() : nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->time_gemm) {  // same bookkeeping as launch_partial_gemm: one event pair per contraction (itcpd_gemm_timing)
        if (c->gemm_events_used == c->gemm_events.size()) {
            cudaEvent_t a, b;
            CUDA_TRY(cudaEventCreate(&a));
            CUDA_TRY(cudaEventCreate(&b));
            c->gemm_events.push_back({a, b});
        }
        e0 = c->gemm_events[c->gemm_events_used].first;
        e1 = c->gemm_events[c->gemm_events_used].second;
        c->gemm_events_used++;
        CUDA_TRY(cudaEventRecord(e0, c->stream));
    }
    for (int rb = 0; rb < rblocks; ++rb) {
        pa.r0 = rb * I8_BN;
        int *eb = c->i8_eb.as<int>() + rb * I8_BN;
        uint8_t *bdig = c->i8_bdig.as<uint8_t>() + (size_t)rb * ktiles * I8_B_BYTES;
        i8_krp_exponent_kernel<<<(unsigned)ceil_div(ceil_div(kext, 256) * I8_BN, 256), 256, 0, c->stream>>>(pa, eb);
        i8_krp_pack_kernel<<<(unsigned)ceil_div(ktiles * I8_BN * 2, 256), 256, 0, c->stream>>>(pa, eb, ktiles, bdig);
        c->launches += 2;
    }
    I8Sched sc;
    sc.num_row_tiles = (int)row_tiles;
    sc.kt_count = (int)ktiles;
    sc.ksplit = ksplit;
    sc.kchunk = kchunk;
    sc.rblocks = rblocks;
    sc.bdig_rb_stride = (long long)ktiles * I8_B_BYTES;
    const uint8_t *bdig = c->i8_bdig.as<uint8_t>();
    const int *eb = c->i8_eb.as<int>();
    if (prepacked && kind == 0) partial_gemm_i8p_kernel<0><<<grid, 192, I8P_SMEM, c->stream>>>(pk.buf.as<uint8_t>(), bdig, ec.buf.as<int>(), eb, out, rows_out, R, sc, part);
    else if (prepacked) partial_gemm_i8p_kernel<1><<<grid, 192, I8P_SMEM, c->stream>>>(pk.buf.as<uint8_t>(), bdig, ec.buf.as<int>(), eb, out, rows_out, R, sc, part);
    else if (kind == 0) partial_gemm_i8_kernel<0><<<grid, 320, I8_SMEM, c->stream>>>(map, bdig, ec.buf.as<int>(), eb, out, rows_out, R, sc, part);
    else partial_gemm_i8_kernel<1><<<grid, 320, I8_SMEM, c->stream>>>(map, bdig, ec.buf.as<int>(), eb, out, rows_out, R, sc, part);
    c->launches++;
    if (ksplit > 1) {
        const int64_t n = rows_out * (int64_t)R;
        i8_splitk_fixup_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, c->stream>>>(part, ksplit, rblocks, rows_out, n, out);
        c->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    if (e1) CUDA_TRY(cudaEventRecord(e1, c->stream));
    return ITCPD_OK;
}
#endif  // ITCPD_I8_PROBE
#endif  // ITCPD_I8_HOST_EMULATION

}  // namespace itcpd
