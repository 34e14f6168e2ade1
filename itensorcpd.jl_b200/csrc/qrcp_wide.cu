// Column-pivoted Householder QR of a (possibly very wide) m x n matrix on the device -- the `qr(A, ColumnNorm())`
// of the pivot-projected solvers' setup (src/optimizers/als_optimizers/randomized/qr_lev_score_sampled.jl:23,127;
// src/algebra/SEQRCS.jl:152,166).  Only the pivot order and diag(R) are consumed by the reference
// (effective rank + projector columns), so Q is kept implicitly as Householder vectors.
//
// Two kernels per elimination step (HBM-bound: the trailing matrix streams once per step):
//   qrw_pivot_house_kernel  (1 CTA)   : arg-max of the column norms (first maximum), column swap, Householder
//                                       generation (dlarfg), diag(R)
//   qrw_apply_kernel        (1 warp per trailing column): a <- (I - tau v v^T) a, EXACT norm of the remaining part
//                                       (no dlaqp2 down-dating needed: the column is in flight anyway), per-CTA maxima
#include "common.cuh"
#include <cfloat>
#include <cstdlib>

namespace itcpd {

constexpr int QW_WARPS = 8;

__global__ void __launch_bounds__(256) qrw_norm_init_kernel(const double *__restrict__ A, int64_t m, int64_t n, double *__restrict__ vn,
                                                            double *__restrict__ bmax_v, int64_t *__restrict__ bmax_i) {
    __shared__ double sv[QW_WARPS];
    __shared__ int64_t si[QW_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t k = blockIdx.x * (int64_t)QW_WARPS + w;
    double nr = -1.0;
    if (k < n) {
        const double *a = A + m * k;
        double s = 0.0;
        for (int64_t i = lane; i < m; i += 32) s = fma(a[i], a[i], s);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        nr = sqrt(s);
        if (lane == 0) vn[k] = nr;
    }
    if (lane == 0) { sv[w] = nr; si[w] = k; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bv = sv[0]; int64_t bi = si[0];
        for (int q = 1; q < QW_WARPS; ++q) if (sv[q] > bv) { bv = sv[q]; bi = si[q]; }
        bmax_v[blockIdx.x] = bv; bmax_i[blockIdx.x] = bi;
    }
}

// scal[0] = tau, scal[1] = beta (= R[j,j])
__global__ void __launch_bounds__(1024) qrw_pivot_house_kernel(double *__restrict__ A, int64_t m, int64_t n, int64_t j, double *__restrict__ vn,
                                                               const double *__restrict__ bmax_v, const int64_t *__restrict__ bmax_i,
                                                               int64_t nblocks, int64_t *__restrict__ jpvt, double *__restrict__ vbuf,
                                                               double *__restrict__ scal, double *__restrict__ rdiag) {
    __shared__ double sv[32];
    __shared__ int64_t si[32];
    __shared__ double s_red[32];
    __shared__ int64_t s_p;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // ---- pivot: first maximum over the per-CTA maxima of the trailing columns ----
    double bv = -1.0; int64_t bi = INT64_MAX;
    for (int64_t b = tid; b < nblocks; b += 1024) {
        const double v = bmax_v[b]; const int64_t ix = bmax_i[b];
        if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, o);
        const int64_t oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[w] = bv; si[w] = bi; }
    __syncthreads();
    if (tid == 0) {
        double v = sv[0]; int64_t ix = si[0];
        for (int q = 1; q < 32; ++q) if (sv[q] > v || (sv[q] == v && si[q] < ix)) { v = sv[q]; ix = si[q]; }
        s_p = (ix == INT64_MAX || ix < j) ? j : ix;
    }
    __syncthreads();
    const int64_t p = s_p;
    if (p != j) {
        double *cj = A + m * j, *cp = A + m * p;
        for (int64_t i = tid; i < m; i += 1024) { const double a = cj[i]; cj[i] = cp[i]; cp[i] = a; }
        if (tid == 0) { const int64_t q = jpvt[j]; jpvt[j] = jpvt[p]; jpvt[p] = q; vn[p] = vn[j]; }
        __syncthreads();
    }
    // ---- Householder generator on A[j:m, j] (LAPACK dlarfg) ----
    double *col = A + m * j;
    const int64_t mj = m - j;
    double ss = 0.0;
    for (int64_t i = 1 + tid; i < mj; i += 1024) ss = fma(col[j + i], col[j + i], ss);
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) s_red[w] = ss;
    __syncthreads();
    double tot = 0.0;
    for (int q = 0; q < 32; ++q) tot += s_red[q];
    const double xnorm = sqrt(tot);
    const double alpha = col[j];
    __syncthreads();
    double tau = 0.0, beta = alpha, scl = 0.0;
    if (xnorm != 0.0) {
        beta = -copysign(hypot(alpha, xnorm), alpha);
        tau = (beta - alpha) / beta;
        scl = 1.0 / (alpha - beta);
    }
    for (int64_t i = 1 + tid; i < mj; i += 1024) { const double v = col[j + i] * scl; col[j + i] = v; vbuf[i] = v; }
    if (tid == 0) { vbuf[0] = 1.0; col[j] = beta; scal[0] = tau; scal[1] = beta; rdiag[j] = beta; }
}

__global__ void __launch_bounds__(256) qrw_apply_kernel(double *__restrict__ A, int64_t m, int64_t n, int64_t j, const double *__restrict__ vbuf,
                                                        const double *__restrict__ scal, double *__restrict__ vn,
                                                        double *__restrict__ bmax_v, int64_t *__restrict__ bmax_i) {
    extern __shared__ double sh_v[];  // the Householder vector, shared by the CTA's 8 columns
    __shared__ double sv[QW_WARPS];
    __shared__ int64_t si[QW_WARPS];
    const int64_t mj = m - j;
    for (int64_t i = threadIdx.x; i < mj; i += 256) sh_v[i] = vbuf[i];
    __syncthreads();
    const double tau = scal[0];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t k = j + 1 + blockIdx.x * (int64_t)QW_WARPS + w;
    double nr = -1.0;
    if (k < n) {
        double *a = A + m * k + j;
        double dot = 0.0;
        for (int64_t i = lane; i < mj; i += 32) dot = fma(sh_v[i], a[i], dot);
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        const double f = tau * dot;
        double s = 0.0;
        for (int64_t i = lane; i < mj; i += 32) {
            const double x = fma(-f, sh_v[i], a[i]);
            a[i] = x;
            if (i > 0) s = fma(x, x, s);
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        nr = sqrt(s);
        if (lane == 0) vn[k] = nr;
    }
    if (lane == 0) { sv[w] = nr; si[w] = k; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bv = sv[0]; int64_t bi = si[0];
        for (int q = 1; q < QW_WARPS; ++q) if (sv[q] > bv) { bv = sv[q]; bi = si[q]; }
        bmax_v[blockIdx.x] = bv; bmax_i[blockIdx.x] = bi;
    }
}

// The same update with the column held in registers: one DRAM read and one write per element (the kernel above walks the column twice
// and takes the second pass from L1/L2, 40 % / 62 % hit rates in the ncu capture).  A lane owns rows j0 + lane + 32 q, q < NR, where
// j0 = j rounded down to a multiple of 4 so that a warp's 256-byte row segments stay 32-byte aligned as j advances; all NR loads
// are in flight before the first use.  Chosen when m - j0 <= 32 * NR for an instantiated NR.  1024 x 44032, all 1024 steps: 83.9 ms
// against 99.5 ms for the two-pass kernel (first 128 steps 5.2 TB/s against 4.0); NR = 32 runs two CTAs per SM at 128 registers with
// 180 bytes of spills -- one CTA per SM without spills measured 89.4 ms (profiles/r2_qrcp_wide_ncu.txt).
template <int NR, int MINB>
__global__ void __launch_bounds__(256, MINB) qrw_apply_reg_kernel(double *__restrict__ A, int64_t m, int64_t n, int64_t j,
                                                                              const double *__restrict__ vbuf, const double *__restrict__ scal,
                                                                              double *__restrict__ vn, double *__restrict__ bmax_v,
                                                                              int64_t *__restrict__ bmax_i) {
    __shared__ double sh_v[32 * NR];   // v on rows [j0, j0 + 32 NR): zero before row j and past row m
    __shared__ double sv[QW_WARPS];
    __shared__ int64_t si[QW_WARPS];
    const int64_t j0 = j & ~(int64_t)3;
    const int pre = (int)(j - j0), len = (int)(m - j0);
    for (int i = threadIdx.x; i < 32 * NR; i += 256) sh_v[i] = (i >= pre && i < len) ? vbuf[i - pre] : 0.0;
    __syncthreads();
    const double tau = scal[0];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t k = j + 1 + blockIdx.x * (int64_t)QW_WARPS + w;
    double nr = -1.0;
    if (k < n) {
        double *a = A + m * k + j0;
        double x[NR];
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const int i = lane + 32 * q;
            x[q] = (i >= pre && i < len) ? a[i] : 0.0;
        }
        double dot = 0.0;
#pragma unroll
        for (int q = 0; q < NR; ++q) dot = fma(sh_v[lane + 32 * q], x[q], dot);
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        const double f = tau * dot;
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const int i = lane + 32 * q;
            const double y = fma(-f, sh_v[i], x[q]);
            if (i >= pre && i < len) a[i] = y;
            if (i > pre) s = fma(y, y, s);      // rows below the pivot row (y is 0 past row m)
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        nr = sqrt(s);
        if (lane == 0) vn[k] = nr;
    }
    if (lane == 0) { sv[w] = nr; si[w] = k; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bv = sv[0]; int64_t bi = si[0];
        for (int q = 1; q < QW_WARPS; ++q) if (sv[q] > bv) { bv = sv[q]; bi = si[q]; }
        bmax_v[blockIdx.x] = bv; bmax_i[blockIdx.x] = bi;
    }
}

template <int NR, int MINB = 2>
static void launch_apply_reg(itcpd_ctx *c, unsigned nblocks, double *A, int64_t m, int64_t n, int64_t j, const double *vbuf, const double *scal,
                             double *vn, double *bmax_v, int64_t *bmax_i) {
    qrw_apply_reg_kernel<NR, MINB><<<nblocks, 256, 0, c->stream>>>(A, m, n, j, vbuf, scal, vn, bmax_v, bmax_i);
}

__global__ void iota_kernel(int64_t *x, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) x[i] = i;
}

// In-place QRCP of A (m x n, lda = m). jpvt_dev: n int64 (0-based on return), rdiag_dev: min(m,n) doubles.
// Performs `steps` eliminations (<= min(m,n)); the first `steps` entries of jpvt are the pivots in order.
int k_qrcp_wide(itcpd_ctx *c, double *A, int64_t m, int64_t n, int64_t steps, int64_t *jpvt_dev, double *rdiag_dev) {
    const int64_t kmax = std::min<int64_t>(std::min(m, n), steps);
    const int64_t nb0 = ceil_div(n, QW_WARPS);
    // workspace: vn[n] | bmax_v[nb0] | vbuf[m] | scal[2] | bmax_i[nb0]
    TRY(c->work2.reserve((size_t)(n + nb0 + m + 2) * 8 + (size_t)nb0 * 8 + 64));
    double *vn = c->work2.as<double>();
    double *bmax_v = vn + n, *vbuf = bmax_v + nb0, *scal = vbuf + m;
    int64_t *bmax_i = reinterpret_cast<int64_t *>(scal + 2);
    iota_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, c->stream>>>(jpvt_dev, n);
    qrw_norm_init_kernel<<<(unsigned)nb0, 256, 0, c->stream>>>(A, m, n, vn, bmax_v, bmax_i);
    c->launches += 2;
    static bool attr[64] = {false};  // function attributes are per device
    if (!attr[c->device & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(qrw_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr[c->device & 63] = true;
    }
    ARG_CHECK((size_t)m * 8 <= 200 * 1024, "QRCP supports at most 25600 rows");
    int64_t nblocks = nb0;
    const bool reg_path = getenv("ITCPD_QRCP_TWO_PASS") == nullptr;   // the two-pass kernel stays for m > 2048 (and for A/B comparisons)
    for (int64_t j = 0; j < kmax; ++j) {
        qrw_pivot_house_kernel<<<1, 1024, 0, c->stream>>>(A, m, n, j, vn, bmax_v, bmax_i, nblocks, jpvt_dev, vbuf, scal, rdiag_dev);
        c->launches++;
        const int64_t trailing = n - j - 1;
        nblocks = ceil_div(std::max<int64_t>(trailing, 0), QW_WARPS);
        if (trailing > 0) {
            const int64_t span = m - (j & ~(int64_t)3);   // rows a warp covers, from the aligned start
            const unsigned nb = (unsigned)nblocks;
            if (!reg_path || span > 2048) qrw_apply_kernel<<<nb, 256, (size_t)(m - j) * 8, c->stream>>>(A, m, n, j, vbuf, scal, vn, bmax_v, bmax_i);
            else if (span > 1024) launch_apply_reg<64, 1>(c, nb, A, m, n, j, vbuf, scal, vn, bmax_v, bmax_i);
            else if (span > 512) launch_apply_reg<32, 2>(c, nb, A, m, n, j, vbuf, scal, vn, bmax_v, bmax_i);
            else if (span > 256) launch_apply_reg<16>(c, nb, A, m, n, j, vbuf, scal, vn, bmax_v, bmax_i);
            else if (span > 128) launch_apply_reg<8>(c, nb, A, m, n, j, vbuf, scal, vn, bmax_v, bmax_i);
            else if (span > 64) launch_apply_reg<4>(c, nb, A, m, n, j, vbuf, scal, vn, bmax_v, bmax_i);
            else launch_apply_reg<2>(c, nb, A, m, n, j, vbuf, scal, vn, bmax_v, bmax_i);
            c->launches++;
        }
    }
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// Explicit unfolding T_(mode): I_mode x prod(others), others in original order, first fastest.
struct UDims { int n; int64_t ext[ITCPD_MAX_ORDER], dim[ITCPD_MAX_ORDER]; };

// mode 0: the unfolding is the tensor itself up to the padding of the leading dimension -- a coalesced copy
__global__ void unfold_kernel(const double *__restrict__ T, UDims d, int mode, int64_t ncols, double *__restrict__ out) {
    const int64_t I = d.dim[mode];
    int64_t stride_mode = 1;
    for (int q = 0; q < mode; ++q) stride_mode *= d.ext[q];
    const int64_t total = I * ncols;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx % I;
        int64_t rem = idx / I, off = 0, str = 1;
        for (int q = 0; q < d.n; ++q) {
            if (q != mode) { off += (rem % d.dim[q]) * str; rem /= d.dim[q]; }
            str *= d.ext[q];
        }
        out[idx] = T[off + i * stride_mode];
    }
}

// mode >= 1: T is [a (the modes before, fastest) | i (the mode) | b (the modes after)], the unfolding is [i | a | b]: a batch of
// 2-D transposes, 32 x 32 tiles through shared memory so that both the read (along a) and the write (along i) are coalesced.
// One read and one write of the tensor (HBM-bound); the strided gathers of the sketch then run on contiguous columns.
__global__ void __launch_bounds__(256) unfold_tiled_kernel(const double *__restrict__ T, UDims d, int mode, int64_t A, int64_t I, int64_t tiles_a,
                                                           int64_t tiles_i, double *__restrict__ out) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    int64_t blk = blockIdx.x;
    const int64_t ta = blk % tiles_a; blk /= tiles_a;
    const int64_t ti = blk % tiles_i;
    const int64_t b = blk / tiles_i;
    int64_t stride_mode = 1;
    for (int q = 0; q < mode; ++q) stride_mode *= d.ext[q];
    const int64_t stride_b = stride_mode * d.ext[mode];
    {
        const int64_t a = ta * 32 + tx;
        if (a < A) {
            int64_t rem = a, off = 0, str = 1;
            for (int q = 0; q < mode; ++q) { off += (rem % d.dim[q]) * str; rem /= d.dim[q]; str *= d.ext[q]; }
            const double *src = T + off + b * stride_b;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int64_t i = ti * 32 + ty + 8 * r;
                if (i < I) tile[ty + 8 * r][tx] = src[i * stride_mode];
            }
        }
    }
    __syncthreads();
    const int64_t i = ti * 32 + tx;
    if (i < I) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int64_t a = ta * 32 + ty + 8 * r;
            if (a < A) out[i + I * (a + A * b)] = tile[tx][ty + 8 * r];
        }
    }
}

int k_unfold(itcpd_ctx *c, int mode, double *out) {
    UDims d;
    d.n = c->order;
    for (int i = 0; i < c->order; ++i) { d.ext[i] = (i == 0) ? c->ld0 : c->dims[i]; d.dim[i] = c->dims[i]; }
    const int64_t I = c->dims[mode], ncols = c->nelem / I;
    if (mode == 0) {
        unfold_kernel<<<c->sm_count * 16, 256, 0, c->stream>>>(c->T.as<double>(), d, mode, ncols, out);
    } else {
        int64_t A = 1;
        for (int q = 0; q < mode; ++q) A *= c->dims[q];
        const int64_t B = ncols / A, tiles_a = ceil_div(A, 32), tiles_i = ceil_div(I, 32);
        ARG_CHECK(tiles_a * tiles_i * B < ((int64_t)1 << 31), "unfolding: too many tiles for one grid");
        unfold_tiled_kernel<<<(unsigned)(tiles_a * tiles_i * B), 256, 0, c->stream>>>(c->T.as<double>(), d, mode, A, I, tiles_a, tiles_i, out);
    }
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

}  // namespace itcpd
