// Gather / sampling kernels of the randomized CP-ALS path (HBM-bound integer + gather work).
//   pivot_hadamard        src/algebra/had_contract.jl:277-295
//   fused_flatten_sample  src/algebra/pivot_mapping.jl:59-85
//   sketched_matricization src/algebra/pivot_mapping.jl:111-140
//   sample_factor_matrices src/math_tools/probability.jl:12-33
#include "common.cuh"

namespace itcpd {

struct SDims {
    int n;
    int64_t ext[ITCPD_MAX_ORDER], dim[ITCPD_MAX_ORDER];
};
static SDims sdims(const itcpd_ctx *c) {
    SDims d;
    d.n = c->order;
    for (int i = 0; i < c->order; ++i) { d.ext[i] = (i == 0) ? c->ld0 : c->dims[i]; d.dim[i] = c->dims[i]; }
    return d;
}
struct SFac { const double *a[ITCPD_MAX_ORDER]; };

// ---- sampled Khatri-Rao rows: K[s, r] = prod_{m != mode} A_m[piv[s, col(m)], r]  (ascending m) ----
__global__ void pivot_hadamard_kernel(SFac fp, SDims d, int mode, int R, int64_t nsamp, const int64_t *__restrict__ piv,
                                      double *__restrict__ K) {
    const int64_t total = nsamp * R;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = idx % nsamp;
        const int r = (int)(idx / nsamp);
        double v = 1.0;
        int col = 0;
        for (int m = 0; m < d.n; ++m) {
            if (m == mode) continue;
            const int64_t i = piv[s + nsamp * col] - 1;  // 1-based as the reference stores them
            v = v * fp.a[m][i + d.dim[m] * (int64_t)r];
            ++col;
        }
        K[idx] = v;
    }
}

int k_pivot_hadamard(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv_dev, double *K_dev) {
    SFac fp;
    for (int m = 0; m < c->order; ++m) fp.a[m] = c->A[m].as<double>();
    const int64_t total = nsamp * c->rank;
    pivot_hadamard_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 16), 256, 0, c->stream>>>(fp, sdims(c), mode, c->rank,
                                                                                                              nsamp, piv_dev, K_dev);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ---- fibre gather: out[i, s] = T[coords(s) with mode index i] ----
__global__ void __launch_bounds__(128) gather_fibers_kernel(const double *__restrict__ T, SDims d, int mode, int64_t nsamp,
                                                            const int64_t *__restrict__ piv, double *__restrict__ out) {
    const int64_t s = blockIdx.x;
    int64_t off = 0, str = 1, stride_mode = 1;
    int col = 0;
    for (int m = 0; m < d.n; ++m) {
        if (m == mode) stride_mode = str;
        else { off += (piv[s + nsamp * col] - 1) * str; ++col; }
        str *= d.ext[m];
    }
    const int64_t I = d.dim[mode];
    for (int64_t i = threadIdx.x; i < I; i += 128) out[i + I * s] = T[off + i * stride_mode];
}

int k_gather_fibers(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv_dev, double *out_dev) {
    if (nsamp == 0) return ITCPD_OK;
    gather_fibers_kernel<<<(unsigned)nsamp, 128, 0, c->stream>>>(c->T.as<double>(), sdims(c), mode, nsamp, piv_dev, out_dev);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ---- sampled MTTKRP in one pass: M_s[i, r] = sum_s T[fibre(s)][i] * K[s, r]   (= T_s K of ProjectionAlgorithm.jl:60-62) ----
// The gathered unfolding T_s (I x nsamp) is never materialised: a CTA owns 64 rows x 64 rank columns and a chunk of the samples,
// stages 16 fibres' segments and 16 rows of K per step in shared memory and accumulates a 4 x 4 register tile per thread.  The
// sample chunks (split-K: with I / 64 row blocks alone only a handful of SMs would have work) write partial tiles that are added
// in chunk order (fixed order: bitwise reproducible).  Bytes: 8 per gathered element (a 32-byte sector per element for modes
// other than the first: inherent to a strided fibre) + 8 R per sample for K.
constexpr int SM_BI = 64, SM_BR = 64, SM_SB = 16;

__global__ void __launch_bounds__(256) sampled_mttkrp_kernel(const double *__restrict__ T, const double *__restrict__ Ts, SDims d, int mode, int64_t nsamp,
                                                             int64_t chunk, const int64_t *__restrict__ piv, const double *__restrict__ K, int R,
                                                             double *__restrict__ part) {
    __shared__ double st[SM_SB][SM_BI + 1], sk[SM_SB][SM_BR + 1];
    __shared__ int64_t s_off[SM_SB];
    const int64_t I = d.dim[mode];
    const int64_t i0 = (int64_t)blockIdx.x * SM_BI;
    const int r0 = blockIdx.y * SM_BR;
    const int64_t s_begin = (int64_t)blockIdx.z * chunk, s_end = min(nsamp, s_begin + chunk);
    int64_t stride_mode = 1;
    for (int m = 0; m < mode; ++m) stride_mode *= d.ext[m];
    const int ti = threadIdx.x & 15, tr = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int64_t sb = s_begin; sb < s_end; sb += SM_SB) {
        if (threadIdx.x < SM_SB) {   // fibre start offsets of this batch (Ts given: the fibres were gathered before, column s of Ts)
            const int64_t s = sb + threadIdx.x;
            int64_t off = 0, str = 1;
            if (Ts) off = I * s;
            else if (s < s_end) {
                int col = 0;
                for (int m = 0; m < d.n; ++m) {
                    if (m != mode) { off += (piv[s + nsamp * col] - 1) * str; ++col; }
                    str *= d.ext[m];
                }
            }
            s_off[threadIdx.x] = off;
        }
        __syncthreads();
        for (int q = threadIdx.x; q < SM_SB * SM_BI; q += 256) {
            const int ss = q / SM_BI, ii = q % SM_BI;   // consecutive threads walk along one fibre
            const int64_t s = sb + ss, i = i0 + ii;
            st[ss][ii] = (s < s_end && i < I) ? (Ts ? Ts[s_off[ss] + i] : T[s_off[ss] + i * stride_mode]) : 0.0;
        }
        for (int q = threadIdx.x; q < SM_SB * SM_BR; q += 256) {
            const int rr = q / SM_SB, ss = q % SM_SB;   // consecutive threads read consecutive samples of one column of K
            const int64_t s = sb + ss;
            const int r = r0 + rr;
            sk[ss][rr] = (s < s_end && r < R) ? K[s + nsamp * (int64_t)r] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int ss = 0; ss < SM_SB; ++ss) {
            double tv[4], kv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) tv[a] = st[ss][ti + 16 * a];
#pragma unroll
            for (int b = 0; b < 4; ++b) kv[b] = sk[ss][tr + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(tv[a], kv[b], acc[a][b]);
        }
        __syncthreads();
    }
    double *dst = part + (size_t)blockIdx.z * (size_t)I * R;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t i = i0 + ti + 16 * a;
            const int r = r0 + tr + 16 * b;
            if (i < I && r < R) dst[i + I * (int64_t)r] = acc[a][b];
        }
}

// M (I x R) = T_s K.  Ts_dev == null: the fibres are read from the tensor through the pivots (T_s is never materialised);
// otherwise Ts_dev is the cached gathered unfolding (pivot-projected solvers: the tensor may be gone).  Scratch: c->work2.
int k_sampled_mttkrp(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv_dev, const double *Ts_dev, const double *K_dev, double *M_dev) {
    const int R = c->rank;
    const int64_t I = c->dims[mode];
    const int64_t rowb = ceil_div(I, SM_BI), colb = ceil_div(R, SM_BR);
    // enough (row block, column block, sample chunk) CTAs for two per SM, chunks a multiple of the 16-sample batch
    int64_t nchunks = std::max<int64_t>(1, std::min<int64_t>(ceil_div(2 * (int64_t)c->sm_count, rowb * colb), ceil_div(nsamp, 4 * SM_SB)));
    const int64_t chunk = ceil_div(ceil_div(nsamp, nchunks), SM_SB) * SM_SB;
    nchunks = ceil_div(nsamp, chunk);
    double *part = M_dev;
    if (nchunks > 1) {
        TRY(c->work2.reserve((size_t)nchunks * (size_t)I * R * 8));
        part = c->work2.as<double>();
    }
    dim3 grid((unsigned)rowb, (unsigned)colb, (unsigned)nchunks);
    sampled_mttkrp_kernel<<<grid, 256, 0, c->stream>>>(c->T.as<double>(), Ts_dev, sdims(c), mode, nsamp, chunk, piv_dev, K_dev, R, part);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    if (nchunks > 1) TRY(k_sum_slices(c, part, I * R, (int)nchunks, M_dev));
    return ITCPD_OK;
}

// ---- sparse-sign sketch of the unfolding: out[i, j] = sum_{e in row j} val[e] T_(mode)[i, col[e]] ----
// (row_ptr, col, val) is the sketch in CSR-by-sketch-row order, entries of a row in increasing
// non-zero order (the order the reference's dict_rows visits them, pivot_mapping.jl:127-137).
// The column number of every non-zero is decoded ONCE into the element offset of that column's first entry (sketch_offsets_kernel,
// in place); the gather loop is then one broadcast load of (offset, value) and one load of T per non-zero -- the 64-bit
// divisions of the decode used to be repeated by every thread of every CTA and bounded the kernel.
__global__ void __launch_bounds__(256) sketch_offsets_kernel(SDims d, int mode, int64_t nnz, int64_t *__restrict__ col) {
    for (int64_t e = blockIdx.x * 256ll + threadIdx.x; e < nnz; e += 256ll * gridDim.x) {
        int64_t rem = col[e], off = 0, str = 1;
        for (int m = 0; m < d.n; ++m) {
            if (m != mode) { off += (rem % d.dim[m]) * str; rem /= d.dim[m]; }
            str *= d.ext[m];
        }
        col[e] = off;
    }
}

__global__ void __launch_bounds__(128) sketch_kernel(const double *__restrict__ T, int64_t I, int64_t stride_mode, const int64_t *__restrict__ row_ptr,
                                                     const int64_t *__restrict__ off, const double *__restrict__ val,
                                                     double *__restrict__ out) {
    const int64_t j = blockIdx.x;
    const int64_t e0 = row_ptr[j], e1 = row_ptr[j + 1];
    for (int64_t i = threadIdx.x + 128ll * blockIdx.y; i < I; i += 128ll * gridDim.y) {
        const double *Ti = T + i * stride_mode;
        double acc = 0.0;
        for (int64_t e = e0; e < e1; ++e) acc = fma(val[e], Ti[off[e]], acc);
        out[i + I * j] = acc;
    }
}

// col -> col * I: the same decode for a sketch that reads the explicit unfolding (I x ncols, the mode's index fastest)
__global__ void __launch_bounds__(256) sketch_offsets_unfolded_kernel(int64_t I, int64_t nnz, int64_t *__restrict__ col) {
    for (int64_t e = blockIdx.x * 256ll + threadIdx.x; e < nnz; e += 256ll * gridDim.x) col[e] *= I;
}

// unfolded_dev: nullptr = gather from the tensor in place (strided for every mode but the first: one 8-byte element per 32-byte
// sector); else the explicit unfolding of `mode` (k_unfold), whose columns are contiguous.  Same products, same order: bitwise equal.
int k_sketch_csr(itcpd_ctx *c, int mode, int l, int64_t nnz, const int64_t *row_ptr_dev, int64_t *col_dev, const double *val_dev, double *out_dev,
                 const double *unfolded_dev) {
    const SDims d = sdims(c);
    int64_t stride_mode = 1;
    for (int m = 0; m < mode; ++m) stride_mode *= d.ext[m];
    if (nnz > 0) {
        const unsigned nb = (unsigned)std::min<int64_t>(ceil_div(nnz, 256), c->sm_count * 16);
        if (unfolded_dev) sketch_offsets_unfolded_kernel<<<nb, 256, 0, c->stream>>>(c->dims[mode], nnz, col_dev);
        else sketch_offsets_kernel<<<nb, 256, 0, c->stream>>>(d, mode, nnz, col_dev);
        c->launches++;
    }
    dim3 grid((unsigned)l, (unsigned)std::min<int64_t>(ceil_div(c->dims[mode], 128), 8));
    sketch_kernel<<<grid, 128, 0, c->stream>>>(unfolded_dev ? unfolded_dev : c->T.as<double>(), c->dims[mode], unfolded_dev ? 1 : stride_mode, row_ptr_dev,
                                               col_dev, val_dev, out_dev);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ---- omega_hadamard (had_contract.jl:300-329): sparse-sign sketch of the Khatri-Rao product, transposed output ----
// out[r, j] = sum_{e in sketch row j} val[e] * prod_{m != mode} A_m[coord_m(col[e]), r]      (R x l, column-major)
__global__ void __launch_bounds__(128) omega_hadamard_kernel(SFac fp, SDims d, int mode, int R, const int64_t *__restrict__ row_ptr,
                                                             const int64_t *__restrict__ col, const double *__restrict__ val,
                                                             double *__restrict__ out) {
    const int64_t j = blockIdx.x;
    for (int r = threadIdx.x; r < R; r += 128) {
        double acc = 0.0;
        for (int64_t e = row_ptr[j]; e < row_ptr[j + 1]; ++e) {
            int64_t rem = col[e];
            double v = 1.0;
            for (int m = 0; m < d.n; ++m) {
                if (m == mode) continue;
                v = v * fp.a[m][rem % d.dim[m] + d.dim[m] * (int64_t)r];
                rem /= d.dim[m];
            }
            acc = fma(val[e], v, acc);
        }
        out[r + (int64_t)R * j] = acc;
    }
}

int k_omega_hadamard(itcpd_ctx *c, int mode, int l, const int64_t *row_ptr_dev, const int64_t *col_dev, const double *val_dev, double *out_dev) {
    SFac fp;
    for (int m = 0; m < c->order; ++m) fp.a[m] = c->A[m].as<double>();
    omega_hadamard_kernel<<<(unsigned)l, 128, 0, c->stream>>>(fp, sdims(c), mode, c->rank, row_ptr_dev, col_dev, val_dev, out_dev);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// transposed sampled KRP: out[r, s] = prod_{m != mode} A_m[piv[s, col(m)], r]   (R x nsamp)
__global__ void pivot_hadamard_t_kernel(SFac fp, SDims d, int mode, int R, int64_t nsamp, const int64_t *__restrict__ piv,
                                        double *__restrict__ out) {
    const int64_t total = nsamp * R;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx % R);
        const int64_t s = idx / R;
        double v = 1.0;
        int col = 0;
        for (int m = 0; m < d.n; ++m) {
            if (m == mode) continue;
            v = v * fp.a[m][piv[s + nsamp * col] - 1 + d.dim[m] * (int64_t)r];
            ++col;
        }
        out[idx] = v;
    }
}

int k_pivot_hadamard_t(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv_dev, double *out_dev) {
    SFac fp;
    for (int m = 0; m < c->order; ++m) fp.a[m] = c->A[m].as<double>();
    const int64_t total = nsamp * c->rank;
    pivot_hadamard_t_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 16), 256, 0, c->stream>>>(fp, sdims(c), mode, c->rank, nsamp, piv_dev, out_dev);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ---- weighted sampling with replacement: inclusive CDF + binary search ----
__global__ void __launch_bounds__(1024) cdf_kernel(const double *__restrict__ w, int64_t n, double *__restrict__ cdf) {
    __shared__ double part[1024];
    __shared__ double wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t chunk = (n + 1023) / 1024;
    const int64_t b = tid * chunk, e = min(n, b + chunk);
    double s = 0.0;
    for (int64_t i = b; i < e; ++i) s += fabs(w[i]);
    // exclusive prefix of the 1024 chunk sums: shuffle scan inside each warp, then over the 32 warp totals (a serial loop in
    // one thread cost 1024 dependent additions, ~13 us of a 17 us kernel).  The association differs from the serial sum by
    // rounding only; the draws are compared with the reference statistically, never bit for bit (its sampler is StatsBase's).
    double incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        double t = wsum[lane], ti = t;
        for (int o = 1; o < 32; o <<= 1) {
            const double v = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += v;
        }
        wsum[lane] = ti - t;   // exclusive prefix of the warp totals
    }
    __syncthreads();
    part[tid] = wsum[wid] + (incl - s);
    double run = part[tid];
    for (int64_t i = b; i < e; ++i) { run += fabs(w[i]); cdf[i] = run; }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// seed_dev != null: the draw's seed is seed + *seed_dev (a device-side draw counter, so that a captured sweep draws fresh samples
// at every replay); the host-driven entry points pass the seed by value
__global__ void sample_kernel(const double *__restrict__ cdf, int64_t n, int64_t nsamp, uint64_t seed, const unsigned long long *__restrict__ seed_dev,
                              uint64_t stream, int64_t *__restrict__ out) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nsamp) return;
    if (seed_dev) seed += (uint64_t)*seed_dev;
    const uint64_t bits = splitmix64(splitmix64(seed ^ (stream * 0xD1B54A32D192ED03ull)) + (uint64_t)s);
    const double u = (double)(bits >> 11) * (1.0 / 9007199254740992.0) * cdf[n - 1];
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    out[s] = lo + 1;  // 1-based
}

__global__ void bump_counter_kernel(unsigned long long *ctr) { *ctr += 1ull; }

int k_sample_rows(itcpd_ctx *c, int skip_mode, int64_t nsamp, uint64_t seed, int64_t *piv_dev, unsigned long long *draw_counter_dev) {
    if (draw_counter_dev) {   // one draw = one seed: advance the device-side draw counter first (host loop: `seed += 1` before every draw)
        bump_counter_kernel<<<1, 1, 0, c->stream>>>(draw_counter_dev);
        c->launches++;
    }
    int col = 0;
    for (int m = 0; m < c->order; ++m) {
        if (m == skip_mode) continue;
        const int64_t n = mode_rows(c, m);
        TRY(c->work.reserve((size_t)n * 8));
        cdf_kernel<<<1, 1024, 0, c->stream>>>(c->lev[m].as<double>(), n, c->work.as<double>());
        sample_kernel<<<(unsigned)ceil_div(nsamp, 256), 256, 0, c->stream>>>(c->work.as<double>(), n, nsamp, seed, draw_counter_dev, (uint64_t)m,
                                                                           piv_dev + nsamp * col);
        c->launches += 2;
        ++col;
    }
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// one weighted draw of nsamp rows from an arbitrary device weight vector (the sharded path samples from gathered scores)
int k_cdf_sample(itcpd_ctx *c, const double *weights_dev, int64_t n, int64_t nsamp, uint64_t seed, uint64_t stream_id, int64_t *out_dev) {
    TRY(c->work.reserve((size_t)n * 8));
    cdf_kernel<<<1, 1024, 0, c->stream>>>(weights_dev, n, c->work.as<double>());
    sample_kernel<<<(unsigned)ceil_div(nsamp, 256), 256, 0, c->stream>>>(c->work.as<double>(), n, nsamp, seed, nullptr, stream_id, out_dev);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ---- small dense products for the sampled normal equations ----
// C (m x n) = A (m x k) * B (k x n), all column-major; m = I_mode rows, k = nsamp, n = R
__global__ void __launch_bounds__(256) gemm_nn_kernel(const double *__restrict__ A, const double *__restrict__ B, int64_t m, int64_t k, int n,
                                                      double *__restrict__ C) {
    __shared__ double sa[16][17], sb[16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t row = blockIdx.x * 16ll + tx;
    const int colc = blockIdx.y * 16 + ty;
    double acc = 0.0;
    for (int64_t k0 = 0; k0 < k; k0 += 16) {
        sa[ty][tx] = (row < m && k0 + ty < k) ? A[row + m * (k0 + ty)] : 0.0;
        sb[ty][tx] = (k0 + tx < k && colc < n) ? B[(k0 + tx) + k * (int64_t)colc] : 0.0;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) acc = fma(sa[q][tx], sb[ty][q], acc);
        __syncthreads();
    }
    if (row < m && colc < n) C[row + m * (int64_t)colc] = acc;
}

int k_small_gemm_nn(itcpd_ctx *c, const double *A, const double *B, int64_t m, int64_t k, int n, double *C) {
    dim3 grid((unsigned)ceil_div(m, 16), (unsigned)ceil_div(n, 16));
    gemm_nn_kernel<<<grid, 256, 0, c->stream>>>(A, B, m, k, n, C);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

}  // namespace itcpd
