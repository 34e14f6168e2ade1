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

// ---- sparse-sign sketch of the unfolding: out[i, j] = sum_{e in row j} val[e] T_(mode)[i, col[e]] ----
// (row_ptr, col, val) is the sketch in CSR-by-sketch-row order, entries of a row in increasing
// non-zero order (the order the reference's dict_rows visits them, pivot_mapping.jl:127-137).
__global__ void __launch_bounds__(128) sketch_kernel(const double *__restrict__ T, SDims d, int mode, const int64_t *__restrict__ row_ptr,
                                                     const int64_t *__restrict__ col, const double *__restrict__ val,
                                                     double *__restrict__ out) {
    const int64_t j = blockIdx.x;
    const int64_t I = d.dim[mode];
    int64_t stride_mode = 1;
    for (int m = 0; m < mode; ++m) stride_mode *= d.ext[m];
    for (int64_t i = threadIdx.x + 128ll * blockIdx.y; i < I; i += 128ll * gridDim.y) {
        double acc = 0.0;
        for (int64_t e = row_ptr[j]; e < row_ptr[j + 1]; ++e) {
            int64_t rem = col[e], off = 0, str = 1;
            for (int m = 0; m < d.n; ++m) {
                if (m != mode) { off += (rem % d.dim[m]) * str; rem /= d.dim[m]; }
                str *= d.ext[m];
            }
            acc = fma(val[e], T[off + i * stride_mode], acc);
        }
        out[i + I * j] = acc;
    }
}

int k_sketch_csr(itcpd_ctx *c, int mode, int l, const int64_t *row_ptr_dev, const int64_t *col_dev, const double *val_dev, double *out_dev) {
    dim3 grid((unsigned)l, (unsigned)std::min<int64_t>(ceil_div(c->dims[mode], 128), 8));
    sketch_kernel<<<grid, 128, 0, c->stream>>>(c->T.as<double>(), sdims(c), mode, row_ptr_dev, col_dev, val_dev, out_dev);
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
    const int tid = threadIdx.x;
    const int64_t chunk = (n + 1023) / 1024;
    const int64_t b = tid * chunk, e = min(n, b + chunk);
    double s = 0.0;
    for (int64_t i = b; i < e; ++i) s += fabs(w[i]);
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        double run = 0.0;
        for (int q = 0; q < 1024; ++q) { const double v = part[q]; part[q] = run; run += v; }
    }
    __syncthreads();
    double run = part[tid];
    for (int64_t i = b; i < e; ++i) { run += fabs(w[i]); cdf[i] = run; }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void sample_kernel(const double *__restrict__ cdf, int64_t n, int64_t nsamp, uint64_t seed, uint64_t stream,
                              int64_t *__restrict__ out) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nsamp) return;
    const uint64_t bits = splitmix64(splitmix64(seed ^ (stream * 0xD1B54A32D192ED03ull)) + (uint64_t)s);
    const double u = (double)(bits >> 11) * (1.0 / 9007199254740992.0) * cdf[n - 1];
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    out[s] = lo + 1;  // 1-based
}

int k_sample_rows(itcpd_ctx *c, int skip_mode, int64_t nsamp, uint64_t seed, int64_t *piv_dev) {
    int col = 0;
    for (int m = 0; m < c->order; ++m) {
        if (m == skip_mode) continue;
        const int64_t n = mode_rows(c, m);
        TRY(c->work.reserve((size_t)n * 8));
        cdf_kernel<<<1, 1024, 0, c->stream>>>(c->lev[m].as<double>(), n, c->work.as<double>());
        sample_kernel<<<(unsigned)ceil_div(nsamp, 256), 256, 0, c->stream>>>(c->work.as<double>(), n, nsamp, seed, (uint64_t)m,
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
    sample_kernel<<<(unsigned)ceil_div(nsamp, 256), 256, 0, c->stream>>>(c->work.as<double>(), n, nsamp, seed, stream_id, out_dev);
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
