// Small device kernels around the GEMM: Gram / Hadamard / normalise / fit (HBM- or latency-bound),
// second-level contractions of the dimension tree, tensor generation, reconstruction.
// Every reduction is fixed-order (no atomics) so that repeated runs and replicated ranks are bitwise equal.
#include "common.cuh"

namespace itcpd {

struct Dims {
    int n;
    int64_t ext[ITCPD_MAX_ORDER];  // storage extents (ext[0] = ld0)
    int64_t dim[ITCPD_MAX_ORDER];  // logical extents
};

static Dims ctx_dims(const itcpd_ctx *c) {
    Dims d;
    d.n = c->order;
    for (int i = 0; i < c->order; ++i) { d.ext[i] = (i == 0) ? c->ld0 : c->dims[i]; d.dim[i] = c->dims[i]; }
    return d;
}

template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *sh) {
    // fixed-order tree: warp shuffle then one warp over the warp sums
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (l < THREADS / 32) ? sh[l] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    }
    return v;  // valid in thread 0
}

// ------------------------------------------------------------------------------------------------
// Gram: G = A^T A  (post_solve, tensor.jl:46-49; compute_als, standard/tensor.jl:12)
// ------------------------------------------------------------------------------------------------
constexpr int GT = 16;       // output tile
constexpr int GCHUNK = 64;   // rows per smem chunk
constexpr int GSLICE = 256;  // rows per CTA slice

__global__ void __launch_bounds__(GT *GT) gram_partial_kernel(const double *__restrict__ A, const double *__restrict__ B, int64_t rows, int R,
                                                              double *__restrict__ part) {  // part = A^T B (B == A: Gram)
    __shared__ double sa[GT][GCHUNK + 1], sb[GT][GCHUNK + 1];
    const int tx = threadIdx.x % GT, ty = threadIdx.x / GT;
    const int r1 = blockIdx.x * GT + tx, r2 = blockIdx.y * GT + ty;
    const int64_t i0 = (int64_t)blockIdx.z * GSLICE;
    const int64_t i1 = min(rows, i0 + GSLICE);
    double acc = 0.0;
    for (int64_t base = i0; base < i1; base += GCHUNK) {
        for (int q = threadIdx.x; q < GT * GCHUNK; q += GT * GT) {
            const int col = q / GCHUNK, ii = q % GCHUNK;
            const int64_t i = base + ii;
            const int ca = blockIdx.x * GT + col, cb = blockIdx.y * GT + col;
            sa[col][ii] = (i < i1 && ca < R) ? A[i + rows * (int64_t)ca] : 0.0;
            sb[col][ii] = (i < i1 && cb < R) ? B[i + rows * (int64_t)cb] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int ii = 0; ii < GCHUNK; ++ii) acc = fma(sa[tx][ii], sb[ty][ii], acc);
        __syncthreads();
    }
    if (r1 < R && r2 < R) part[(size_t)blockIdx.z * R * R + r1 + (size_t)R * r2] = acc;
}

__global__ void sum_slices_kernel(const double *__restrict__ part, int64_t n, int slices, double *__restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int z = 0; z < slices; ++z) s += part[(size_t)z * n + i];
    out[i] = s;
}

int k_sum_slices(itcpd_ctx *c, const double *part, int64_t n, int slices, double *out) {
    sum_slices_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, c->stream>>>(part, n, slices, out);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

int k_gram(itcpd_ctx *c, const double *A, int64_t rows, int R, double *G) {
    const int slices = (int)std::max<int64_t>(1, ceil_div(rows, GSLICE));
    TRY(c->redux.reserve((size_t)slices * R * R * 8));
    dim3 grid((unsigned)ceil_div(R, GT), (unsigned)ceil_div(R, GT), (unsigned)slices);
    gram_partial_kernel<<<grid, GT * GT, 0, c->stream>>>(A, A, rows, R, c->redux.as<double>());
    sum_slices_kernel<<<(unsigned)ceil_div((int64_t)R * R, 256), 256, 0, c->stream>>>(c->redux.as<double>(), (int64_t)R * R, slices, G);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// cross Gram C = A^T B of two rows x R matrices (cp_cp_contract, src/algebra/cp_contract.jl:32-52)
int k_cross_gram(itcpd_ctx *c, const double *A, const double *B, int64_t rows, int R, double *C) {
    const int slices = (int)std::max<int64_t>(1, ceil_div(rows, GSLICE));
    TRY(c->redux.reserve((size_t)slices * R * R * 8));
    dim3 grid((unsigned)ceil_div(R, GT), (unsigned)ceil_div(R, GT), (unsigned)slices);
    gram_partial_kernel<<<grid, GT * GT, 0, c->stream>>>(A, B, rows, R, c->redux.as<double>());
    sum_slices_kernel<<<(unsigned)ceil_div((int64_t)R * R, 256), 256, 0, c->stream>>>(c->redux.as<double>(), (int64_t)R * R, slices, C);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// out2[0] = lp^T (hadamard_n X_n) lc   with X_n = A_prev,n^T A_n   (inner product of two CPDs)
// out2[1] = lc^T (hadamard_n G_n)  lc   with G_n = A_n^T A_n        (squared norm of the current CPD)
struct CrossPtrs { const double *x[ITCPD_MAX_ORDER]; const double *g[ITCPD_MAX_ORDER]; int n; };
__global__ void __launch_bounds__(256) cpd_diff_final_kernel(CrossPtrs p, const double *__restrict__ lp, const double *__restrict__ lc, int R,
                                                             double *__restrict__ out2) {
    __shared__ double sh[8];
    double a = 0.0, b = 0.0;
    for (int e = threadIdx.x; e < R * R; e += 256) {
        double hx = p.x[0][e], hg = p.g[0][e];
        for (int m = 1; m < p.n; ++m) { hx = hx * p.x[m][e]; hg = hg * p.g[m][e]; }
        const int r1 = e % R, r2 = e / R;
        a = fma(hx, lp[r1] * lc[r2], a);
        b = fma(hg, lc[r1] * lc[r2], b);
    }
    a = block_sum<256>(a, sh);
    b = block_sum<256>(b, sh);
    if (threadIdx.x == 0) { out2[0] = a; out2[1] = b; }
}

int k_cpd_diff_terms(itcpd_ctx *c, double *out2) {
    const int N = c->order, R = c->rank;
    TRY(c->work2.reserve((size_t)2 * N * R * R * 8));
    CrossPtrs p;
    p.n = N;
    double *base = c->work2.as<double>();
    // NOTE: k_cross_gram uses c->redux as scratch; results go to work2
    for (int n = 0; n < N; ++n) {
        double *X = base + (size_t)(2 * n) * R * R, *G = base + (size_t)(2 * n + 1) * R * R;
        TRY(k_cross_gram(c, c->prevA[n].as<double>(), c->A[n].as<double>(), c->dims[n], R, X));
        TRY(k_cross_gram(c, c->A[n].as<double>(), c->A[n].as<double>(), c->dims[n], R, G));
        if (comm_active(c) && n == N - 1) TRY(comm_allreduce_sum(c, X, (int64_t)2 * R * R));  // sharded factor: X and G are adjacent
        p.x[n] = X;
        p.g[n] = G;
    }
    cpd_diff_final_kernel<<<1, 256, 0, c->stream>>>(p, c->prev_lambda.as<double>(), c->lambda.as<double>(), R, out2);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ------------------------------------------------------------------------------------------------
// Gamma = hadamard_{m != skip} G_m, ascending m starting from ones (MttkrpAlgorithm.jl:18-31)
// ------------------------------------------------------------------------------------------------
struct GramPtrs { const double *g[ITCPD_MAX_ORDER]; int n; };

__global__ void gram_hadamard_kernel(GramPtrs p, int skip, int64_t n, double *__restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = 1.0;
    for (int m = 0; m < p.n; ++m)
        if (m != skip) v = v * p.g[m][i];
    out[i] = v;
}

int k_gram_hadamard(itcpd_ctx *c, int skip_mode, double *Gamma) {
    GramPtrs p;
    p.n = c->order;
    for (int m = 0; m < c->order; ++m) p.g[m] = c->G[m].as<double>();
    const int64_t n = (int64_t)c->rank * c->rank;
    gram_hadamard_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, c->stream>>>(p, skip_mode, n, Gamma);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ------------------------------------------------------------------------------------------------
// row_norm (math_tools/row_norm.jl:4-24): lambda_r = sqrt(sum_i X[i,r]^2), A = X ./ lambda
// split in two so that the slab-sharded mode can all-reduce the sums of squares in between
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsumsq_kernel(const double *__restrict__ X, int64_t rows, double *__restrict__ out) {
    __shared__ double sh[8];
    const double *x = X + rows * (int64_t)blockIdx.x;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < rows; i += 256) s = fma(x[i], x[i], s);
    s = block_sum<256>(s, sh);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

__global__ void scale_cols_kernel(const double *__restrict__ X, int64_t rows, int R, const double *__restrict__ sumsq,
                                  double *__restrict__ A, double *__restrict__ lambda) {
    const int r = blockIdx.y;
    const double l = sqrt(sumsq[r]);
    if (blockIdx.x == 0 && threadIdx.x == 0) lambda[r] = l;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x)
        A[i + rows * (int64_t)r] = X[i + rows * (int64_t)r] / l;
}

int k_colnorm_scale(itcpd_ctx *c, const double *X, int64_t rows, int R, double *A, double *lambda, bool rows_are_slab) {
    TRY(c->work2.reserve((size_t)R * 8));
    colsumsq_kernel<<<R, 256, 0, c->stream>>>(X, rows, c->work2.as<double>());
    c->launches++;
    if (rows_are_slab && comm_active(c)) TRY(comm_allreduce_sum(c, c->work2.as<double>(), R));
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(rows, 256), 64), (unsigned)R);
    scale_cols_kernel<<<grid, 256, 0, c->stream>>>(X, rows, R, c->work2.as<double>(), A, lambda);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ------------------------------------------------------------------------------------------------
// fit scalars (fit_check.jl:28-29, converge_checks.jl:5-11)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inner_partial_kernel(const double *__restrict__ M, const double *__restrict__ A,
                                                            const double *__restrict__ lambda, int64_t rows, int R,
                                                            double *__restrict__ part) {
    __shared__ double sh[8];
    const int64_t n = rows * R;
    double s = 0.0;
    for (int64_t q = blockIdx.x * 256ll + threadIdx.x; q < n; q += (int64_t)gridDim.x * 256) {
        const int r = (int)(q / rows);
        s = fma(M[q], A[q] * lambda[r], s);
    }
    s = block_sum<256>(s, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) fit_final_kernel(const double *__restrict__ part, int nparts, GramPtrs p,
                                                        const double *__restrict__ lambda, int R, double *__restrict__ out2) {
    __shared__ double sh[8];
    // nparts <= 256: one partial per thread, fixed-order tree (a serial loop in one thread costs 256 dependent L2 round trips)
    const double inner = block_sum<256>((int)threadIdx.x < nparts ? part[threadIdx.x] : 0.0, sh);
    double q = 0.0;
    for (int e = threadIdx.x; e < R * R; e += 256) {
        double h = p.g[0][e];
        for (int m = 1; m < p.n; ++m) h = h * p.g[m][e];
        q = fma(h, lambda[e % R] * lambda[e / R], q);
    }
    q = block_sum<256>(q, sh);
    if (threadIdx.x == 0) { out2[0] = inner; out2[1] = q; }
}

int k_fit_terms(itcpd_ctx *c, double *out2, bool reduce_now) {
    const int N = c->order, R = c->rank;
    const int64_t rows = mode_rows(c, N - 1);
    const int nparts = (int)std::min<int64_t>(ceil_div(rows * R, 256 * 4), 256);
    TRY(c->redux.reserve((size_t)nparts * 8 + 64));
    inner_partial_kernel<<<nparts, 256, 0, c->stream>>>(c->M[N - 1].as<double>(), c->A[N - 1].as<double>(), c->lambda.as<double>(),
                                                        rows, R, c->redux.as<double>());
    GramPtrs p;
    p.n = N;
    for (int m = 0; m < N; ++m) p.g[m] = c->G[m].as<double>();
    fit_final_kernel<<<1, 256, 0, c->stream>>>(c->redux.as<double>(), nparts, p, c->lambda.as<double>(), R, out2);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    // inner is a slab-partial sum (model_norm2 is replicated); the sweep driver defers this reduction to the result fetch
    if (reduce_now && comm_active(c)) TRY(comm_allreduce_sum(c, out2, 1));
    return ITCPD_OK;
}

// ------------------------------------------------------------------------------------------------
// plain (non-packed) Khatri-Rao expansion: W[k, r] = prod_f A_f[i_f(k), r], zero on padded rows
// ------------------------------------------------------------------------------------------------
struct KrpArgs {
    const double *fac[ITCPD_MAX_ORDER];
    int64_t ext[ITCPD_MAX_ORDER], dim[ITCPD_MAX_ORDER];
    int nf;
    int64_t kext;
    int R;
};

__global__ void krp_expand_kernel(KrpArgs a, double *__restrict__ W) {
    const int64_t total = a.kext * a.R;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = idx % a.kext;
        const int r = (int)(idx / a.kext);
        double v = 1.0;
        int64_t rem = k;
        for (int f = 0; f < a.nf; ++f) {
            const int64_t i = rem % a.ext[f];
            rem /= a.ext[f];
            v = (i < a.dim[f]) ? v * a.fac[f][i + a.dim[f] * (int64_t)r] : 0.0;
        }
        W[idx] = v;
    }
}

static int krp_expand(itcpd_ctx *c, int first, int last, DevBuf &dst, int64_t *kext_out) {
    KrpArgs a;
    memset(&a, 0, sizeof(a));
    a.R = c->rank;
    a.kext = 1;
    for (int m = first; m <= last; ++m) {
        a.fac[a.nf] = c->A[m].as<double>();
        a.ext[a.nf] = (m == 0) ? c->ld0 : c->dims[m];
        a.dim[a.nf] = c->dims[m];
        a.kext *= a.ext[a.nf];
        a.nf++;
    }
    *kext_out = a.kext;
    TRY(dst.reserve((size_t)a.kext * a.R * 8));
    const int64_t total = a.kext * a.R;
    krp_expand_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 16), 256, 0, c->stream>>>(a, dst.as<double>());
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ------------------------------------------------------------------------------------------------
// second level of the dimension tree: from the partial P[(group modes), r] to M_mode
//   out[i, r] = sum_b wb[b, r] sum_f wf[f, r] P[f + F (i + I b) + F I B r]
// ------------------------------------------------------------------------------------------------
// front empty: coalesced over i, b split over 8 lanes of the block, fixed-order smem reduction
__global__ void __launch_bounds__(256) partial_first_kernel(const double *__restrict__ P, const double *__restrict__ wb, int64_t I,
                                                            int64_t Ilog, int64_t B, double *__restrict__ out) {
    __shared__ double sh[8][33];
    const int ti = threadIdx.x & 31, tb = threadIdx.x >> 5;
    const int r = blockIdx.y;
    const int64_t i = blockIdx.x * 32ll + ti;
    double acc = 0.0;
    if (i < Ilog) {
        const double *p = P + i + I * B * (int64_t)r;
        if (wb) {
            const double *w = wb + B * (int64_t)r;
            for (int64_t b = tb; b < B; b += 8) acc = fma(p[I * b], w[b], acc);
        } else {
            for (int64_t b = tb; b < B; b += 8) acc += p[I * b];
        }
    }
    sh[tb][ti] = acc;
    __syncthreads();
    if (tb == 0 && i < Ilog) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += sh[q][ti];
        out[i + Ilog * (int64_t)r] = s;
    }
}

// front non-empty: one WARP per (i, r), lanes run over the contiguous front index with 16-byte loads
__global__ void __launch_bounds__(256) partial_general_kernel(const double *__restrict__ P, const double *__restrict__ wf,
                                                              const double *__restrict__ wb, int64_t F, int64_t I, int64_t B,
                                                              double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = blockIdx.x * 8ll + (threadIdx.x >> 5);
    const int r = blockIdx.y;
    if (i >= I) return;
    const double *p = P + F * i + F * I * B * (int64_t)r;
    const double *f = wf + F * (int64_t)r;
    double acc = 0.0;
    if ((F & 1) == 0) {
        for (int64_t b = 0; b < B; ++b) {
            const double2 *pb = reinterpret_cast<const double2 *>(p + F * I * b);
            const double2 *f2 = reinterpret_cast<const double2 *>(f);
            double in0 = 0.0, in1 = 0.0;
            for (int64_t q = lane; q < F / 2; q += 32) {
                const double2 pv = pb[q], fv = f2[q];
                in0 = fma(pv.x, fv.x, in0);
                in1 = fma(pv.y, fv.y, in1);
            }
            const double inner = in0 + in1;
            acc = wb ? fma(inner, wb[b + B * (int64_t)r], acc) : acc + inner;
        }
    } else {
        for (int64_t b = 0; b < B; ++b) {
            double inner = 0.0;
            const double *pb = p + F * I * b;
            for (int64_t q = lane; q < F; q += 32) inner = fma(pb[q], f[q], inner);
            acc = wb ? fma(inner, wb[b + B * (int64_t)r], acc) : acc + inner;
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[i + I * (int64_t)r] = acc;
}

int k_partial_mttkrp(itcpd_ctx *c, const double *P, int gfirst, int glast, int64_t ld_first, int mode, double *out) {
    const int R = c->rank;
    int64_t F = 1, B = 1;
    const double *wf = nullptr, *wb = nullptr;
    if (mode > gfirst) {
        TRY(krp_expand(c, gfirst, mode - 1, c->krp_scratch[0], &F));
        wf = c->krp_scratch[0].as<double>();
    }
    if (mode < glast) {
        TRY(krp_expand(c, mode + 1, glast, c->krp_scratch[1], &B));
        wb = c->krp_scratch[1].as<double>();
    }
    const int64_t Ilog = c->dims[mode];
    if (wf == nullptr) {
        // the target mode is the first of its group (no front modes; a front group of total extent 1 still carries
        // a 1 x R weight and takes the general path)
        const int64_t I = ld_first;
        dim3 grid((unsigned)ceil_div(Ilog, 32), (unsigned)R);
        partial_first_kernel<<<grid, 256, 0, c->stream>>>(P, wb, I, Ilog, B, out);
    } else {
        dim3 grid((unsigned)ceil_div(Ilog, 8), (unsigned)R);
        partial_general_kernel<<<grid, 256, 0, c->stream>>>(P, wf, wb, F, Ilog, B, out);
    }
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ------------------------------------------------------------------------------------------------
// direct one-pass MTTKRP with plain FMAs (cross-check of the tree/GEMM path; not the fast path)
// ------------------------------------------------------------------------------------------------
struct FacPtrs { const double *a[ITCPD_MAX_ORDER]; };

__global__ void __launch_bounds__(256) direct_mttkrp_kernel(const double *__restrict__ T, Dims d, FacPtrs fp, int mode, int R,
                                                            double *__restrict__ out) {
    __shared__ double sh[8];
    const int64_t i = blockIdx.x;
    const int r0 = blockIdx.y * 4;
    int64_t nother = 1, stride_mode = 1;
    for (int m = 0; m < d.n; ++m) {
        if (m != mode) nother *= d.dim[m];
        if (m < mode) stride_mode *= d.ext[m];
    }
    double acc[4] = {0, 0, 0, 0};
    for (int64_t q = threadIdx.x; q < nother; q += 256) {
        int64_t rem = q, off = i * stride_mode, str = 1;
        double w[4] = {1, 1, 1, 1};
        for (int m = 0; m < d.n; ++m) {
            if (m != mode) {
                const int64_t im = rem % d.dim[m];
                rem /= d.dim[m];
                off += im * str;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (r0 + j < R) w[j] *= fp.a[m][im + d.dim[m] * (int64_t)(r0 + j)];
            }
            str *= d.ext[m];
        }
        const double tv = T[off];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fma(tv, w[j], acc[j]);
    }
    for (int j = 0; j < 4; ++j) {
        const double s = block_sum<256>(acc[j], sh);
        if (threadIdx.x == 0 && r0 + j < R) out[i + d.dim[mode] * (int64_t)(r0 + j)] = s;
    }
}

int k_direct_mttkrp(itcpd_ctx *c, int mode, double *out) {
    FacPtrs fp;
    for (int m = 0; m < c->order; ++m) fp.a[m] = c->A[m].as<double>();
    dim3 grid((unsigned)c->dims[mode], (unsigned)ceil_div(c->rank, 4));
    direct_mttkrp_kernel<<<grid, 256, 0, c->stream>>>(c->T.as<double>(), ctx_dims(c), fp, mode, c->rank, out);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ------------------------------------------------------------------------------------------------
// counter-based N(0,1) generator: Philox4x32-10 (Salmon et al. 2011) + Box-Muller
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// two independent standard normals for pair index `pairidx` of stream `seed`
__device__ __forceinline__ void normal_pair(uint64_t seed, uint64_t pairidx, double &z0, double &z1) {
    uint32_t ctr[4] = {(uint32_t)pairidx, (uint32_t)(pairidx >> 32), 0u, 0u};
    philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = ((uint64_t)ctr[1] << 32) | ctr[0], b = ((uint64_t)ctr[3] << 32) | ctr[2];
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);  // (0,1]
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);          // [0,1)
    const double rad = sqrt(-2.0 * log(u1));
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    z0 = rad * co;
    z1 = rad * s;
}

__device__ __forceinline__ double normal_at(uint64_t seed, uint64_t e) {
    double z0, z1;
    normal_pair(seed, e >> 1, z0, z1);
    return (e & 1) ? z1 : z0;
}

__global__ void generate_kernel(double *__restrict__ T, int64_t ld0, int64_t dim0, int64_t nstore, uint64_t seed, int64_t elem_offset) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nstore; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = s % ld0, rest = s / ld0;
        T[s] = (i0 < dim0) ? normal_at(seed, (uint64_t)(i0 + dim0 * rest + elem_offset)) : 0.0;
    }
}

int k_generate(itcpd_ctx *c, uint64_t seed, int64_t elem_offset) {
    generate_kernel<<<c->sm_count * 16, 256, 0, c->stream>>>(c->T.as<double>(), c->ld0, c->dims[0], c->nstore, seed, elem_offset);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

__global__ void randn_kernel(double *__restrict__ dst, int64_t n, uint64_t seed, uint64_t off) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x)
        dst[s] = normal_at(seed, (uint64_t)s + off);
}

__global__ void add_noise_kernel(double *__restrict__ T, int64_t ld0, int64_t dim0, int64_t nstore, uint64_t seed, double sigma) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nstore; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = s % ld0, rest = s / ld0;
        if (i0 < dim0) T[s] += sigma * normal_at(seed, (uint64_t)(i0 + dim0 * rest));
    }
}

int k_add_noise(itcpd_ctx *c, uint64_t seed, double sigma) {
    add_noise_kernel<<<c->sm_count * 16, 256, 0, c->stream>>>(c->T.as<double>(), c->ld0, c->dims[0], c->nstore, seed, sigma);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

int k_randn_matrix(itcpd_ctx *c, double *dst, int64_t n, uint64_t seed, uint64_t stream_offset) {
    randn_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), 148 * 8), 256, 0, c->stream>>>(dst, n, seed, stream_offset);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ------------------------------------------------------------------------------------------------
// sum of squares (norm(T)), fixed order two-stage
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const double *__restrict__ x, int64_t n, double *__restrict__ part) {
    __shared__ double sh[8];
    double s = 0.0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) s = fma(x[i], x[i], s);
    s = block_sum<256>(s, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) sum_final_kernel(const double *__restrict__ part, int n, double *__restrict__ out) {
    __shared__ double sh[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += part[i];
    s = block_sum<256>(s, sh);
    if (threadIdx.x == 0) out[0] = s;
}

int k_sumsq(itcpd_ctx *c, const double *x, int64_t n, double *out_dev) {
    const int parts = (int)std::min<int64_t>(std::max<int64_t>(1, ceil_div(n, 4096)), (int64_t)c->sm_count * 8);
    TRY(c->redux.reserve((size_t)parts * 8));
    sumsq_partial_kernel<<<parts, 256, 0, c->stream>>>(x, n, c->redux.as<double>());
    sum_final_kernel<<<1, 256, 0, c->stream>>>(c->redux.as<double>(), parts, out_dev);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ------------------------------------------------------------------------------------------------
// leading-dimension padding (mode 0 is stored with an even leading dimension for TMA)
// ------------------------------------------------------------------------------------------------
__global__ void pad_in_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t dim0, int64_t ld0, int64_t nstore) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nstore; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = s % ld0, rest = s / ld0;
        dst[s] = (i0 < dim0) ? src[i0 + dim0 * rest] : 0.0;
    }
}
__global__ void pad_out_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t dim0, int64_t ld0, int64_t nelem) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nelem; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = e % dim0, rest = e / dim0;
        dst[e] = src[i0 + ld0 * rest];
    }
}
int k_pad_copy_in(itcpd_ctx *c, const double *src, double *dst) {
    pad_in_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(src, dst, c->dims[0], c->ld0, c->nstore);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}
int k_pad_copy_out(itcpd_ctx *c, const double *src, double *dst) {
    pad_out_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(src, dst, c->dims[0], c->ld0, c->nelem);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// ------------------------------------------------------------------------------------------------
// reconstruct (algebra/reconstruct.jl:2-9) and fused residual ||T - That||^2
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reconstruct_kernel(const double *__restrict__ T, Dims d, FacPtrs fp, const double *__restrict__ lambda,
                                                          int R, int64_t nelem, double *__restrict__ out, double *__restrict__ part) {
    __shared__ double sh[8];
    double rs = 0.0;
    for (int64_t e = blockIdx.x * 256ll + threadIdx.x; e < nelem; e += (int64_t)gridDim.x * 256) {
        int64_t idx[ITCPD_MAX_ORDER];
        int64_t rem = e, off = 0, str = 1;
        for (int m = 0; m < d.n; ++m) {
            idx[m] = rem % d.dim[m];
            rem /= d.dim[m];
            off += idx[m] * str;
            str *= d.ext[m];
        }
        double v = 0.0;
        for (int r = 0; r < R; ++r) {
            double w = lambda[r];
            for (int m = 0; m < d.n; ++m) w *= fp.a[m][idx[m] + d.dim[m] * (int64_t)r];
            v += w;
        }
        if (out) out[e] = v;
        if (part) { const double dlt = T[off] - v; rs = fma(dlt, dlt, rs); }
    }
    if (part) {
        rs = block_sum<256>(rs, sh);
        if (threadIdx.x == 0) part[blockIdx.x] = rs;
    }
}

int k_reconstruct(itcpd_ctx *c, double *out_dense, double *resid_sumsq_dev) {
    FacPtrs fp;
    for (int m = 0; m < c->order; ++m) fp.a[m] = c->A[m].as<double>();
    const int parts = (int)std::min<int64_t>(std::max<int64_t>(1, ceil_div(c->nelem, 1024)), (int64_t)c->sm_count * 8);
    double *part = nullptr;
    if (resid_sumsq_dev) {
        TRY(c->redux.reserve((size_t)parts * 8));
        part = c->redux.as<double>();
    }
    reconstruct_kernel<<<parts, 256, 0, c->stream>>>(c->T.as<double>(), ctx_dims(c), fp, c->lambda.as<double>(), c->rank, c->nelem,
                                                     out_dense, part);
    c->launches++;
    if (resid_sumsq_dev) {
        sum_final_kernel<<<1, 256, 0, c->stream>>>(part, parts, resid_sumsq_dev);
        c->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

}  // namespace itcpd
