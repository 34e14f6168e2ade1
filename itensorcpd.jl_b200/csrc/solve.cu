// R x R normal-equation solve (src/algebra/ldiv_solve.jl:13-29):
//   cholesky(Hermitian(Gamma), RowMaximum(), check=true, tol) \ B      -> LAPACK dpstrf + permuted dpotrs
//   on failure: qr(Gamma, ColumnNorm()) \ B                            -> xGELSY min-norm (see qrcp below)
// Latency-bound block-level kernels: one CTA factorises (matrix resident in shared memory when it
// fits), then one thread per right-hand side runs the two triangular solves against the factor.
#include "common.cuh"
#include <cfloat>

namespace itcpd {

constexpr int CH_THREADS = 512;

// status words written by the factorisation: [0] path (0 chol / 1 qrcp), [1] rank, [2] info
__global__ void __launch_bounds__(CH_THREADS) pivoted_cholesky_kernel(const double *__restrict__ Gin, int n, double tol,
                                                                      double *__restrict__ Wg, int *__restrict__ piv,
                                                                      int *__restrict__ status, int use_smem) {
    extern __shared__ double sm_dyn[];
    __shared__ double red_v[CH_THREADS / 32];
    __shared__ int red_i[CH_THREADS / 32];
    __shared__ int s_p;
    __shared__ double s_ajj, s_stop;
    __shared__ int s_fail;
    double *W = use_smem ? sm_dyn : Wg;
    const int tid = threadIdx.x;
    for (int e = tid; e < n * n; e += CH_THREADS) W[e] = Gin[e];
    for (int e = tid; e < n; e += CH_THREADS) piv[e] = e;
    if (tid == 0) s_fail = 0;
    __syncthreads();

    int rank = n;
    for (int j = 0; j < n; ++j) {
        // ---- pivot: first maximum of the remaining (updated) diagonal, dpstf2 semantics ----
        double bv = -DBL_MAX;
        int bi = n;
        bool has_nan = false;
        for (int i = j + tid; i < n; i += CH_THREADS) {
            const double d = W[i + n * i];
            if (d != d) has_nan = true;
            if (d > bv) { bv = d; bi = i; }
        }
        if (has_nan) { bv = DBL_MAX; bi = -1; }  // NaN poisons the factorisation (LAPACK: disnan(ajj) -> fail)
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, bv, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { red_v[tid >> 5] = bv; red_i[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            double v = red_v[0];
            int ix = red_i[0];
            for (int w = 1; w < CH_THREADS / 32; ++w)
                if (red_v[w] > v || (red_v[w] == v && red_i[w] < ix)) { v = red_v[w]; ix = red_i[w]; }
            s_p = ix;
            s_ajj = v;
            if (j == 0) {
                s_stop = (tol < 0.0) ? n * DBL_EPSILON * v : tol;
                if (ix < 0 || !(v > 0.0)) s_fail = 1;            // dpstf2: ajj <= 0 or NaN at the start
            } else if (ix < 0 || !(v > s_stop)) {
                s_fail = 1;                                       // pivot <= tol: rank deficient
            }
        }
        __syncthreads();
        if (s_fail) { rank = j; break; }
        const int p = s_p;
        // ---- symmetric interchange j <-> p on the full square ----
        if (p != j) {
            for (int k = tid; k < n; k += CH_THREADS) {  // rows
                const double a = W[j + n * k], b = W[p + n * k];
                W[j + n * k] = b; W[p + n * k] = a;
            }
            __syncthreads();
            for (int k = tid; k < n; k += CH_THREADS) {  // columns
                const double a = W[k + n * j], b = W[k + n * p];
                W[k + n * j] = b; W[k + n * p] = a;
            }
            if (tid == 0) { const int q = piv[j]; piv[j] = piv[p]; piv[p] = q; }
            __syncthreads();
        }
        const double d = sqrt(s_ajj);
        // ---- scale row j of the factor ----
        for (int k = j + 1 + tid; k < n; k += CH_THREADS) W[j + n * k] = W[j + n * k] / d;
        if (tid == 0) W[j + n * j] = d;
        __syncthreads();
        // ---- trailing update on the full square (keeps both triangles so interchanges stay trivial) ----
        const int m = n - j - 1;
        for (int e = tid; e < m * m; e += CH_THREADS) {
            const int i = j + 1 + e % m, k = j + 1 + e / m;
            W[i + n * k] = fma(-W[j + n * i], W[j + n * k], W[i + n * k]);
        }
        __syncthreads();
    }
    if (use_smem)
        for (int e = tid; e < n * n; e += CH_THREADS) Wg[e] = W[e];
    if (tid == 0) { status[0] = (rank == n) ? ITCPD_SOLVE_CHOLESKY : ITCPD_SOLVE_QRCP; status[1] = rank; status[2] = (rank == n) ? 0 : 1; }
}

// One thread per right-hand side (a row of M): x = P (U^T U)^{-1} P^T b, written to row `i` of X.
// U (upper triangle of Wg, column-major) is staged in shared memory when it fits.
constexpr int TS_THREADS = 64;

__global__ void __launch_bounds__(TS_THREADS) chol_solve_rows_kernel(const double *__restrict__ Wg, const int *__restrict__ piv,
                                                                     const int *__restrict__ status, const double *__restrict__ M,
                                                                     int64_t rows, int n, double *__restrict__ X, int u_in_smem,
                                                                     int b_in_smem, double *__restrict__ bglob, int fwd_only_rank) {
    extern __shared__ double sm_dyn[];
    if (fwd_only_rank < 0 && status[0] != ITCPD_SOLVE_CHOLESKY) return;  // the QRCP path handles this system
    const double *U = Wg;
    double *sb = sm_dyn;
    if (u_in_smem) {
        for (int e = threadIdx.x; e < n * n; e += TS_THREADS) sm_dyn[e] = Wg[e];
        U = sm_dyn;
        sb = sm_dyn + (size_t)n * n;
        __syncthreads();
    }
    const int64_t i = blockIdx.x * (int64_t)TS_THREADS + threadIdx.x;
    if (i >= rows) return;
    // element k of this thread's vector
    double *bp;
    int64_t bs;
    if (b_in_smem) { bp = sb + threadIdx.x; bs = TS_THREADS; }
    else { bp = bglob + i; bs = rows; }
#define BV(k) bp[(int64_t)(k) * bs]
    const int nn = (fwd_only_rank >= 0) ? fwd_only_rank : n;
    for (int k = 0; k < n; ++k) BV(k) = M[i + rows * (int64_t)piv[k]];
    // forward: U^T y = P^T b
    for (int k = 0; k < nn; ++k) {
        double s0 = BV(k), s1 = 0.0;
        const double *uk = U + (size_t)n * k;
        int l = 0;
        for (; l + 1 < k; l += 2) { s0 = fma(-uk[l], BV(l), s0); s1 = fma(-uk[l + 1], BV(l + 1), s1); }
        if (l < k) s0 = fma(-uk[l], BV(l), s0);
        BV(k) = (s0 + s1) / uk[k];
    }
    if (fwd_only_rank >= 0) {  // leverage score: ||y||^2 (k_leverage)
        double s = 0.0;
        for (int k = 0; k < nn; ++k) s = fma(BV(k), BV(k), s);
        X[i] = s;
        return;
    }
    // backward: U x = y
    for (int k = n - 1; k >= 0; --k) {
        double s0 = BV(k), s1 = 0.0;
        int l = k + 1;
        for (; l + 1 < n; l += 2) { s0 = fma(-U[k + (size_t)n * l], BV(l), s0); s1 = fma(-U[k + (size_t)n * (l + 1)], BV(l + 1), s1); }
        if (l < n) s0 = fma(-U[k + (size_t)n * l], BV(l), s0);
        BV(k) = (s0 + s1) / U[k + (size_t)n * k];
    }
    for (int k = 0; k < n; ++k) X[i + rows * (int64_t)piv[k]] = BV(k);
#undef BV
}

static int smem_limit(itcpd_ctx *) { return 227 * 1024 - 2048; }

static int run_cholesky(itcpd_ctx *c, const double *Gamma, int R, double tol, int *status_dev) {
    TRY(c->solve_ws.reserve((size_t)R * R * 8 * 2 + 1024));
    TRY(c->ipiv.reserve((size_t)R * 4 * 2));
    const size_t need = (size_t)R * R * 8;
    const int use_smem = need <= (size_t)smem_limit(c);
    static bool attr = false;
    if (!attr) {
        CUDA_TRY(cudaFuncSetAttribute(pivoted_cholesky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit(c)));
        CUDA_TRY(cudaFuncSetAttribute(chol_solve_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit(c)));
        attr = true;
    }
    pivoted_cholesky_kernel<<<1, CH_THREADS, use_smem ? need : 0, c->stream>>>(Gamma, R, tol, c->solve_ws.as<double>(), c->ipiv.as<int>(),
                                                                                 status_dev, use_smem);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

static int run_tri_solves(itcpd_ctx *c, const double *M, int64_t rows, int R, double *X, const int *status_dev, int fwd_only_rank) {
    const size_t u_bytes = (size_t)R * R * 8, b_bytes = (size_t)R * TS_THREADS * 8;
    int u_in = 0, b_in = 0;
    size_t smem = 0;
    if (u_bytes + b_bytes <= (size_t)smem_limit(c)) { u_in = 1; b_in = 1; smem = u_bytes + b_bytes; }
    else if (b_bytes <= (size_t)smem_limit(c)) { b_in = 1; smem = b_bytes; }
    double *bglob = nullptr;
    if (!b_in) {
        TRY(c->work.reserve((size_t)rows * R * 8));
        bglob = c->work.as<double>();
    }
    chol_solve_rows_kernel<<<(unsigned)ceil_div(rows, TS_THREADS), TS_THREADS, smem, c->stream>>>(
        c->solve_ws.as<double>(), c->ipiv.as<int>(), status_dev, M, rows, R, X, u_in, b_in, bglob, fwd_only_rank);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

int qrcp_minnorm_solve(itcpd_ctx *c, const double *Gamma, const double *M, int64_t rows, int R, double *X, int *status_dev);  // qrcp.cu

int k_solve(itcpd_ctx *c, const double *Gamma, const double *M, int64_t rows, int R, double tol, double *X, int *status_dev) {
    TRY(run_cholesky(c, Gamma, R, tol, status_dev));
    TRY(run_tri_solves(c, M, rows, R, X, status_dev, -1));
    // rank-deficient systems are re-solved by the pivoted-QR min-norm path; it is a no-op (device-side
    // early exit on status[0]) when the Cholesky succeeded, so no host round trip is needed here.
    TRY(qrcp_minnorm_solve(c, Gamma, M, rows, R, X, status_dev));
    return ITCPD_OK;
}

// leverage scores (math_tools/probability.jl:3-10): p_i = ||Q[i,:]||^2 / min(I,R) with A = QR.
// Q = A U^{-1} for the Cholesky factor of the Gram matrix, so ||Q[i,:]||^2 = ||U^{-T} a_i||^2.
__global__ void fill_kernel(double *x, int64_t n, double v) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}
__global__ void scale_kernel(double *x, int64_t n, double v) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) x[i] *= v;
}

int k_leverage(itcpd_ctx *c, const double *A, const double *G, int64_t rows, int R, double *lev_out) {
    if (rows <= R) {  // square/wide factor: Q is orthogonal, every row has unit norm
        fill_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, c->stream>>>(lev_out, rows, 1.0 / (double)rows);
        c->launches++;
        return ITCPD_OK;
    }
    TRY(c->status.reserve(64));
    int *st = c->status.as<int>() + 8;
    TRY(run_cholesky(c, G, R, -1.0, st));
    int h[3];
    CUDA_TRY(cudaMemcpyAsync(h, st, 12, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    TRY(run_tri_solves(c, A, rows, R, lev_out, st, h[1]));
    scale_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, c->stream>>>(lev_out, rows, 1.0 / (double)std::min<int64_t>(rows, R));
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

}  // namespace itcpd
