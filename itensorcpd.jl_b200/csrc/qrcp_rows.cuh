// Per-right-hand-side half of the rank-deficient fallback (xGELSY steps 4: dormqr, dtrsm, dormrz; see qrcp.cu):
// one thread owns one right-hand side.  Shared by qrcp_solve_rows_kernel (qrcp.cu) and by the one-thread-per-row
// normal-equation solve kernel (solve.cu), which takes this path in place when the pivoted Cholesky failed.
#pragma once
#include "common.cuh"

namespace itcpd {

struct QrcpWs {
    double *ws;     // A[m*n] | tau[n] | tauz[n] | ... (qrcp.cu: workspace layout)
    int *jpvt;
    double *bglob;  // rows x m scratch, right-hand side i strided by `rows`
};

// b <- Q^T b; solve T11; zero tail; apply Z^T; scatter through jpvt.  M: rows x m (row i = right-hand side i).
__device__ __forceinline__ void qrcp_row_solve(const double *__restrict__ ws, const int *__restrict__ jpvt, int rnk,
                                               const double *__restrict__ M, int64_t rows, int m, int n, double *__restrict__ X,
                                               double *__restrict__ bglob, int64_t i) {
    const double *A = ws, *tau = A + (size_t)m * n, *tauz = tau + n;
    double *b = bglob + i;
#define BV(k) b[(int64_t)(k) * rows]
    for (int k = 0; k < m; ++k) BV(k) = M[i + rows * (int64_t)k];
    // Q^T b = H_{n-1} ... H_0 b applied in order 0..n-1
    for (int j = 0; j < n; ++j) {
        double dot = BV(j);
        for (int q = j + 1; q < m; ++q) dot = fma(A[q + (size_t)m * j], BV(q), dot);
        const double f = tau[j] * dot;
        BV(j) -= f;
        for (int q = j + 1; q < m; ++q) BV(q) = fma(-f, A[q + (size_t)m * j], BV(q));
    }
    // T11 y = (Q^T b)(0:rnk)
    for (int k = rnk - 1; k >= 0; --k) {
        double s = BV(k);
        for (int q = k + 1; q < rnk; ++q) s = fma(-A[k + (size_t)m * q], BV(q), s);
        BV(k) = s / A[k + (size_t)m * k];
    }
    for (int k = rnk; k < n; ++k) BV(k) = 0.0;
    // Z^T y: Z = Z_0 Z_1 ... Z_{rnk-1}; Z^T y applies Z_{rnk-1}^T first ... LAPACK dormrz('L','T') loops i = 0..rnk-1
    const int l = n - rnk;
    if (l > 0) {
        for (int j = 0; j < rnk; ++j) {
            double dot = BV(j);
            for (int q = 0; q < l; ++q) dot = fma(A[j + (size_t)m * (rnk + q)], BV(rnk + q), dot);
            const double f = tauz[j] * dot;
            BV(j) -= f;
            for (int q = 0; q < l; ++q) BV(rnk + q) = fma(-f, A[j + (size_t)m * (rnk + q)], BV(rnk + q));
        }
    }
    for (int k = 0; k < n; ++k) X[i + rows * (int64_t)jpvt[k]] = BV(k);
#undef BV
}

int qrcp_workspace(itcpd_ctx *c, int m, int n, int64_t rows, QrcpWs *w);
int qrcp_factor_only(itcpd_ctx *c, const double *A, int m, int n, int64_t rows, int *status_dev, int force);
int qrcp_rows_only(itcpd_ctx *c, int m, int n, const double *Bt, int64_t rows, double *X, int *status_dev, int force);

}  // namespace itcpd
