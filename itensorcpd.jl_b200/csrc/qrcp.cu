// Rank-deficient fallback of ldiv_solve! (src/algebra/ldiv_solve.jl:19-21): `qr(A, ColumnNorm()) \ B`.
// Julia's ldiv!(::QRPivoted, B, rcond = n*eps) is the LAPACK xGELSY algorithm:
//   1. Householder QR with column pivoting                      (dgeqp3 / dlaqp2 norm down-dating)
//   2. rank by incremental condition estimation                 (dlaic1, stop when smax*rcond > smin)
//   3. RZ factorisation of the leading rank rows [R11 R12]      (dtzrzf / dlatrz)
//   4. x = P Z^T [T11^{-1} (Q^T b)(1:rank); 0]                    (dormqr, dtrsm, dormrz)
// Everything runs on the device: one CTA factorises the R x R matrix in global memory (rare path:
// it only runs when the pivoted Cholesky met a pivot <= tol), then one thread per right-hand side
// applies Q^T, the triangular solve and Z^T.  The kernels exit immediately unless status[0] == QRCP.
#include "common.cuh"
#include "qrcp_rows.cuh"
#include <cfloat>

namespace itcpd {

constexpr int QT = 512;

__device__ __forceinline__ double blk_sum(double v, double *sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < QT / 32; ++w) s += sh[w];  // every thread gets the total, fixed order
    return s;
}

// LAPACK dlaic1 (incremental condition estimation), restated from its published formulas.
// job 1: largest singular value, job 2: smallest.  Returns sestpr, s, c.
__device__ void laic1(int job, int j, const double *x, double sest, const double *w, double gamma, double *sestpr, double *s_out,
                      double *c_out) {
    const double eps = DBL_EPSILON * 0.5;  // LAPACK dlamch('Epsilon') = relative machine eps (2^-53)
    double alpha = 0.0;
    for (int i = 0; i < j; ++i) alpha += x[i] * w[i];
    const double absalp = fabs(alpha), absgam = fabs(gamma), absest = fabs(sest);
    double s, c, tmp, s1, s2, b, t, zeta1, zeta2, norma, cc, test, sine, cosine;
    if (job == 1) {
        if (sest == 0.0) {
            s1 = fmax(absgam, absalp);
            if (s1 == 0.0) { s = 0.0; c = 1.0; *sestpr = 0.0; }
            else { s = alpha / s1; c = gamma / s1; tmp = sqrt(s * s + c * c); s /= tmp; c /= tmp; *sestpr = s1 * tmp; }
        } else if (absgam <= eps * absest) {
            s = 1.0; c = 0.0; tmp = fmax(absest, absalp); s1 = absest / tmp; s2 = absalp / tmp;
            *sestpr = tmp * sqrt(s1 * s1 + s2 * s2);
        } else if (absalp <= eps * absest) {
            s1 = absgam; s2 = absest;
            if (s1 <= s2) { s = 1.0; c = 0.0; *sestpr = s2; } else { s = 0.0; c = 1.0; *sestpr = s1; }
        } else if (absest <= eps * absalp || absest <= eps * absgam) {
            s1 = absgam; s2 = absalp;
            if (s1 <= s2) { tmp = s1 / s2; s = sqrt(1.0 + tmp * tmp); *sestpr = s2 * s; c = (gamma / s2) / s; s = copysign(1.0, alpha) / s; }
            else { tmp = s2 / s1; c = sqrt(1.0 + tmp * tmp); *sestpr = s1 * c; s = (alpha / s1) / c; c = copysign(1.0, gamma) / c; }
        } else {
            zeta1 = alpha / absest; zeta2 = gamma / absest;
            b = (1.0 - zeta1 * zeta1 - zeta2 * zeta2) * 0.5;
            cc = zeta1 * zeta1;
            if (b > 0.0) t = cc / (b + sqrt(b * b + cc)); else t = sqrt(b * b + cc) - b;
            sine = -zeta1 / t; cosine = -zeta2 / (1.0 + t);
            tmp = sqrt(sine * sine + cosine * cosine);
            s = sine / tmp; c = cosine / tmp;
            *sestpr = sqrt(t + 1.0) * absest;
        }
    } else {
        if (sest == 0.0) {
            *sestpr = 0.0;
            if (fmax(absgam, absalp) == 0.0) { sine = 1.0; cosine = 0.0; } else { sine = -gamma; cosine = alpha; }
            s1 = fmax(fabs(sine), fabs(cosine));
            s = sine / s1; c = cosine / s1; tmp = sqrt(s * s + c * c); s /= tmp; c /= tmp;
        } else if (absgam <= eps * absest) {
            s = 0.0; c = 1.0; *sestpr = absgam;
        } else if (absalp <= eps * absest) {
            s1 = absgam; s2 = absest;
            if (s1 <= s2) { s = 0.0; c = 1.0; *sestpr = s1; } else { s = 1.0; c = 0.0; *sestpr = s2; }
        } else if (absest <= eps * absalp || absest <= eps * absgam) {
            s1 = absgam; s2 = absalp;
            if (s1 <= s2) { tmp = s1 / s2; c = sqrt(1.0 + tmp * tmp); *sestpr = absest * (tmp / c); s = -(gamma / s2) / c; c = copysign(1.0, alpha) / c; }
            else { tmp = s2 / s1; s = sqrt(1.0 + tmp * tmp); *sestpr = absest / s; c = (alpha / s1) / s; s = -copysign(1.0, gamma) / s; }
        } else {
            zeta1 = alpha / absest; zeta2 = gamma / absest;
            norma = fmax(1.0 + zeta1 * zeta1 + fabs(zeta1 * zeta2), fabs(zeta1 * zeta2) + zeta2 * zeta2);
            test = 1.0 + 2.0 * (zeta1 - zeta2) * (zeta1 + zeta2);
            if (test >= 0.0) {
                b = (zeta1 * zeta1 + zeta2 * zeta2 + 1.0) * 0.5;
                cc = zeta2 * zeta2;
                t = cc / (b + sqrt(fabs(b * b - cc)));
                sine = zeta1 / (1.0 - t); cosine = -zeta2 / t;
                *sestpr = sqrt(t + 4.0 * eps * eps * norma) * absest;
            } else {
                b = (zeta2 * zeta2 + zeta1 * zeta1 - 1.0) * 0.5;
                cc = zeta1 * zeta1;
                if (b >= 0.0) t = -cc / (b + sqrt(b * b + cc)); else t = b - sqrt(b * b + cc);
                sine = -zeta1 / t; cosine = -zeta2 / (1.0 + t);
                *sestpr = sqrt(1.0 + t + 4.0 * eps * eps * norma) * absest;
            }
            tmp = sqrt(sine * sine + cosine * cosine);
            s = sine / tmp; c = cosine / tmp;
        }
    }
    *s_out = s; *c_out = c;
}

// LAPACK dlarfg: Householder generator. x0 = alpha, x[1..m-1]; returns beta, tau, scales v in place (v0 = 1 implied)
// (block-parallel; all threads must call)
__device__ void larfg_block(int m, double *alpha, double *x, int incx, double *tau, double *sh) {
    double ss = 0.0;
    for (int i = threadIdx.x; i < m - 1; i += QT) { const double v = x[(size_t)i * incx]; ss = fma(v, v, ss); }
    const double xnorm = sqrt(blk_sum(ss, sh));
    const double a = *alpha;
    __syncthreads();
    if (xnorm == 0.0) { if (threadIdx.x == 0) *tau = 0.0; __syncthreads(); return; }
    const double beta = -copysign(hypot(a, xnorm), a);
    const double scal = 1.0 / (a - beta);
    for (int i = threadIdx.x; i < m - 1; i += QT) x[(size_t)i * incx] *= scal;
    if (threadIdx.x == 0) { *tau = (beta - a) / beta; *alpha = beta; }
    __syncthreads();
}

// workspace layout (doubles): A[n*n] | tau[n] | tauz[n] | vn1[n] | vn2[n] | wmin[n] | wmax[n] | wv[n]
__global__ void __launch_bounds__(QT) qrcp_factor_kernel(const double *__restrict__ Gin, int m, int n, double *__restrict__ ws, int *__restrict__ jpvt,
                                                         int *__restrict__ status, int force) {
    if (!force && status[0] != ITCPD_SOLVE_QRCP) return;
    __shared__ double sh[QT / 32];
    __shared__ double s_v[QT / 32];
    __shared__ int s_i[QT / 32];
    __shared__ int s_p;
    __shared__ int s_rank;
    double *A = ws, *tau = A + (size_t)m * n, *tauz = tau + n, *vn1 = tauz + n, *vn2 = vn1 + n, *wmin = vn2 + n, *wmax = wmin + n,
           *wv = wmax + n;
    const int tid = threadIdx.x;
    for (int64_t e = tid; e < (int64_t)m * n; e += QT) A[e] = Gin[e];
    for (int e = tid; e < n; e += QT) jpvt[e] = e;
    __syncthreads();
    // initial column norms
    for (int k = 0; k < n; ++k) {
        double ss = 0.0;
        for (int i = tid; i < m; i += QT) ss = fma(A[i + (size_t)m * k], A[i + (size_t)m * k], ss);
        const double nr = sqrt(blk_sum(ss, sh));
        if (tid == 0) { vn1[k] = nr; vn2[k] = nr; }
    }
    __syncthreads();
    const double tol3z = sqrt(DBL_EPSILON * 0.5);
    for (int j = 0; j < n; ++j) {
        // pivot = first max of vn1[j:]
        double bv = -1.0; int bi = n;
        for (int k = j + tid; k < n; k += QT) if (vn1[k] > bv) { bv = vn1[k]; bi = k; }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, bv, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { s_v[tid >> 5] = bv; s_i[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            double v = s_v[0]; int ix = s_i[0];
            for (int w = 1; w < QT / 32; ++w) if (s_v[w] > v || (s_v[w] == v && s_i[w] < ix)) { v = s_v[w]; ix = s_i[w]; }
            s_p = ix;
        }
        __syncthreads();
        const int p = s_p;
        if (p != j && p < n) {
            for (int i = tid; i < m; i += QT) { const double a = A[i + (size_t)m * j]; A[i + (size_t)m * j] = A[i + (size_t)m * p]; A[i + (size_t)m * p] = a; }
            if (tid == 0) { const int q = jpvt[j]; jpvt[j] = jpvt[p]; jpvt[p] = q; vn1[p] = vn1[j]; vn2[p] = vn2[j]; }
            __syncthreads();
        }
        // Householder on A[j:, j]
        if (j < m - 1) larfg_block(m - j, &A[j + (size_t)m * j], &A[j + 1 + (size_t)m * j], 1, &tau[j], sh);
        else { if (tid == 0) tau[j] = 0.0; __syncthreads(); }
        const double tj = tau[j];
        // apply H to the trailing columns and down-date the norms (dlaqp2)
        for (int k = j + 1; k < n; ++k) {
            double dot = 0.0;
            for (int i = j + 1 + tid; i < m; i += QT) dot = fma(A[i + (size_t)m * j], A[i + (size_t)m * k], dot);
            dot = blk_sum(dot, sh) + A[j + (size_t)m * k];
            const double f = tj * dot;
            __syncthreads();
            if (tid == 0) A[j + (size_t)m * k] -= f;
            for (int i = j + 1 + tid; i < m; i += QT) A[i + (size_t)m * k] = fma(-f, A[i + (size_t)m * j], A[i + (size_t)m * k]);
            __syncthreads();
            if (vn1[k] != 0.0) {
                double temp = fabs(A[j + (size_t)m * k]) / vn1[k];
                temp = fmax(0.0, 1.0 - temp * temp);
                const double r = vn1[k] / vn2[k];
                const double temp2 = temp * r * r;
                if (temp2 <= tol3z) {
                    double ss = 0.0;
                    for (int i = j + 1 + tid; i < m; i += QT) ss = fma(A[i + (size_t)m * k], A[i + (size_t)m * k], ss);
                    const double nr = sqrt(blk_sum(ss, sh));
                    __syncthreads();
                    if (tid == 0) { vn1[k] = nr; vn2[k] = nr; }
                } else {
                    __syncthreads();
                    if (tid == 0) vn1[k] = vn1[k] * sqrt(temp);
                }
                __syncthreads();
            }
        }
    }
    // ---- rank by incremental condition estimation (Julia qr.jl ldiv! / dgelsy) ----
    if (tid == 0) {
        const double rcond = n * DBL_EPSILON;
        int rnk = 0;
        double smax = fabs(A[0]), smin = smax;
        if (smax != 0.0) {
            rnk = 1;
            wmin[0] = 1.0; wmax[0] = 1.0;
            while (rnk < n) {
                const int i = rnk;
                double sminpr, s1, c1, smaxpr, s2, c2;
                laic1(2, rnk, wmin, smin, &A[(size_t)m * i], A[i + (size_t)m * i], &sminpr, &s1, &c1);
                laic1(1, rnk, wmax, smax, &A[(size_t)m * i], A[i + (size_t)m * i], &smaxpr, &s2, &c2);
                if (smaxpr * rcond > sminpr) break;
                for (int q = 0; q < rnk; ++q) { wmin[q] *= s1; wmax[q] *= s2; }
                wmin[rnk] = c1; wmax[rnk] = c2;
                smin = sminpr; smax = smaxpr;
                rnk++;
            }
        }
        s_rank = rnk;
        status[1] = rnk;
    }
    __syncthreads();
    const int rnk = s_rank;
    // ---- RZ factorisation of A[0:rnk, :] = [T11 0] Z  (dlatrz), l = n - rnk trailing columns ----
    const int l = n - rnk;
    if (l > 0) {
        for (int i = rnk - 1; i >= 0; --i) {
            // generate reflector from [A(i,i), A(i, rnk:n)]
            larfg_block(l + 1, &A[i + (size_t)m * i], &A[i + (size_t)m * rnk], m, &tauz[i], sh);
            const double tz = tauz[i];
            // apply to rows 0..i-1 from the right: w = A(0:i,i) + A(0:i, rnk:n) v ; A(:,i) -= tz w ; A(:,rnk:n) -= tz w v^T
            for (int r0 = tid; r0 < i; r0 += QT) {
                double w = A[r0 + (size_t)m * i];
                for (int q = 0; q < l; ++q) w = fma(A[r0 + (size_t)m * (rnk + q)], A[i + (size_t)m * (rnk + q)], w);
                wv[r0] = w;
            }
            __syncthreads();
            for (int r0 = tid; r0 < i; r0 += QT) {
                const double w = tz * wv[r0];
                A[r0 + (size_t)m * i] -= w;
                for (int q = 0; q < l; ++q) A[r0 + (size_t)m * (rnk + q)] = fma(-w, A[i + (size_t)m * (rnk + q)], A[r0 + (size_t)m * (rnk + q)]);
            }
            __syncthreads();
        }
    }
}

// one thread per right-hand side: b <- Q^T b; solve T11; zero tail; apply Z^T; scatter through jpvt (qrcp_rows.cuh)
__global__ void __launch_bounds__(64) qrcp_solve_rows_kernel(const double *__restrict__ ws, const int *__restrict__ jpvt,
                                                             const int *__restrict__ status, const double *__restrict__ M, int64_t rows,
                                                             int m, int n, double *__restrict__ X, double *__restrict__ bglob, int force) {
    if (!force && status[0] != ITCPD_SOLVE_QRCP) return;
    const int64_t i = blockIdx.x * 64ll + threadIdx.x;
    if (i >= rows) return;
    qrcp_row_solve(ws, jpvt, status[1], M, rows, m, n, X, bglob, i);
}

// workspace of the fallback: the factor behind the Cholesky factor in solve_ws, the pivots behind the Cholesky pivots
int qrcp_workspace(itcpd_ctx *c, int m, int n, int64_t rows, QrcpWs *w) {
    const size_t chol_doubles = (size_t)(n | 1) * n + n;            // the Cholesky factor lives in front (solve.cu)
    const size_t ws_doubles = (size_t)m * n + 8 * (size_t)n;
    TRY(c->solve_ws.reserve((chol_doubles + ws_doubles) * 8 + 1024));
    TRY(c->ipiv.reserve((size_t)n * 4 * 2));
    if (rows > 0) TRY(c->work.reserve((size_t)rows * m * 8));
    w->ws = c->solve_ws.as<double>() + chol_doubles;
    w->jpvt = c->ipiv.as<int>() + n;
    w->bglob = c->work.as<double>();
    return ITCPD_OK;
}

// the factorisation half alone: depends on A (and the status word) only, so the sweep driver runs it right behind the
// pivoted Cholesky on the side stream; it exits immediately unless the Cholesky failed (or `force`)
int qrcp_factor_only(itcpd_ctx *c, const double *A, int m, int n, int64_t rows, int *status_dev, int force) {
    ARG_CHECK(m >= n && n >= 1, "pivoted-QR least squares needs a tall (m >= n) matrix");
    QrcpWs w;
    TRY(qrcp_workspace(c, m, n, rows, &w));
    qrcp_factor_kernel<<<1, QT, 0, c->stream>>>(A, m, n, w.ws, w.jpvt, status_dev, force);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

int qrcp_rows_only(itcpd_ctx *c, int m, int n, const double *Bt, int64_t rows, double *X, int *status_dev, int force) {
    QrcpWs w;
    TRY(qrcp_workspace(c, m, n, rows, &w));
    qrcp_solve_rows_kernel<<<(unsigned)ceil_div(rows, 64), 64, 0, c->stream>>>(w.ws, w.jpvt, status_dev, Bt, rows, m, n, X, w.bglob, force);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// min-norm least squares  X (rows x n) = (A \\ B^T)^T  for a tall column-major A (m x n, m >= n) and the rows x m
// matrix Bt whose row i is right-hand side i.  `force` = 0: run only when status[0] says the Cholesky failed.
int qrcp_ls_solve(itcpd_ctx *c, const double *A, int m, int n, const double *Bt, int64_t rows, double *X, int *status_dev, int force) {
    TRY(qrcp_factor_only(c, A, m, n, rows, status_dev, force));
    return qrcp_rows_only(c, m, n, Bt, rows, X, status_dev, force);
}

int qrcp_minnorm_solve(itcpd_ctx *c, const double *Gamma, const double *M, int64_t rows, int R, double *X, int *status_dev) {
    return qrcp_ls_solve(c, Gamma, R, R, M, rows, X, status_dev, 0);
}

}  // namespace itcpd
