// R x R normal-equation solve (src/algebra/ldiv_solve.jl:13-29):
//   cholesky(Hermitian(Gamma), RowMaximum(), check=true, tol) \ B      -> LAPACK dpstrf + permuted dpotrs
//   on failure: qr(Gamma, ColumnNorm()) \ B                            -> xGELSY min-norm (qrcp.cu)
// Latency-bound block/warp-level kernels:
//   * pivoted_cholesky_kernel: one CTA, matrix resident in shared memory (odd leading dimension, no bank
//     conflicts), left-looking dpstf2 order: 3 block barriers per column.
//   * chol_solve_warp_kernel: one WARP per right-hand side (a row of M); the vector lives in registers
//     (lane l owns entries l, l+32, ...), each substitution step is one shuffle broadcast + E FMAs.
#include "common.cuh"
#include "qrcp_rows.cuh"
#include <cfloat>

namespace itcpd {

// fused all-reduce, first half: wait until every peer has published its partial M for this exchange (system-scope flags
// written by the peers' signal kernels into OUR memory).  The epoch is a kernel argument (host-counted exchanges) or, for
// sweeps that are replayed from a CUDA graph, a word in device memory that the signal kernel advanced (peer_graph.cu).
// Bounded: a peer that died must surface as a launch failure on this rank, not as a hung box.  All threads must call.
__device__ __forceinline__ void peer_wait_all(const PeerSrc &src) {
    if (!src.flags) return;
    if ((int)threadIdx.x < src.n) {
        const long long epoch = src.epoch_dev ? *src.epoch_dev : src.epoch;
        unsigned long long t0 = 0, spins = 0;
        while (src.flags[threadIdx.x] < epoch) {
            if ((++spins & 0xfffff) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > ITCPD_PEER_TIMEOUT_NS) {
                    if (blockIdx.x == 0) printf("itcpd: peer %d never published exchange %lld\n", (int)threadIdx.x, epoch);
                    __trap();
                }
            }
        }
    }
    __syncthreads();
    __threadfence_system();
}

constexpr int CH_THREADS = 256;

// status words written by the factorisation: [0] path (0 chol / 1 qrcp), [1] rank, [2] info
// Wg receives the factor with leading dimension ldw = n | 1.
__global__ void __launch_bounds__(CH_THREADS) pivoted_cholesky_kernel(const double *__restrict__ Gin, int n, double tol,
                                                                      double *__restrict__ Wg, int *__restrict__ piv,
                                                                      int *__restrict__ status, int use_smem) {
    extern __shared__ double sm_dyn[];
    __shared__ int s_p, s_fail;
    __shared__ double s_ajj, s_stop;
    __shared__ int s_piv[1024];
    const int ldw = n | 1;
    const int NT = blockDim.x;
    double *W = use_smem ? sm_dyn : Wg;                       // full symmetric copy; rows < j become the factor
    double *dd = use_smem ? sm_dyn + (size_t)ldw * n : Wg + (size_t)ldw * n;  // running diagonal (dpstf2 "work")
    const int tid = threadIdx.x;
    for (int e = tid; e < n * n; e += NT) W[(e % n) + (size_t)ldw * (e / n)] = Gin[e];
    for (int e = tid; e < n; e += NT) { s_piv[e] = e; dd[e] = Gin[e + (size_t)n * e]; }
    if (tid == 0) s_fail = 0;
    __syncthreads();

    int rank = n;
    for (int j = 0; j < n; ++j) {
        // ---- pivot: first maximum of the running diagonal (warp 0 only) ----
        if (tid < 32) {
            double bv = -DBL_MAX;
            int bi = n;
            bool bad = false;
            for (int i = j + tid; i < n; i += 32) {
                const double d = dd[i];
                if (d != d) bad = true;
                if (d > bv) { bv = d; bi = i; }
            }
            if (bad) { bv = DBL_MAX; bi = -1; }  // NaN poisons the factorisation (LAPACK: disnan -> fail)
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, bv, o);
                const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (tid == 0) {
                s_p = bi;
                s_ajj = bv;
                if (j == 0) {
                    s_stop = (tol < 0.0) ? n * DBL_EPSILON * bv : tol;
                    if (bi < 0 || !(bv > 0.0)) s_fail = 1;   // dpstf2: ajj <= 0 or NaN at the start
                } else if (bi < 0 || !(bv > s_stop)) {
                    s_fail = 1;                               // pivot <= tol: rank deficient
                }
            }
        }
        __syncthreads();
        if (s_fail) { rank = j; break; }
        const int p = s_p;
        // ---- symmetric interchange j <-> p (single phase: off-diagonal rows and columns, then the 2x2 block) ----
        if (p != j) {
            for (int t = tid; t < n; t += NT) {
                if (t != j && t != p) {
                    double a = W[j + (size_t)ldw * t], b = W[p + (size_t)ldw * t];
                    W[j + (size_t)ldw * t] = b; W[p + (size_t)ldw * t] = a;
                    a = W[t + (size_t)ldw * j]; b = W[t + (size_t)ldw * p];
                    W[t + (size_t)ldw * j] = b; W[t + (size_t)ldw * p] = a;
                }
            }
            if (tid == 0) {
                const double a = W[j + (size_t)ldw * j];
                W[j + (size_t)ldw * j] = W[p + (size_t)ldw * p];
                W[p + (size_t)ldw * p] = a;
                const double b = W[j + (size_t)ldw * p];
                W[j + (size_t)ldw * p] = W[p + (size_t)ldw * j];
                W[p + (size_t)ldw * j] = b;
                const double d = dd[j]; dd[j] = dd[p]; dd[p] = d;
                const int q = s_piv[j]; s_piv[j] = s_piv[p]; s_piv[p] = q;
            }
            __syncthreads();
        }
        // ---- row j of the factor: U[j,k] = (A[j,k] - sum_{l<j} U[l,j] U[l,k]) / U[j,j]; diagonal down-date ----
        const double d = sqrt(s_ajj);
        for (int k = j + 1 + tid; k < n; k += NT) {
            const double *cj = W + (size_t)ldw * j, *ck = W + (size_t)ldw * k;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int l = 0;
            for (; l + 3 < j; l += 4) {
                s0 = fma(cj[l], ck[l], s0); s1 = fma(cj[l + 1], ck[l + 1], s1);
                s2 = fma(cj[l + 2], ck[l + 2], s2); s3 = fma(cj[l + 3], ck[l + 3], s3);
            }
            for (; l < j; ++l) s0 = fma(cj[l], ck[l], s0);
            const double u = (ck[j] - ((s0 + s1) + (s2 + s3))) / d;
            W[j + (size_t)ldw * k] = u;
            dd[k] = fma(-u, u, dd[k]);
        }
        if (tid == 0) W[j + (size_t)ldw * j] = d;
        __syncthreads();
    }
    if (use_smem)
        for (int e = tid; e < ldw * n; e += NT) Wg[e] = W[e];
    for (int e = tid; e < n; e += NT) piv[e] = s_piv[e];
    if (tid == 0) { status[0] = (rank == n) ? ITCPD_SOLVE_CHOLESKY : ITCPD_SOLVE_QRCP; status[1] = rank; status[2] = (rank == n) ? 0 : 1; }
}

// ---- latency-tuned variant for n <= 128 ("team" kernel) --------------------------------------------------------
// Same arithmetic, operation for operation, as pivoted_cholesky_kernel (bitwise identical factor, pivots and status),
// but organised around the dependent chain of one column: thread k owns column k and keeps its running diagonal in
// a register; the pivot search is three warp redux ops on an order-preserving integer key (no FP64 compares, no
// shuffle tree) plus one named barrier across the <= 4 owning warps; the swap and the row computation share a second
// barrier (a single warp needs only __syncwarp).  All 256 threads of the CTA take part in the load / store phases.
constexpr int CHT_MAX_N = 128;
constexpr int CHT_THREADS = 256;

__device__ __forceinline__ unsigned long long ordered_key(double d) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ void team_sync(int nw, int team) {
    if (nw == 1) __syncwarp();
    else asm volatile("bar.sync 1, %0;" ::"r"(team) : "memory");
}

__global__ void __launch_bounds__(CHT_THREADS) pivoted_cholesky_team_kernel(const double *__restrict__ Gin, int n, double tol,
                                                                            double *__restrict__ Wg, int *__restrict__ piv,
                                                                            int *__restrict__ status) {
    extern __shared__ double sm_dyn[];
    __shared__ unsigned long long s_key[4];
    __shared__ int s_idx[4], s_bad[4];
    __shared__ double s_swapdd;
    __shared__ int s_piv[CHT_MAX_N];
    __shared__ int s_rank;
    const int ldw = n | 1;
    const int tid = threadIdx.x;
    const int team = (n + 31) & ~31, nw = team >> 5;
    double *W = sm_dyn;
    for (int e = tid; e < n * n; e += CHT_THREADS) W[(e % n) + (size_t)ldw * (e / n)] = Gin[e];
    if (tid < n) s_piv[tid] = tid;
    if (tid == 0) s_rank = n;
    __syncthreads();
    if (tid < team) {
        const int k = tid, lane = tid & 31, w = tid >> 5;
        const bool mine = k < n;
        double *ck = W + (size_t)ldw * (mine ? k : 0);
        double ddk = mine ? ck[k] : 0.0;  // running diagonal (dpstf2 "work") of my column
        double stop = 0.0;
        int rank = n;
        for (int j = 0; j < n; ++j) {
            // ---- pivot: first maximum of the running diagonal over k >= j ----
            const bool elig = mine && k >= j;
            const unsigned long long key = elig ? ordered_key(ddk) : 0ull;
            const unsigned hi = (unsigned)(key >> 32);
            const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
            const unsigned lo = (hi == mh) ? (unsigned)key : 0u;
            const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
            const bool win = elig && hi == mh && (unsigned)key == ml;
            const unsigned mi = __reduce_min_sync(0xffffffffu, win ? (unsigned)k : 0x7fffffffu);
            int bb = __any_sync(0xffffffffu, elig && ddk != ddk);  // NaN poisons the factorisation
            unsigned long long bk = ((unsigned long long)mh << 32) | ml;
            int bi = (int)mi;
            if (nw > 1) {
                if (lane == 0) { s_key[w] = bk; s_idx[w] = bi; s_bad[w] = bb; }
                team_sync(nw, team);
                bk = s_key[0]; bi = s_idx[0]; bb = s_bad[0];
                for (int q = 1; q < nw; ++q) {
                    const unsigned long long kq = s_key[q];
                    if (kq > bk) { bk = kq; bi = s_idx[q]; }  // ties keep the lower warp = the lower index
                    bb |= s_bad[q];
                }
            }
            const double bv = key_value(bk);
            bool fail;
            if (j == 0) {
                stop = (tol < 0.0) ? n * DBL_EPSILON * bv : tol;
                fail = bb || !(bv > 0.0);
            } else {
                fail = bb || !(bv > stop);
            }
            if (fail) { rank = j; break; }  // uniform over the team
            const int p = bi;
            const double d = sqrt(bv);
            // ---- symmetric interchange j <-> p ----
            if (p != j && mine) {
                if (k != j && k != p) {
                    double a = ck[j], b = ck[p];
                    ck[j] = b; ck[p] = a;
                    a = W[k + (size_t)ldw * j]; b = W[k + (size_t)ldw * p];
                    W[k + (size_t)ldw * j] = b; W[k + (size_t)ldw * p] = a;
                } else if (k == j) {
                    const double a = W[j + (size_t)ldw * j];
                    W[j + (size_t)ldw * j] = W[p + (size_t)ldw * p];
                    W[p + (size_t)ldw * p] = a;
                    const double b = W[j + (size_t)ldw * p];
                    W[j + (size_t)ldw * p] = W[p + (size_t)ldw * j];
                    W[p + (size_t)ldw * j] = b;
                    const int q = s_piv[j]; s_piv[j] = s_piv[p]; s_piv[p] = q;
                    s_swapdd = ddk;  // my running diagonal moves to column p
                }
            }
            team_sync(nw, team);
            if (p != j && k == p) ddk = s_swapdd;
            // ---- row j of the factor and the diagonal down-date ----
            if (mine && k > j) {
                const double *cj = W + (size_t)ldw * j;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int l = 0;
                for (; l + 3 < j; l += 4) {
                    s0 = fma(cj[l], ck[l], s0); s1 = fma(cj[l + 1], ck[l + 1], s1);
                    s2 = fma(cj[l + 2], ck[l + 2], s2); s3 = fma(cj[l + 3], ck[l + 3], s3);
                }
                for (; l < j; ++l) s0 = fma(cj[l], ck[l], s0);
                const double u = (ck[j] - ((s0 + s1) + (s2 + s3))) / d;
                ck[j] = u;
                ddk = fma(-u, u, ddk);
            } else if (k == j) {
                ck[j] = d;
            }
            if (nw == 1) __syncwarp();  // (several warps: the next pivot barrier orders these writes before the next swap)
        }
        if (tid == 0) s_rank = rank;
    }
    __syncthreads();
    const int rank = s_rank;
    for (int e = tid; e < ldw * n; e += CHT_THREADS) Wg[e] = W[e];
    for (int e = tid; e < n; e += CHT_THREADS) piv[e] = s_piv[e];
    if (tid == 0) { status[0] = (rank == n) ? ITCPD_SOLVE_CHOLESKY : ITCPD_SOLVE_QRCP; status[1] = rank; status[2] = (rank == n) ? 0 : 1; }
}

// ---- right-looking variant for n <= 64 (chol_alg = 2, and chol_alg = 3 where the factorisation is exposed) -----
// The left-looking dot is what the team kernel's dependent chain is made of (profiles/r1_cholesky_probe.txt).  Here
// thread k owns ORIGINAL column k of the trailing matrix in registers for the whole factorisation and nothing is ever
// swapped: a step is  pivot search -> u_k = A[q,k] / d  (q = pivot column; a uniform dynamic register index, read
// through a jump table) -> publish u through shared memory -> rank-1 update of the register column (independent FMAs).
// Positions follow LAPACK's interchange bookkeeping (the column at position j moves to the pivot's position), so pivots,
// tie-breaking (first maximum in POSITION order) and the output layout equal the other kernels'; the rounding differs
// (sequential fma down-dates instead of a 4-way split dot), so parity is by tolerance, not bitwise.
template <int NMAX>
__device__ __forceinline__ double rl_fetch(const double (&col)[NMAX], int q) {
    double a = 0.0;
#define RL_CASE(i) case i: a = col[(i) < NMAX ? (i) : 0]; break;
    switch (q) {
        RL_CASE(0) RL_CASE(1) RL_CASE(2) RL_CASE(3) RL_CASE(4) RL_CASE(5) RL_CASE(6) RL_CASE(7)
        RL_CASE(8) RL_CASE(9) RL_CASE(10) RL_CASE(11) RL_CASE(12) RL_CASE(13) RL_CASE(14) RL_CASE(15)
        RL_CASE(16) RL_CASE(17) RL_CASE(18) RL_CASE(19) RL_CASE(20) RL_CASE(21) RL_CASE(22) RL_CASE(23)
        RL_CASE(24) RL_CASE(25) RL_CASE(26) RL_CASE(27) RL_CASE(28) RL_CASE(29) RL_CASE(30) RL_CASE(31)
        RL_CASE(32) RL_CASE(33) RL_CASE(34) RL_CASE(35) RL_CASE(36) RL_CASE(37) RL_CASE(38) RL_CASE(39)
        RL_CASE(40) RL_CASE(41) RL_CASE(42) RL_CASE(43) RL_CASE(44) RL_CASE(45) RL_CASE(46) RL_CASE(47)
        RL_CASE(48) RL_CASE(49) RL_CASE(50) RL_CASE(51) RL_CASE(52) RL_CASE(53) RL_CASE(54) RL_CASE(55)
        RL_CASE(56) RL_CASE(57) RL_CASE(58) RL_CASE(59) RL_CASE(60) RL_CASE(61) RL_CASE(62) RL_CASE(63)
        default: break;
    }
#undef RL_CASE
    return a;
}

template <int NMAX>
__global__ void __launch_bounds__(CHT_THREADS) pivoted_cholesky_rl_kernel(const double *__restrict__ Gin, int n, double tol,
                                                                          double *__restrict__ Wg, int *__restrict__ piv,
                                                                          int *__restrict__ status) {
    extern __shared__ double sm_dyn[];            // Gamma staging, then Uo[l + ldw * original column] = row l of the factor
    __shared__ __align__(16) double s_u[NMAX];    // u of the current step by original column; 0 once a column is eliminated
    __shared__ unsigned long long s_key[2];
    __shared__ int s_idx[2], s_bad[2];
    __shared__ int s_orig[NMAX];                  // position -> original column
    __shared__ int s_rank;
    const int ldw = n | 1;
    const int tid = threadIdx.x;
    constexpr int team = NMAX, nw = NMAX / 32;
    double *W = sm_dyn;
    for (int e = tid; e < n * n; e += CHT_THREADS) W[(e % n) + (size_t)ldw * (e / n)] = Gin[e];
    if (tid < NMAX) { s_orig[tid] = tid; s_u[tid] = 0.0; }
    if (tid == 0) s_rank = n;
    __syncthreads();
    if (tid < team) {
        const int k = tid, lane = tid & 31, w = tid >> 5;
        const bool mine = k < n;
        double col[NMAX];
#pragma unroll
        for (int i = 0; i < NMAX; ++i) col[i] = (mine && i < n) ? W[i + (size_t)ldw * k] : 0.0;
        double ddk = mine ? W[k + (size_t)ldw * k] : 0.0;
        bool alive = mine;
        int posk = k;
        double stop = 0.0;
        int rank = n;
        team_sync(nw, team);  // every column is in registers: the staging area becomes Uo
        for (int j = 0; j < n; ++j) {
            // ---- pivot: first maximum of the running diagonal in position order ----
            const unsigned long long key = alive ? ordered_key(ddk) : 0ull;
            const unsigned hi = (unsigned)(key >> 32);
            const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
            const unsigned lo = (hi == mh) ? (unsigned)key : 0u;
            const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
            const bool win = alive && hi == mh && (unsigned)key == ml;
            const unsigned mi = __reduce_min_sync(0xffffffffu, win ? (((unsigned)posk << 8) | (unsigned)k) : 0x7fffffffu);
            int bb = __any_sync(0xffffffffu, alive && ddk != ddk);
            unsigned long long bk = ((unsigned long long)mh << 32) | ml;
            int bi = (int)mi;
            if (nw > 1) {
                if (lane == 0) { s_key[w] = bk; s_idx[w] = bi; s_bad[w] = bb; }
                team_sync(nw, team);
                const unsigned long long k0 = s_key[0], k1 = s_key[1];
                const int i0 = s_idx[0], i1 = s_idx[1];
                if (k1 > k0 || (k1 == k0 && i1 < i0)) { bk = k1; bi = i1; } else { bk = k0; bi = i0; }
                bb = s_bad[0] | s_bad[1];
            }
            const double bv = key_value(bk);
            bool fail;
            if (j == 0) {
                stop = (tol < 0.0) ? n * DBL_EPSILON * bv : tol;
                fail = bb || !(bv > 0.0);
            } else {
                fail = bb || !(bv > stop);
            }
            if (fail) { rank = j; break; }  // uniform over the team
            const int q = bi & 0xff, p = bi >> 8;
            const double d = sqrt(bv);
            // ---- interchange bookkeeping only: the pivot goes to position j, the column that sat there to position p ----
            if (mine) {
                if (k == q) { posk = j; s_orig[j] = k; }
                else if (posk == j) { posk = p; s_orig[p] = k; }
            }
            // ---- row j of the factor ----
            double u = 0.0;
            if (alive) {
                if (k == q) {
                    W[j + (size_t)ldw * k] = d;
                    s_u[k] = 0.0;
                    alive = false;
                } else {
                    u = rl_fetch<NMAX>(col, q) / d;
                    W[j + (size_t)ldw * k] = u;
                    s_u[k] = u;
                    ddk = fma(-u, u, ddk);
                }
            }
            team_sync(nw, team);
            // ---- rank-1 update of my column (rows of eliminated columns see u = 0) ----
            if (alive) {
#pragma unroll
                for (int i = 0; i < NMAX; i += 2) {
                    const double2 uu = *reinterpret_cast<const double2 *>(&s_u[i]);
                    col[i] = fma(-uu.x, u, col[i]);
                    col[i + 1] = fma(-uu.y, u, col[i + 1]);
                }
            }
            if (nw == 1) __syncwarp();  // (two warps: the next pivot barrier orders these reads before the next s_u writes)
        }
        if (tid == 0) s_rank = rank;
    }
    __syncthreads();
    const int rank = s_rank;
    for (int e = tid; e < ldw * n; e += CHT_THREADS) {
        const int c = e / ldw, l = e - c * ldw;
        Wg[e] = (l <= c && l < rank) ? W[l + (size_t)ldw * s_orig[c]] : 0.0;
    }
    for (int e = tid; e < n; e += CHT_THREADS) piv[e] = s_orig[e];
    if (tid == 0) { status[0] = (rank == n) ? ITCPD_SOLVE_CHOLESKY : ITCPD_SOLVE_QRCP; status[1] = rank; status[2] = (rank == n) ? 0 : 1; }
}

// ---- right-looking variant for 64 < n <= 128 (chol_alg = 2 / 3) -------------------------------------------------------
// A 128-entry column does not fit one thread's registers, so every column is owned by TWO threads: thread (k, h) keeps rows
// 64 h .. 64 h + 63 of original column k (k = tid & 127, h = tid >> 7: 256 threads).  The pivot search runs on the h = 0 threads
// (they carry the running diagonals; both halves update theirs identically from the published u), the thread whose half holds
// the pivot row fetches u_k = A[q, k] / d and publishes it, and after one block barrier both halves apply the rank-1 update to
// their 64 rows.  Two block barriers per step; every element sees exactly the operations of pivoted_cholesky_rl_kernel, so for
// n <= 64 the two kernels are bitwise identical (tests/test_cholesky_emulation_cpu.py) and pivots / tie-breaking are LAPACK's.
// The team kernel needs 268 us at n = 128 (profiles/r1_cholesky_probe.txt) and, at 132 KB of shared memory, cannot share an SM
// with a GEMM CTA: this is the factorisation the rank-128 sweeps wait for.
__global__ void __launch_bounds__(CHT_THREADS) pivoted_cholesky_rl2_kernel(const double *__restrict__ Gin, int n, double tol,
                                                                           double *__restrict__ Wg, int *__restrict__ piv,
                                                                           int *__restrict__ status) {
    extern __shared__ double sm_dyn[];            // Gamma staging, then Uo[l + ldw * original column] = row l of the factor
    __shared__ __align__(16) double s_u[128];     // u of the current step by original column; 0 once a column is eliminated
    __shared__ unsigned long long s_key[4];
    __shared__ int s_idx[4], s_bad[4];
    __shared__ int s_orig[128];                   // position -> original column
    __shared__ int s_rank;
    const int ldw = n | 1;
    const int tid = threadIdx.x;
    const int k = tid & 127, h = tid >> 7, lane = tid & 31, w = (tid >> 5) & 3;
    double *W = sm_dyn;
    for (int e = tid; e < n * n; e += CHT_THREADS) W[(e % n) + (size_t)ldw * (e / n)] = Gin[e];
    if (tid < 128) { s_orig[tid] = tid; s_u[tid] = 0.0; }
    if (tid == 0) s_rank = n;
    __syncthreads();
    const bool mine = k < n;
    double col[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) col[i] = (mine && 64 * h + i < n) ? W[(64 * h + i) + (size_t)ldw * k] : 0.0;
    double ddk = mine ? W[k + (size_t)ldw * k] : 0.0;
    bool alive = mine;
    int posk = k;
    double stop = 0.0;
    int rank = n;
    __syncthreads();  // every column is in registers: the staging area becomes Uo
    for (int j = 0; j < n; ++j) {
        // ---- pivot: first maximum of the running diagonal in position order (h = 0 threads: warps 0 .. 3) ----
        if (h == 0) {
            const unsigned long long key = alive ? ordered_key(ddk) : 0ull;
            const unsigned hi = (unsigned)(key >> 32);
            const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
            const unsigned lo = (hi == mh) ? (unsigned)key : 0u;
            const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
            const bool win = alive && hi == mh && (unsigned)key == ml;
            const unsigned mi = __reduce_min_sync(0xffffffffu, win ? (((unsigned)posk << 8) | (unsigned)k) : 0x7fffffffu);
            const int bad = __any_sync(0xffffffffu, alive && ddk != ddk);
            if (lane == 0) { s_key[w] = ((unsigned long long)mh << 32) | ml; s_idx[w] = (int)mi; s_bad[w] = bad; }
        }
        __syncthreads();
        unsigned long long bk = s_key[0];
        int bi = s_idx[0], bb = s_bad[0];
#pragma unroll
        for (int q4 = 1; q4 < 4; ++q4) {
            const unsigned long long kq = s_key[q4];
            const int iq = s_idx[q4];
            if (kq > bk || (kq == bk && iq < bi)) { bk = kq; bi = iq; }
            bb |= s_bad[q4];
        }
        const double bv = key_value(bk);
        bool fail;
        if (j == 0) {
            stop = (tol < 0.0) ? n * DBL_EPSILON * bv : tol;
            fail = bb || !(bv > 0.0);
        } else {
            fail = bb || !(bv > stop);
        }
        if (fail) { rank = j; break; }  // uniform over the block
        const int q = bi & 0xff, p = bi >> 8;
        const double d = sqrt(bv);
        // ---- interchange bookkeeping only: the pivot goes to position j, the column that sat there to position p ----
        if (h == 0 && mine) {
            if (k == q) { posk = j; s_orig[j] = k; }
            else if (posk == j) { posk = p; s_orig[p] = k; }
        }
        // ---- row j of the factor: the half that holds row q of my column fetches it ----
        if (alive) {
            if (k == q) {
                if (h == 0) { W[j + (size_t)ldw * k] = d; s_u[k] = 0.0; }
                alive = false;
            } else if (h == (q >> 6)) {
                const double u = rl_fetch<64>(col, q & 63) / d;
                W[j + (size_t)ldw * k] = u;
                s_u[k] = u;
            }
        }
        __syncthreads();
        // ---- rank-1 update of my 64 rows (rows of eliminated columns see u = 0); the next step's first barrier orders these
        // reads of s_u before its next writes ----
        if (alive) {
            const double u = s_u[k];
            ddk = fma(-u, u, ddk);
#pragma unroll
            for (int i = 0; i < 64; i += 2) {
                const double2 uu = *reinterpret_cast<const double2 *>(&s_u[64 * h + i]);
                col[i] = fma(-uu.x, u, col[i]);
                col[i + 1] = fma(-uu.y, u, col[i + 1]);
            }
        }
    }
    if (tid == 0) s_rank = rank;
    __syncthreads();
    rank = s_rank;
    for (int e = tid; e < ldw * n; e += CHT_THREADS) {
        const int c = e / ldw, l = e - c * ldw;
        Wg[e] = (l <= c && l < rank) ? W[l + (size_t)ldw * s_orig[c]] : 0.0;
    }
    for (int e = tid; e < n; e += CHT_THREADS) piv[e] = s_orig[e];
    if (tid == 0) { status[0] = (rank == n) ? ITCPD_SOLVE_CHOLESKY : ITCPD_SOLVE_QRCP; status[1] = rank; status[2] = (rank == n) ? 0 : 1; }
}

// One warp per right-hand side: x = P (U^T U)^{-1} P^T b with b = row i of M, result to row i of X.
// Lane l keeps entries k = l + 32 e (e < E) in registers.  mode: 0 full solve, 1 forward only and
// return ||y[0:nn]||^2 (leverage scores).
constexpr int TSW_WARPS = 8;

template <int E>
__global__ void __launch_bounds__(TSW_WARPS * 32) chol_solve_warp_kernel(const double *__restrict__ Wg, const int *__restrict__ piv,
                                                                         const int *__restrict__ status, PeerSrc src,
                                                                         int64_t rows, int n, double *__restrict__ X, int u_in_smem,
                                                                         int fwd_only, int nn_dev_slot, QrcpWs qr) {
    extern __shared__ double sm_dyn[];
    peer_wait_all(src);
    const bool chol_ok = fwd_only || status[0] == ITCPD_SOLVE_CHOLESKY;
    const int ldw = n | 1;
    const double *U = Wg;
    __shared__ double s_rd[1024];  // reciprocal diagonal (what OpenBLAS' trsm kernels multiply by)
    __shared__ int s_pv[1024];
    double *s_rows = sm_dyn + (u_in_smem ? (((size_t)ldw * n + 1) & ~(size_t)1) : 0);  // TSW_WARPS x 1024 doubles, only when src.reduced_out
    if (chol_ok) {
        for (int e = threadIdx.x; e < n; e += TSW_WARPS * 32) { s_rd[e] = 1.0 / Wg[e + (size_t)ldw * e]; s_pv[e] = piv[e]; }
        if (u_in_smem) {
            // the workspace is 16-byte aligned and holds ldw * n + n doubles: stage it with 16-byte loads, a batch in flight per thread
            if (((reinterpret_cast<uintptr_t>(Wg) | reinterpret_cast<uintptr_t>(sm_dyn)) & 15) == 0) {
                const int n2 = (ldw * n + 1) >> 1;
                const double2 *g2 = reinterpret_cast<const double2 *>(Wg);
                double2 *s2 = reinterpret_cast<double2 *>(sm_dyn);
#pragma unroll 4
                for (int e = threadIdx.x; e < n2; e += TSW_WARPS * 32) s2[e] = g2[e];
            } else {
                for (int e = threadIdx.x; e < ldw * n; e += TSW_WARPS * 32) sm_dyn[e] = Wg[e];
            }
            U = sm_dyn;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t i = blockIdx.x * (int64_t)TSW_WARPS + (threadIdx.x >> 5);
    if (i >= rows) return;
    if (!chol_ok) {
        // rank deficient (the pivoted Cholesky met a pivot <= tol): reduce the peers' rows if any, then lane 0 runs the pivoted-QR
        // min-norm solve of this row (qrcp_rows.cuh; its factorisation ran right behind the Cholesky).  Rare path: one lane per row.
        const double *Mrow = src.p[0];
        if (src.reduced_out) {
            for (int k = lane; k < n; k += 32) {
                const int64_t off = i + rows * (int64_t)k;
                double v = 0.0;
                for (int q = 0; q < src.n; ++q) v += src.p[q][off];
                src.reduced_out[off] = v;
            }
            __syncwarp();
            Mrow = src.reduced_out;
        }
        if (lane == 0) qrcp_row_solve(qr.ws, qr.jpvt, status[1], Mrow, rows, n, n, X, qr.bglob, i);
        return;
    }
    const int nn = fwd_only ? status[nn_dev_slot] : n;
    double b[E];
    if (src.reduced_out) {
        // fused all-reduce: every peer's row is read ONCE (unpermuted, coalesced, fixed rank order: identical bits on
        // every rank), stored as the reduced matrix (later readers: the fit) and permuted through a
        // per-warp shared-memory row
        double *rowbuf = s_rows + (size_t)(threadIdx.x >> 5) * 1024;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = lane + 32 * e;
            if (k < n) {
                const int64_t off = i + rows * (int64_t)k;
                double v = 0.0;
                for (int q = 0; q < src.n; ++q) v += src.p[q][off];
                src.reduced_out[off] = v;
                rowbuf[k] = v;
            }
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = lane + 32 * e;
            b[e] = (k < n) ? rowbuf[s_pv[k]] : 0.0;
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = lane + 32 * e;
            b[e] = (k < n) ? src.p[0][i + rows * (int64_t)s_pv[k]] : 0.0;
        }
    }
    // forward: U^T y = P^T b  (column-oriented: after y_k is known, b_l -= U[k,l] y_k for l > k)
    for (int k = 0; k < nn; ++k) {
        const int owner = k & 31, eo = k >> 5;
        double yk = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (e == eo) yk = b[e];
        yk = __shfl_sync(0xffffffffu, yk, owner) * s_rd[k];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int l = lane + 32 * e;
            if (l == k) b[e] = yk;
            else if (l > k && l < nn) b[e] = fma(-U[k + (size_t)ldw * l], yk, b[e]);
        }
    }
    if (fwd_only == 2) {
        // y = U^{-T} P^T a_i is row i of Q_1 = A P U^{-1} (Cholesky-QR): store it, zero beyond the numerical rank
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = lane + 32 * e;
            if (k < n) X[i + rows * (int64_t)k] = (k < nn) ? b[e] : 0.0;
        }
        return;
    }
    if (fwd_only) {
        double s = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (lane + 32 * e < nn) s = fma(b[e], b[e], s);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) X[i] = s;
        return;
    }
    // backward: U x = y  (after x_k is known, y_l -= U[l,k] x_k for l < k)
    for (int k = n - 1; k >= 0; --k) {
        const int owner = k & 31, eo = k >> 5;
        double xk = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e)
            if (e == eo) xk = b[e];
        xk = __shfl_sync(0xffffffffu, xk, owner) * s_rd[k];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int l = lane + 32 * e;
            if (l == k) b[e] = xk;
            else if (l < k) b[e] = fma(-U[l + (size_t)ldw * k], xk, b[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int k = lane + 32 * e;
        if (k < n) X[i + rows * (int64_t)s_pv[k]] = b[e];
    }
}

// dynamic shared memory budget: 227 KB per CTA minus the static arrays of these kernels (pivots, reciprocal diagonal)
static int smem_limit(itcpd_ctx *) { return 208 * 1024; }

static int run_cholesky(itcpd_ctx *c, const double *Gamma, int R, double tol, int *status_dev) {
    const int ldw = R | 1;
    TRY(c->solve_ws.reserve(((size_t)ldw * R + R) * 8 + ((size_t)R * R + 8 * (size_t)R) * 8 + 1024));
    TRY(c->ipiv.reserve((size_t)R * 4 * 2));
    const size_t need = ((size_t)ldw * R + R) * 8;
    const int use_smem = need <= (size_t)smem_limit(c);
    static bool attr[64] = {false};  // function attributes are per device
    if (!attr[c->device & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(pivoted_cholesky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit(c)));
        attr[c->device & 63] = true;
    }
    const bool right_looking = c->chol_alg == 2 || (c->chol_alg == 3 && (c->chol_exposed || R > 64));
    if (right_looking && R > 64 && R <= 128) {  // two threads per column (experimental)
        static bool attr_r[64] = {false};
        if (!attr_r[c->device & 63]) {
            CUDA_TRY(cudaFuncSetAttribute(pivoted_cholesky_rl2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit(c)));
            attr_r[c->device & 63] = true;
        }
        pivoted_cholesky_rl2_kernel<<<1, CHT_THREADS, (size_t)ldw * R * 8, c->stream>>>(Gamma, R, tol, c->solve_ws.as<double>(), c->ipiv.as<int>(), status_dev);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        return ITCPD_OK;
    }
    if (right_looking && R <= 64) {
        const size_t smem = (size_t)ldw * R * 8;
        if (R <= 32) pivoted_cholesky_rl_kernel<32><<<1, CHT_THREADS, smem, c->stream>>>(Gamma, R, tol, c->solve_ws.as<double>(), c->ipiv.as<int>(), status_dev);
        else pivoted_cholesky_rl_kernel<64><<<1, CHT_THREADS, smem, c->stream>>>(Gamma, R, tol, c->solve_ws.as<double>(), c->ipiv.as<int>(), status_dev);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        return ITCPD_OK;
    }
    if (c->chol_alg >= 1 && R <= CHT_MAX_N) {
        static bool attr_t[64] = {false};
        if (!attr_t[c->device & 63]) {
            CUDA_TRY(cudaFuncSetAttribute(pivoted_cholesky_team_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit(c)));
            attr_t[c->device & 63] = true;
        }
        pivoted_cholesky_team_kernel<<<1, CHT_THREADS, (size_t)ldw * R * 8, c->stream>>>(Gamma, R, tol, c->solve_ws.as<double>(),
                                                                                          c->ipiv.as<int>(), status_dev);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        return ITCPD_OK;
    }
    const int threads = R <= 64 ? 64 : (R <= 128 ? 128 : CH_THREADS);
    pivoted_cholesky_kernel<<<1, threads, use_smem ? need : 0, c->stream>>>(Gamma, R, tol, c->solve_ws.as<double>(), c->ipiv.as<int>(),
                                                                                 status_dev, use_smem);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

template <int E>
static int launch_tsw(itcpd_ctx *c, const PeerSrc &M, int64_t rows, int R, double *X, const int *status_dev, int fwd_only, int slot) {
    const size_t u_bytes = (((size_t)(R | 1) * R + 1) & ~(size_t)1) * 8;   // rounded up to 16 bytes (staged with 16-byte loads)
    const int u_in = u_bytes + (size_t)TSW_WARPS * 1024 * 8 <= (size_t)smem_limit(c);
    auto kern = chol_solve_warp_kernel<E>;
    static bool attr[64] = {false};  // function attributes are per device
    if (!attr[c->device & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit(c)));
        attr[c->device & 63] = true;
    }
    const size_t row_bytes = M.reduced_out ? (size_t)TSW_WARPS * 1024 * 8 : 0;
    QrcpWs w;
    memset(&w, 0, sizeof(w));
    if (!fwd_only) TRY(qrcp_workspace(c, R, R, rows, &w));
    kern<<<(unsigned)ceil_div(rows, TSW_WARPS), TSW_WARPS * 32, (u_in ? u_bytes : 0) + row_bytes, c->stream>>>(
        c->solve_ws.as<double>(), c->ipiv.as<int>(), status_dev, M, rows, R, X, u_in, fwd_only, slot, w);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

static int run_tri_solves(itcpd_ctx *c, const PeerSrc &M, int64_t rows, int R, double *X, const int *status_dev, int fwd_only, int slot) {
    const int E = (int)ceil_div(R, 32);
    if (E <= 1) return launch_tsw<1>(c, M, rows, R, X, status_dev, fwd_only, slot);
    if (E <= 2) return launch_tsw<2>(c, M, rows, R, X, status_dev, fwd_only, slot);
    if (E <= 4) return launch_tsw<4>(c, M, rows, R, X, status_dev, fwd_only, slot);
    if (E <= 8) return launch_tsw<8>(c, M, rows, R, X, status_dev, fwd_only, slot);
    if (E <= 16) return launch_tsw<16>(c, M, rows, R, X, status_dev, fwd_only, slot);
    if (E <= 32) return launch_tsw<32>(c, M, rows, R, X, status_dev, fwd_only, slot);
    set_error("rank %d is above the 1024 limit of the solve kernels", R);
    return ITCPD_ERR_UNSUPPORTED;
}

static PeerSrc single_src(const double *M) {
    PeerSrc s;
    memset(&s, 0, sizeof(s));
    s.p[0] = M;
    s.n = 1;
    return s;
}

// The factorisation only needs Gamma (the Gram-Hadamard), not the MTTKRP, so the sweep driver runs it on a side stream
// underneath the GEMM pass; k_solve_apply then joins and applies it to the rows of M.  The factorisation half of the
// rank-deficient fallback (pivoted QR; a device-side early exit unless the Cholesky stopped at a pivot <= tol) rides along,
// so that the apply step is a single launch.
int k_solve_factor(itcpd_ctx *c, const double *Gamma, int R, double tol, int *status_dev) {
    TRY(run_cholesky(c, Gamma, R, tol, status_dev));
    return qrcp_factor_only(c, Gamma, R, R, 0, status_dev, 0);
}

// rank-deficient systems take the pivoted-QR min-norm path INSIDE the row-solve kernel (lane 0 of the row's warp runs
// qrcp_row_solve; the factorisation half ran behind the Cholesky), so a mode update has no no-op launches on its critical path
static int apply_rows(itcpd_ctx *c, const PeerSrc &src, int64_t rows, int R, double *X, int *status_dev) {
    return run_tri_solves(c, src, rows, R, X, status_dev, 0, 1);
}

int k_solve_apply(itcpd_ctx *c, const double *Gamma, const double *M, int64_t rows, int R, double *X, int *status_dev) {
    (void)Gamma;
    return apply_rows(c, single_src(M), rows, R, X, status_dev);
}

// fused all-reduce + solve: the right-hand sides are the sum of the peers' partial MTTKRPs (read over NVLink while
// loading); src.reduced_out (this rank's M buffer) receives the reduced matrix for the rank-deficient fallback
int k_solve_apply_peers(itcpd_ctx *c, const double *Gamma, const PeerSrc &src, int64_t rows, int R, double *X, int *status_dev) {
    (void)Gamma;
    return apply_rows(c, src, rows, R, X, status_dev);
}

int k_solve(itcpd_ctx *c, const double *Gamma, const double *M, int64_t rows, int R, double tol, double *X, int *status_dev) {
    TRY(k_solve_factor(c, Gamma, R, tol, status_dev));
    return k_solve_apply(c, Gamma, M, rows, R, X, status_dev);
}

// leverage scores (math_tools/probability.jl:3-10): p_i = ||Q[i,:]||^2 / min(I,R) with A = QR  (the reference takes a Householder QR).
// Cholesky-QR with a correction: U = chol(A^T A) (pivoted: a numerically rank-deficient factor is truncated at its rank), Q_1 = A P U^{-1}
// by forward substitution.  In floating point Q_1^T Q_1 = G_2 = I + E with |E| ~ cond(A)^2 eps, and EXACTLY
//     ||Q[i,:]||^2 = a_i^T (A^T A)^{-1} a_i = q_i^T G_2^{-1} q_i        (q_i = row i of Q_1),
// so the scores are the quadratic form with G_2^{-1} ~ I - E + E^2 = 3 I - 3 G_2 + G_2^2 (error |E|^3): the accuracy of a second
// Cholesky-QR pass without a second factorisation.  Plain ||q_i||^2 (round 1) was off by cond(A)^2 eps -- 1e-4 for the nearly collinear
// factors of late ALS sweeps; with the correction the error is ~ cond(A) eps + (cond(A)^2 eps)^3
// (tests/test_gpu_sampled.py::test_leverage_scores_ill_conditioned: cond 1e6, scores within 1e-9 of a Householder QR).
__global__ void fill_kernel(double *x, int64_t n, double v) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}

// W = 3 I - 3 G2 + G2 G2   (R x R; G2 symmetric)
__global__ void __launch_bounds__(256) neumann_kernel(const double *__restrict__ G2, int R, double *__restrict__ W) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= R * R) return;
    const int r1 = e % R, r2 = e / R;
    double s = 0.0;
    for (int k = 0; k < R; ++k) s = fma(G2[r1 + (size_t)R * k], G2[k + (size_t)R * r2], s);
    W[e] = s - 3.0 * G2[e] + (r1 == r2 ? 3.0 : 0.0);
}

// lev[i] = scale * q_i^T W q_i.  R <= 128: W staged in shared memory, one warp per row with the row broadcast from shared memory
// and lane l owning the outputs t_k = sum_j W[k,j] q_j for k = l, l + 32, ... (no shuffle inside the j loop: one reduction per row).
// Larger R: W read through the L1/L2, same arithmetic order.
constexpr int QF_WARPS = 8, QF_ROWS = 1;   // rows per warp: one -- a factor of 1024 rows then fills 128 CTAs instead of 32 (33 -> 10 us)
constexpr int QF_KPL = 4;                  // outputs per lane kept as independent accumulation chains (R <= 128)

__global__ void __launch_bounds__(QF_WARPS * 32) quadform_rows_kernel(const double *__restrict__ Q, const double *__restrict__ W, int64_t rows, int R,
                                                                       double scale, double *__restrict__ lev, int w_in_smem) {
    extern __shared__ double sm_q[];          // [W: R x R when staged] [q rows: QF_WARPS x R]
    const double *Wp = W;
    if (w_in_smem) {
        for (int e = threadIdx.x; e < R * R; e += QF_WARPS * 32) sm_q[e] = W[e];
        Wp = sm_q;
    }
    double *qrow = sm_q + (w_in_smem ? (size_t)R * R : 0) + (size_t)(threadIdx.x >> 5) * R;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int rr = 0; rr < QF_ROWS; ++rr) {
        const int64_t i = ((int64_t)blockIdx.x * QF_WARPS + (threadIdx.x >> 5)) * QF_ROWS + rr;
        if (i >= rows) break;   // uniform over the warp
        for (int k = lane; k < R; k += 32) qrow[k] = Q[i + rows * (int64_t)k];
        __syncwarp();
        double acc = 0.0;
        for (int k0 = lane; k0 < R; k0 += 32 * QF_KPL) {
            // t_k = sum_j W[k, j] q_j for up to QF_KPL values of k at once: the chains are independent, each sums j in ascending order
            double t[QF_KPL];
            int kk[QF_KPL];
#pragma unroll
            for (int u = 0; u < QF_KPL; ++u) { t[u] = 0.0; kk[u] = min(k0 + 32 * u, R - 1); }   // clamped rows are computed and dropped
            for (int j = 0; j < R; ++j) {
                const double qj = qrow[j];
#pragma unroll
                for (int u = 0; u < QF_KPL; ++u) t[u] = fma(Wp[kk[u] + (size_t)R * j], qj, t[u]);   // W is symmetric: lanes read consecutive words
            }
#pragma unroll
            for (int u = 0; u < QF_KPL; ++u)
                if (k0 + 32 * u < R) acc = fma(qrow[k0 + 32 * u], t[u], acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) lev[i] = scale * acc;
        __syncwarp();
    }
}

static int run_quadform(itcpd_ctx *c, const double *Q, const double *W, int64_t rows, int R, double scale, double *lev) {
    const unsigned grid = (unsigned)ceil_div(rows, QF_WARPS * QF_ROWS);
    const int w_in = R <= 128;
    const size_t smem = ((w_in ? (size_t)R * R : 0) + (size_t)QF_WARPS * R) * 8;
    static bool attr[64] = {false};
    if (!attr[c->device & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(quadform_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit(c)));
        attr[c->device & 63] = true;
    }
    ARG_CHECK(smem <= (size_t)smem_limit(c), "rank too large for the leverage kernels");
    quadform_rows_kernel<<<grid, QF_WARPS * 32, smem, c->stream>>>(Q, W, rows, R, scale, lev, w_in);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// `rows` local rows of a factor with rows_total rows overall (rows_total == rows unless the factor is slab-sharded; G is the
// GLOBAL Gram then, and G2 is all-reduced by the caller-supplied hook so that every rank applies the same correction)
static int leverage_impl(itcpd_ctx *c, const double *A, const double *G, int64_t rows, int64_t rows_total, int R, double *lev_out) {
    if (rows_total <= R) {  // square/wide factor: Q is orthogonal, every row has unit norm
        fill_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, c->stream>>>(lev_out, rows, 1.0 / (double)rows_total);
        c->launches++;
        return ITCPD_OK;
    }
    TRY(c->status.reserve(256));
    int *st = c->status.as<int>() + 32;
    TRY(run_cholesky(c, G, R, -1.0, st));
    TRY(c->lev_q.reserve(((size_t)rows * R + 2 * (size_t)R * R) * 8));
    double *Q1 = c->lev_q.as<double>(), *G2 = Q1 + (size_t)rows * R, *W = G2 + (size_t)R * R;
    TRY(run_tri_solves(c, single_src(A), rows, R, Q1, st, 2, 1));            // Q_1 = A P U^{-1}, zero beyond the numerical rank
    TRY(k_gram(c, Q1, rows, R, G2));
    if (comm_active(c) && rows != rows_total) TRY(comm_allreduce_sum(c, G2, (int64_t)R * R));
    neumann_kernel<<<(unsigned)ceil_div((int64_t)R * R, 256), 256, 0, c->stream>>>(G2, R, W);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return run_quadform(c, Q1, W, rows, R, 1.0 / (double)std::min<int64_t>(rows_total, R), lev_out);
}

int k_leverage(itcpd_ctx *c, const double *A, const double *G, int64_t rows, int R, double *lev_out) {
    return leverage_impl(c, A, G, rows, rows, R, lev_out);
}

// the same for `rows` local rows of a factor with rows_total rows overall (slab-sharded factor; G is the global Gram)
int k_leverage_rows(itcpd_ctx *c, const double *A, const double *G, int64_t rows_local, int64_t rows_total, int R, double *lev_out) {
    return leverage_impl(c, A, G, rows_local, rows_total, R, lev_out);
}

}  // namespace itcpd
