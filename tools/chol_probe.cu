// Standalone micro-benchmark (NOT part of the library): where does the R x R pivoted Cholesky spend its time on B200?
//   1. dependent-chain latencies of the FP64 / warp-collective / barrier instructions the kernel is made of
//   2. per-phase clock64 breakdown of the shipped "team" kernel (copy of solve.cu: pivoted_cholesky_team_kernel)
//   3. candidate: one warp, E columns per lane (no block barriers), with and without a reciprocal instead of the division
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/chol_probe tools/chol_probe.cu
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

// ---------------------------------------------------------------------------------------------------------------
// 1. latencies
// ---------------------------------------------------------------------------------------------------------------
template <int OP>
__global__ void lat_kernel(double *out, long long *cyc, double seed, int iters) {
    __shared__ int chase[64];
    double x = seed + threadIdx.x * 1e-3, y = 1.0000001;
    unsigned v = threadIdx.x;
    if (threadIdx.x < 64) chase[threadIdx.x] = (threadIdx.x + 7) & 63;
    __syncthreads();
    int idx = threadIdx.x & 63;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) x = fma(x, y, 1e-9);
        if (OP == 1) x = x + y;
        if (OP == 2) x = x * y;
        if (OP == 3) x = sqrt(x) + 1.5;
        if (OP == 4) x = 2.0 / x + 0.5;
        if (OP == 5) v = __reduce_max_sync(0xffffffffu, v ^ (threadIdx.x & 31)) + 1u;
        if (OP == 6) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1u;
        if (OP == 7) idx = chase[idx];
        if (OP == 8) asm volatile("bar.sync 1, 64;" ::: "memory");
        if (OP == 9) __syncthreads();
        if (OP == 10) x = (x > y) ? x - 0.25 : x + 0.5;  // DSETP + select + DADD chain
        if (OP == 11) __syncwarp();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x + v + idx;
}

// ---------------------------------------------------------------------------------------------------------------
// shared helpers (same as solve.cu)
// ---------------------------------------------------------------------------------------------------------------
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

// ---------------------------------------------------------------------------------------------------------------
// 2. the shipped team kernel with optional phase clocks (thread n-1 is active in every phase of every column)
// ---------------------------------------------------------------------------------------------------------------
template <bool PROF>
__global__ void __launch_bounds__(256) team_kernel(const double *__restrict__ Gin, int n, double tol, double *__restrict__ Wg,
                                                   int *__restrict__ piv, int *__restrict__ status, long long *__restrict__ prof) {
    extern __shared__ double sm_dyn[];
    __shared__ unsigned long long s_key[4];
    __shared__ int s_idx[4], s_bad[4];
    __shared__ double s_swapdd;
    __shared__ int s_piv[128];
    __shared__ int s_rank;
    const int ldw = n | 1;
    const int tid = threadIdx.x;
    const int team = (n + 31) & ~31, nw = team >> 5;
    double *W = sm_dyn;
    long long pa = 0, ps = 0, pd = 0, pv = 0, pl = 0, tl0 = 0;
    if (PROF) tl0 = clock64();
    for (int e = tid; e < n * n; e += 256) W[(e % n) + (size_t)ldw * (e / n)] = Gin[e];
    if (tid < n) s_piv[tid] = tid;
    if (tid == 0) s_rank = n;
    __syncthreads();
    if (PROF) pl = clock64() - tl0;
    if (tid < team) {
        const int k = tid, lane = tid & 31, w = tid >> 5;
        const bool mine = k < n;
        double *ck = W + (size_t)ldw * (mine ? k : 0);
        double ddk = mine ? ck[k] : 0.0;
        double stop = 0.0;
        int rank = n;
        for (int j = 0; j < n; ++j) {
            long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
            if (PROF) c0 = clock64();
            const bool elig = mine && k >= j;
            const unsigned long long key = elig ? ordered_key(ddk) : 0ull;
            const unsigned hi = (unsigned)(key >> 32);
            const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
            const unsigned lo = (hi == mh) ? (unsigned)key : 0u;
            const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
            const bool win = elig && hi == mh && (unsigned)key == ml;
            const unsigned mi = __reduce_min_sync(0xffffffffu, win ? (unsigned)k : 0x7fffffffu);
            int bb = __any_sync(0xffffffffu, elig && ddk != ddk);
            unsigned long long bk = ((unsigned long long)mh << 32) | ml;
            int bi = (int)mi;
            if (nw > 1) {
                if (lane == 0) { s_key[w] = bk; s_idx[w] = bi; s_bad[w] = bb; }
                team_sync(nw, team);
                bk = s_key[0]; bi = s_idx[0]; bb = s_bad[0];
                for (int q = 1; q < nw; ++q) {
                    const unsigned long long kq = s_key[q];
                    if (kq > bk) { bk = kq; bi = s_idx[q]; }
                    bb |= s_bad[q];
                }
            }
            const double bv = key_value(bk);
            bool fail;
            if (j == 0) { stop = (tol < 0.0) ? n * DBL_EPSILON * bv : tol; fail = bb || !(bv > 0.0); }
            else fail = bb || !(bv > stop);
            if (fail) { rank = j; break; }
            const int p = bi;
            if (PROF) c1 = clock64();
            const double d = sqrt(bv);
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
                    s_swapdd = ddk;
                }
            }
            team_sync(nw, team);
            if (p != j && k == p) ddk = s_swapdd;
            if (PROF) c2 = clock64();
            if (mine && k > j) {
                const double *cj = W + (size_t)ldw * j;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int l = 0;
                for (; l + 3 < j; l += 4) {
                    s0 = fma(cj[l], ck[l], s0); s1 = fma(cj[l + 1], ck[l + 1], s1);
                    s2 = fma(cj[l + 2], ck[l + 2], s2); s3 = fma(cj[l + 3], ck[l + 3], s3);
                }
                for (; l < j; ++l) s0 = fma(cj[l], ck[l], s0);
                const double num = ck[j] - ((s0 + s1) + (s2 + s3));
                if (PROF) { c3 = clock64(); if (num == 1.2345e-300) c3++; }
                const double u = num / d;
                ck[j] = u;
                ddk = fma(-u, u, ddk);
            } else if (k == j) {
                ck[j] = d;
                if (PROF) c3 = clock64();
            }
            if (nw == 1) __syncwarp();
            if (PROF) { if (ddk == 1.2345e-300) c3++; c4 = clock64(); pa += c1 - c0; ps += c2 - c1; pd += c3 - c2; pv += c4 - c3; }
        }
        if (tid == 0) s_rank = rank;
        if (PROF && k == n - 1) { prof[0] = pa; prof[1] = ps; prof[2] = pd; prof[3] = pv; prof[4] = pl; }
    }
    __syncthreads();
    const int rank = s_rank;
    for (int e = tid; e < ldw * n; e += 256) Wg[e] = W[e];
    for (int e = tid; e < n; e += 256) piv[e] = s_piv[e];
    if (tid == 0) { status[0] = (rank == n) ? 0 : 1; status[1] = rank; status[2] = (rank == n) ? 0 : 1; }
    if (PROF && tid == 0) prof[5] = clock64() - tl0;
}

// ---------------------------------------------------------------------------------------------------------------
// 3. candidate: ONE warp, E columns per lane (k = lane + 32 e); same arithmetic as the team kernel when RCP == false
// ---------------------------------------------------------------------------------------------------------------
template <int E, bool RCP>
__global__ void __launch_bounds__(256) warp_kernel(const double *__restrict__ Gin, int n, double tol, double *__restrict__ Wg,
                                                   int *__restrict__ piv, int *__restrict__ status) {
    extern __shared__ double sm_dyn[];
    __shared__ int s_piv[32 * E];
    __shared__ int s_rank;
    const int ldw = n | 1;
    const int tid = threadIdx.x;
    double *W = sm_dyn;
    for (int e = tid; e < n * n; e += 256) W[(e % n) + (size_t)ldw * (e / n)] = Gin[e];
    if (tid < n) s_piv[tid] = tid;
    if (tid == 0) s_rank = n;
    __syncthreads();
    if (tid < 32) {
        const int lane = tid;
        double dd[E];
        double *ck[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = lane + 32 * e;
            ck[e] = W + (size_t)ldw * (k < n ? k : 0);
            dd[e] = (k < n) ? ck[e][k] : 0.0;
        }
        double stop = 0.0;
        int rank = n;
        for (int j = 0; j < n; ++j) {
            unsigned long long key = 0ull;
            unsigned kb = 0x7fffffffu;
            bool bad = false;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int k = lane + 32 * e;
                if (k < n && k >= j) {
                    const unsigned long long kk = ordered_key(dd[e]);
                    bad |= dd[e] != dd[e];
                    if (kk > key) { key = kk; kb = (unsigned)k; }
                }
            }
            const unsigned hi = (unsigned)(key >> 32);
            const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
            const unsigned lo = (hi == mh) ? (unsigned)key : 0u;
            const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
            const bool win = kb != 0x7fffffffu && hi == mh && (unsigned)key == ml;
            const unsigned mi = __reduce_min_sync(0xffffffffu, win ? kb : 0x7fffffffu);
            const int bb = __any_sync(0xffffffffu, bad);
            const double bv = key_value(((unsigned long long)mh << 32) | ml);
            bool fail;
            if (j == 0) { stop = (tol < 0.0) ? n * DBL_EPSILON * bv : tol; fail = bb || !(bv > 0.0); }
            else fail = bb || !(bv > stop);
            if (fail) { rank = j; break; }
            const int p = (int)mi;
            const double d = sqrt(bv);
            const double rd = RCP ? 1.0 / d : 0.0;
            if (p != j) {
                // running diagonal of column j moves to column p
                double ddj = 0.0;
#pragma unroll
                for (int e = 0; e < E; ++e) if (e == (j >> 5)) ddj = dd[e];
                ddj = __shfl_sync(0xffffffffu, ddj, j & 31);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int k = lane + 32 * e;
                    if (k < n) {
                        if (k != j && k != p) {
                            double a = ck[e][j], b = ck[e][p];
                            ck[e][j] = b; ck[e][p] = a;
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
                        } else {
                            dd[e] = ddj;  // k == p
                        }
                    }
                }
            }
            __syncwarp();
            const double *cj = W + (size_t)ldw * j;
            double s[E][4];
#pragma unroll
            for (int e = 0; e < E; ++e) s[e][0] = s[e][1] = s[e][2] = s[e][3] = 0.0;
            int l = 0;
            for (; l + 3 < j; l += 4) {
                const double a0 = cj[l], a1 = cj[l + 1], a2 = cj[l + 2], a3 = cj[l + 3];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    s[e][0] = fma(a0, ck[e][l], s[e][0]); s[e][1] = fma(a1, ck[e][l + 1], s[e][1]);
                    s[e][2] = fma(a2, ck[e][l + 2], s[e][2]); s[e][3] = fma(a3, ck[e][l + 3], s[e][3]);
                }
            }
            for (; l < j; ++l) {
                const double a0 = cj[l];
#pragma unroll
                for (int e = 0; e < E; ++e) s[e][0] = fma(a0, ck[e][l], s[e][0]);
            }
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int k = lane + 32 * e;
                if (k < n && k > j) {
                    const double num = ck[e][j] - ((s[e][0] + s[e][1]) + (s[e][2] + s[e][3]));
                    const double u = RCP ? num * rd : num / d;
                    ck[e][j] = u;
                    dd[e] = fma(-u, u, dd[e]);
                } else if (k == j) {
                    ck[e][j] = d;
                }
            }
            __syncwarp();
        }
        if (tid == 0) s_rank = rank;
    }
    __syncthreads();
    const int rank = s_rank;
    for (int e = tid; e < ldw * n; e += 256) Wg[e] = W[e];
    for (int e = tid; e < n; e += 256) piv[e] = s_piv[e];
    if (tid == 0) { status[0] = (rank == n) ? 0 : 1; status[1] = rank; status[2] = (rank == n) ? 0 : 1; }
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
static std::vector<double> gamma_matrix(int n, unsigned seed) {
    // Hadamard product of two Gram matrices of column-normalised random 512 x n factors (what the ALS solve sees)
    std::vector<double> G((size_t)n * n, 1.0);
    srand(seed);
    for (int f = 0; f < 2; ++f) {
        const int rows = 512;
        std::vector<double> X((size_t)rows * n);
        for (int c = 0; c < n; ++c) {
            double ss = 0.0;
            for (int r = 0; r < rows; ++r) { const double v = (rand() / (double)RAND_MAX) - 0.5; X[r + (size_t)rows * c] = v; ss += v * v; }
            const double inv = 1.0 / std::sqrt(ss);
            for (int r = 0; r < rows; ++r) X[r + (size_t)rows * c] *= inv;
        }
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < n; ++b) {
                double d = 0.0;
                for (int r = 0; r < rows; ++r) d += X[r + (size_t)rows * a] * X[r + (size_t)rows * b];
                G[a + (size_t)n * b] *= d;
            }
    }
    return G;
}

static double residual(const std::vector<double> &G, int n, const std::vector<double> &W, const std::vector<int> &piv) {
    const int ldw = n | 1;
    double worst = 0.0, amax = 0.0;
    for (int a = 0; a < n; ++a)
        for (int b = a; b < n; ++b) {
            double s = 0.0;
            for (int l = 0; l <= a; ++l) s += W[l + (size_t)ldw * a] * W[l + (size_t)ldw * b];
            const double ref = G[piv[a] + (size_t)n * piv[b]];
            worst = std::fmax(worst, std::fabs(s - ref));
            amax = std::fmax(amax, std::fabs(ref));
        }
    return worst / amax;
}

struct Dev {
    double *G = nullptr, *W = nullptr;
    int *piv = nullptr, *status = nullptr;
    long long *prof = nullptr;
};

template <typename L>
static int time_it(const char *name, int n, L launch, Dev &dv, const std::vector<double> &G, const std::vector<double> *refW, const std::vector<int> *refP) {
    const int ldw = n | 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 5; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 200;
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<double> W((size_t)ldw * n);
    std::vector<int> piv(n), st(3);
    CK(cudaMemcpy(W.data(), dv.W, W.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(piv.data(), dv.piv, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(st.data(), dv.status, 12, cudaMemcpyDeviceToHost));
    int same = -1;
    if (refW) same = (memcmp(W.data(), refW->data(), W.size() * 8) == 0 && memcmp(piv.data(), refP->data(), n * 4) == 0) ? 1 : 0;
    printf("  %-26s n=%3d  %8.2f us/launch  rank=%d  resid=%.2e  bitwise_vs_team=%d\n", name, n, 1e3 * ms / reps, st[1], residual(G, n, W, piv), same);
    fflush(stdout);
    return 0;
}

int main() {
    CK(cudaSetDevice(0));
    Dev dv;
    CK(cudaMalloc(&dv.G, 128 * 128 * 8));
    CK(cudaMalloc(&dv.W, 129 * 128 * 8));
    CK(cudaMalloc(&dv.piv, 128 * 4));
    CK(cudaMalloc(&dv.status, 64));
    CK(cudaMalloc(&dv.prof, 64));
    double *out;
    long long *cyc;
    CK(cudaMalloc(&out, 256 * 8));
    CK(cudaMalloc(&cyc, 8));

    printf("== dependent-chain latencies (cycles per op, one warp unless noted) ==\n");
    const int iters = 4096;
    const char *names[12] = {"DFMA", "DADD", "DMUL", "sqrt(x)+1.5", "2/x+0.5", "REDUX.MAX+IADD", "SHFL+IADD", "LDS chase",
                             "bar.sync 1,64 (2 warps)", "__syncthreads (8 warps)", "DSETP+sel+DADD", "__syncwarp"};
    for (int op = 0; op < 12; ++op) {
        const int threads = (op == 8) ? 64 : (op == 9 ? 256 : 32);
        switch (op) {
            case 0: lat_kernel<0><<<1, threads>>>(out, cyc, 1.0, iters); break;
            case 1: lat_kernel<1><<<1, threads>>>(out, cyc, 1.0, iters); break;
            case 2: lat_kernel<2><<<1, threads>>>(out, cyc, 1.0, iters); break;
            case 3: lat_kernel<3><<<1, threads>>>(out, cyc, 1.0, iters); break;
            case 4: lat_kernel<4><<<1, threads>>>(out, cyc, 1.0, iters); break;
            case 5: lat_kernel<5><<<1, threads>>>(out, cyc, 1.0, iters); break;
            case 6: lat_kernel<6><<<1, threads>>>(out, cyc, 1.0, iters); break;
            case 7: lat_kernel<7><<<1, threads>>>(out, cyc, 1.0, iters); break;
            case 8: lat_kernel<8><<<1, threads>>>(out, cyc, 1.0, iters); break;
            case 9: lat_kernel<9><<<1, threads>>>(out, cyc, 1.0, iters); break;
            case 10: lat_kernel<10><<<1, threads>>>(out, cyc, 1.0, iters); break;
            case 11: lat_kernel<11><<<1, threads>>>(out, cyc, 1.0, iters); break;
        }
        CK(cudaDeviceSynchronize());
        long long c = 0;
        CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        printf("  %-26s %8.1f\n", names[op], (double)c / iters);
    }
    fflush(stdout);

    CK(cudaFuncSetAttribute(team_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CK(cudaFuncSetAttribute(team_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CK(cudaFuncSetAttribute(warp_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CK(cudaFuncSetAttribute(warp_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CK(cudaFuncSetAttribute(warp_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CK(cudaFuncSetAttribute(warp_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CK(cudaFuncSetAttribute(warp_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));

    const int sizes[4] = {32, 50, 64, 128};
    for (int si = 0; si < 4; ++si) {
        const int n = sizes[si], ldw = n | 1;
        const size_t smem = (size_t)ldw * n * 8;
        std::vector<double> G = gamma_matrix(n, 1234 + n);
        CK(cudaMemcpy(dv.G, G.data(), (size_t)n * n * 8, cudaMemcpyHostToDevice));
        printf("== n = %d ==\n", n);
        // phase breakdown
        team_kernel<true><<<1, 256, smem>>>(dv.G, n, 1e-6, dv.W, dv.piv, dv.status, dv.prof);
        CK(cudaDeviceSynchronize());
        long long pr[6];
        CK(cudaMemcpy(pr, dv.prof, 48, cudaMemcpyDeviceToHost));
        printf("  team kernel phases (cycles, thread n-1): pivot %lld  swap+barrier %lld  dot %lld  div+store %lld | load %lld  total %lld  (per column: %.0f)\n",
               pr[0], pr[1], pr[2], pr[3], pr[4], pr[5], (double)(pr[0] + pr[1] + pr[2] + pr[3]) / n);
        fflush(stdout);
        if (time_it("team (shipped)", n, [&] { team_kernel<false><<<1, 256, smem>>>(dv.G, n, 1e-6, dv.W, dv.piv, dv.status, nullptr); }, dv, G, nullptr, nullptr)) return 1;
        std::vector<double> refW((size_t)ldw * n);
        std::vector<int> refP(n);
        CK(cudaMemcpy(refW.data(), dv.W, refW.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(refP.data(), dv.piv, n * 4, cudaMemcpyDeviceToHost));
        if (n <= 32) {
            if (time_it("one warp E=1", n, [&] { warp_kernel<1, false><<<1, 256, smem>>>(dv.G, n, 1e-6, dv.W, dv.piv, dv.status); }, dv, G, &refW, &refP)) return 1;
        }
        if (n <= 64) {
            if (time_it("one warp E=2", n, [&] { warp_kernel<2, false><<<1, 256, smem>>>(dv.G, n, 1e-6, dv.W, dv.piv, dv.status); }, dv, G, &refW, &refP)) return 1;
            if (time_it("one warp E=2 reciprocal", n, [&] { warp_kernel<2, true><<<1, 256, smem>>>(dv.G, n, 1e-6, dv.W, dv.piv, dv.status); }, dv, G, &refW, &refP)) return 1;
        }
        if (time_it("one warp E=4", n, [&] { warp_kernel<4, false><<<1, 256, smem>>>(dv.G, n, 1e-6, dv.W, dv.piv, dv.status); }, dv, G, &refW, &refP)) return 1;
        if (time_it("one warp E=4 reciprocal", n, [&] { warp_kernel<4, true><<<1, 256, smem>>>(dv.G, n, 1e-6, dv.W, dv.piv, dv.status); }, dv, G, &refW, &refP)) return 1;
    }
    printf("done\n");
    return 0;
}
