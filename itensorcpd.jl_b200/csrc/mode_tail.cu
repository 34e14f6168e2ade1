// Fused tail of one mode update inside the device-resident sweep (optimize.jl:25-30): row_norm (row_norm.jl:4-24), post_solve's Gram
// (standard/tensor.jl:46-49), for the last mode the FitCheck scalars (fit_check.jl:28-29, converge_checks.jl:5-11) and the sweep log --
// three launches instead of the nine of the hook-by-hook path, and ONE small all-reduce instead of two when the mode is slab-sharded.
//
//   xtx_partial_kernel   part[z] = X_z^T X_z over row slices z (X = the unnormalised solve output); for the last mode also the column
//                        dots d_r = sum_i M[i,r] X[i,r] = <T, That> per column (A lambda = X, so the fit needs no pass over A)
//   mode_finalize_kernel one CTA: sums the slices in fixed order, (sharded mode) all-reduces the R x R matrix over the peer-mapped
//                        exchange buffer, lambda_r = sqrt((X^T X)_rr), G = (X^T X) ./ (lambda lambda^T)  (= A^T A for A = X ./ lambda),
//                        fit scalars, sweep-log append
//   scale_cols_lambda_kernel   A = X ./ lambda  (true division, no zero guard: like the reference)
//
// G and lambda come from the Gram of X instead of a second pass over the normalised factor: identical in exact arithmetic, a few ulp
// apart in floating point; the fit trajectories stay within the 1e-9 bar (tests/test_gpu_fullsize.py, tests/test_gpu_config_a.py,
// bench.py `parity`).  The hook-by-hook entry points (itcpd_normalize / itcpd_post_solve / itcpd_fit_terms) keep the reference's
// literal operation order (kernels.cu); tests/test_gpu_dense.py::test_per_hook_path_equals_fused_sweeps ties the two together.
// Every reduction is fixed-order: results are bitwise reproducible and identical on replicated ranks.
#include "common.cuh"
#include "peer.cuh"

namespace itcpd {

constexpr int XT = 16;       // output tile
constexpr int XCHUNK = 64;   // rows per shared-memory chunk
constexpr int XSLICE = 128;  // rows per CTA slice

__global__ void __launch_bounds__(XT *XT) xtx_partial_kernel(const double *__restrict__ X, const double *__restrict__ M, int64_t rows, int R,
                                                             double *__restrict__ part, double *__restrict__ dotpart) {
    __shared__ double sa[XT][XCHUNK + 1], sb[XT][XCHUNK + 1], sm[XT][XCHUNK + 1];
    const int tx = threadIdx.x % XT, ty = threadIdx.x / XT;
    const int r1 = blockIdx.x * XT + tx, r2 = blockIdx.y * XT + ty;
    const bool diag = blockIdx.x == blockIdx.y;
    const bool dots = diag && M != nullptr;
    const int64_t i0 = (int64_t)blockIdx.z * XSLICE;
    const int64_t i1 = min(rows, i0 + XSLICE);
    double acc = 0.0, dacc = 0.0;
    for (int64_t base = i0; base < i1; base += XCHUNK) {
        for (int q = threadIdx.x; q < XT * XCHUNK; q += XT * XT) {
            const int col = q / XCHUNK, ii = q % XCHUNK;   // consecutive threads read consecutive rows of one column
            const int64_t i = base + ii;
            const int ca = blockIdx.x * XT + col, cb = blockIdx.y * XT + col;
            const double va = (i < i1 && ca < R) ? X[i + rows * (int64_t)ca] : 0.0;
            sa[col][ii] = va;
            sb[col][ii] = diag ? va : ((i < i1 && cb < R) ? X[i + rows * (int64_t)cb] : 0.0);
            if (dots) sm[col][ii] = (i < i1 && ca < R) ? M[i + rows * (int64_t)ca] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int ii = 0; ii < XCHUNK; ++ii) acc = fma(sa[tx][ii], sb[ty][ii], acc);
        if (dots && ty == 0) {
#pragma unroll 8
            for (int ii = 0; ii < XCHUNK; ++ii) dacc = fma(sm[tx][ii], sa[tx][ii], dacc);
        }
        __syncthreads();
    }
    if (r1 < R && r2 < R) part[(size_t)blockIdx.z * R * R + r1 + (size_t)R * r2] = acc;
    if (dots && ty == 0 && r1 < R) dotpart[(size_t)blockIdx.z * R + r1] = dacc;
}

struct TailGrams { const double *g[ITCPD_MAX_ORDER]; int n; int mode; };

template <int THREADS>
__device__ __forceinline__ double tail_block_sum(double v, double *sh) {
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

// One CTA.  xtx: R x R scratch in global memory (sums of the slices; the all-reduced matrix when sharded).
__global__ void __launch_bounds__(256) mode_finalize_kernel(const double *__restrict__ part, const double *__restrict__ dotpart, int slices, int R,
                                                            double *xtx, double *G, double *lambda,
                                                            PeerPtrs peers, long long *epoch_dev, size_t small_off, int64_t slot_doubles,
                                                            TailGrams grams, int with_fit, double *__restrict__ fit2,
                                                            const int *__restrict__ status, int nmodes, double *__restrict__ log,
                                                            unsigned long long *__restrict__ counter, unsigned long long cap) {
    __shared__ double sh[8];
    __shared__ long long s_e;
    const int n2 = R * R;
    // 1. slices -> X^T X, fixed order
    for (int e = threadIdx.x; e < n2; e += 256) {
        double s = 0.0;
        for (int z = 0; z < slices; ++z) s += part[(size_t)z * n2 + e];
        xtx[e] = s;
    }
    // 2. slab-sharded mode: all-reduce over the peer-mapped exchange buffers (write mine, publish, wait, sum in rank order:
    //    bitwise identical on every rank); double buffered by the parity of the device-side epoch (peer_graph.cu)
    if (epoch_dev) {
        if (threadIdx.x == 0) { s_e = *epoch_dev + 1; *epoch_dev = s_e; }
        __syncthreads();
        const long long ep = s_e;
        const size_t off = small_off + (size_t)(ep & 1) * (size_t)slot_doubles * 8;
        double *mine = reinterpret_cast<double *>(peers.base[peers.rank] + off);
        for (int e = threadIdx.x; e < n2; e += 256) mine[e] = xtx[e];
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < peers.n) {
            volatile long long *d = reinterpret_cast<long long *>(peers.base[threadIdx.x]) + 16 + peers.rank;
            *d = ep;
            bounded_wait(reinterpret_cast<const volatile long long *>(peers.base[peers.rank]) + 16 + threadIdx.x, ep, (int)threadIdx.x);
        }
        __syncthreads();
        __threadfence_system();
        for (int e = threadIdx.x; e < n2; e += 256) {
            double v = 0.0;
            for (int q = 0; q < peers.n; ++q) v += reinterpret_cast<const volatile double *>(peers.base[q] + off)[e];
            xtx[e] = v;
        }
    }
    __syncthreads();
    // 3. lambda and the Gram of the normalised factor
    for (int r = threadIdx.x; r < R; r += 256) lambda[r] = sqrt(xtx[r + (size_t)R * r]);
    __syncthreads();
    for (int e = threadIdx.x; e < n2; e += 256) G[e] = xtx[e] / (lambda[e % R] * lambda[e / R]);
    if (!with_fit) return;
    __syncthreads();
    // 4. fit scalars: <T, That> = sum_r sum_i M[i,r] X[i,r] (slab-partial when sharded: reduced when the log is fetched),
    //    ||That||^2 = lambda^T (hadamard_n G_n) lambda
    double a = 0.0;
    for (int r = threadIdx.x; r < R; r += 256) {
        double s = 0.0;
        for (int z = 0; z < slices; ++z) s += dotpart[(size_t)z * R + r];
        a += s;
    }
    a = tail_block_sum<256>(a, sh);
    double q = 0.0;
    for (int e = threadIdx.x; e < n2; e += 256) {
        double h = (grams.mode == 0) ? G[e] : grams.g[0][e];
        for (int m = 1; m < grams.n; ++m) h = h * ((m == grams.mode) ? G[e] : grams.g[m][e]);
        q = fma(h, lambda[e % R] * lambda[e / R], q);
    }
    q = tail_block_sum<256>(q, sh);
    if (threadIdx.x == 0) {
        fit2[0] = a;
        fit2[1] = q;
        const unsigned long long idx = *counter;   // sweep log (api.cu: log_sweep_kernel)
        if (idx < cap) {
            int fb = 0;
            for (int m = 0; m < nmodes; ++m) fb += (status[3 * m] == ITCPD_SOLVE_QRCP);
            log[idx] = a;
            log[cap + idx] = q;
            log[2 * cap + idx] = (double)fb;
        }
        *counter = idx + 1;
    }
}

__global__ void scale_cols_lambda_kernel(const double *__restrict__ X, int64_t rows, const double *__restrict__ lambda, double *__restrict__ A) {
    const int r = blockIdx.y;
    const double l = lambda[r];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x)
        A[i + rows * (int64_t)r] = X[i + rows * (int64_t)r] / l;
}

bool mode_tail_supported(const itcpd_ctx *c) { return c->fused_tail && c->rank <= 128; }

// X (rows x R, the solve output) -> A[mode], lambda, G[mode]; with_fit: also fit2 + the sweep-log entry (last mode of a sweep)
int k_mode_tail(itcpd_ctx *c, int mode, bool with_fit, const int *status_dev, int nmodes) {
    const int R = c->rank;
    const int64_t rows = c->dims[mode];
    const double *X = c->X.as<double>();
    const bool sharded = comm_active(c) && mode == c->order - 1;
    const int slices = (int)std::max<int64_t>(1, ceil_div(rows, XSLICE));
    const size_t n2 = (size_t)R * R;
    TRY(c->redux.reserve(((size_t)slices * (n2 + R) + n2 + R) * 8));
    double *part = c->redux.as<double>(), *dotpart = part + (size_t)slices * n2, *xtx = dotpart + (size_t)slices * R, *dotsum = xtx + n2;
    dim3 grid((unsigned)ceil_div(R, XT), (unsigned)ceil_div(R, XT), (unsigned)slices);
    xtx_partial_kernel<<<grid, XT * XT, 0, c->stream>>>(X, with_fit ? c->M[mode].as<double>() : nullptr, rows, R, part, dotpart);
    c->launches++;
    PeerPtrs peers;
    memset(&peers, 0, sizeof(peers));
    long long *epoch_dev = nullptr;
    const double *fin_part = part, *fin_dots = dotpart;
    int fin_slices = slices;
    if (sharded) {
        if (peer_graph_active(c)) {
            ARG_CHECK((int64_t)n2 <= c->peer_small_doubles, "small all-reduce larger than the exchange buffer's small slots");
            peers = peer_ptrs(c);
            epoch_dev = c->peer_epochs.as<long long>() + 1;
        } else {
            // NCCL path: sum the slices, all-reduce the R x R matrix, then finalize from the single reduced "slice".  The column dots
            // stay slab-partial by design, so they are summed over the slices here and handed over as one slice as well.
            TRY(k_sum_slices(c, part, (int64_t)n2, slices, xtx));
            if (with_fit) TRY(k_sum_slices(c, dotpart, R, slices, dotsum));
            TRY(comm_allreduce_sum(c, xtx, (int64_t)n2));
            fin_part = xtx;
            fin_dots = dotsum;
            fin_slices = 1;
        }
    }
    TailGrams tg;
    memset(&tg, 0, sizeof(tg));
    tg.n = c->order;
    tg.mode = mode;
    for (int m = 0; m < c->order; ++m) tg.g[m] = c->G[m].as<double>();
    mode_finalize_kernel<<<1, 256, 0, c->stream>>>(fin_part, fin_dots, fin_slices, R, xtx, c->G[mode].as<double>(), c->lambda.as<double>(), peers, epoch_dev,
                                                   c->peer_small_off, c->peer_small_doubles, tg, with_fit ? 1 : 0, c->fit2.as<double>(), status_dev, nmodes,
                                                   c->sweep_log.as<double>() + 1, reinterpret_cast<unsigned long long *>(c->sweep_log.p),
                                                   (unsigned long long)c->sweep_log_cap);
    c->launches++;
    dim3 sgrid((unsigned)std::min<int64_t>(ceil_div(rows, 256), 64), (unsigned)R);
    scale_cols_lambda_kernel<<<sgrid, 256, 0, c->stream>>>(X, rows, c->lambda.as<double>(), c->A[mode].as<double>());
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

}  // namespace itcpd
