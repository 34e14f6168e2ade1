// Slab-sharded variant of the sampled least-squares update (SURVEY 8e "Sampled path"): every rank holds the slab
// T[..., off : off + slab] of the last mode and the matching rows of the last factor; all other state is replicated.
//   mode != last : a sample is OWNED by the rank whose slab contains its last-mode coordinate.  Each rank builds the
//                  owned rows of the sampled Khatri-Rao product and the owned fibres (zeros elsewhere); the sampled
//                  normal equations K'K (R x R) and T_s K (I_n x R) are sums over ranks -> two all-reduces.
//                  normal=false needs the full K and T_s on every rank: the zero-padded pieces are all-reduced.
//   mode == last : K comes from replicated factors (no exchange); the fibres run along the sharded mode, so each rank
//                  solves for ITS rows of the factor; column norms and the Gram are all-reduced as in the dense path.
// Pivots are GLOBAL 1-based coordinates and identical on every rank (same seed, replicated leverage scores; the
// sharded factor's scores are all-gathered before sampling).
// STATUS: compiled, algebra covered by the world_size-2 gloo test against the oracle; not yet run on a multi-GPU box.
#include "common.cuh"

namespace itcpd {

struct ShDims {
    int n;
    int64_t ext[ITCPD_MAX_ORDER], dim[ITCPD_MAX_ORDER];
};
struct ShFac { const double *a[ITCPD_MAX_ORDER]; };

static ShDims shdims(const itcpd_ctx *c) {
    ShDims d;
    d.n = c->order;
    for (int i = 0; i < c->order; ++i) { d.ext[i] = (i == 0) ? c->ld0 : c->dims[i]; d.dim[i] = c->dims[i]; }
    return d;
}

// K[s, r] = owned(s) ? prod_{m != mode} A_m[row_m(s), r] : 0   (mode != last; the last factor is indexed by its LOCAL row)
__global__ void pivot_hadamard_owned_kernel(ShFac fp, ShDims d, int mode, int R, int64_t nsamp, const int64_t *__restrict__ piv,
                                            int64_t off, double *__restrict__ K) {
    const int64_t total = nsamp * R;
    const int last = d.n - 1;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = idx % nsamp;
        const int r = (int)(idx / nsamp);
        const int64_t g = piv[s + nsamp * (int64_t)(last - 1)] - 1 - off;  // the last mode is the final pivot column
        double v = 0.0;
        if (g >= 0 && g < d.dim[last]) {
            v = 1.0;
            int col = 0;
            for (int m = 0; m < d.n; ++m) {
                if (m == mode) continue;
                const int64_t i = (m == last) ? g : piv[s + nsamp * col] - 1;
                v = v * fp.a[m][i + d.dim[m] * (int64_t)r];
                ++col;
            }
        }
        K[idx] = v;
    }
}

// out[i, s] = owned(s) ? T_slab[coords(s) with mode index i] : 0   (mode != last)
__global__ void __launch_bounds__(128) gather_fibers_owned_kernel(const double *__restrict__ T, ShDims d, int mode, int64_t nsamp,
                                                                  const int64_t *__restrict__ piv, int64_t off_last,
                                                                  double *__restrict__ out) {
    const int64_t s = blockIdx.x;
    const int last = d.n - 1;
    const int64_t I = d.dim[mode];
    const int64_t g = piv[s + nsamp * (int64_t)(last - 1)] - 1 - off_last;
    if (g < 0 || g >= d.dim[last]) {
        for (int64_t i = threadIdx.x; i < I; i += 128) out[i + I * s] = 0.0;
        return;
    }
    int64_t off = 0, str = 1, stride_mode = 1;
    int col = 0;
    for (int m = 0; m < d.n; ++m) {
        if (m == mode) stride_mode = str;
        else { off += ((m == last) ? g : piv[s + nsamp * col] - 1) * str; ++col; }
        str *= d.ext[m];
    }
    for (int64_t i = threadIdx.x; i < I; i += 128) out[i + I * s] = T[off + i * stride_mode];
}

int64_t sharded_last_rows(const itcpd_ctx *c) { return c->dims[c->order - 1] * (int64_t)comm_size(c); }
int64_t sharded_last_offset(const itcpd_ctx *c) { return c->dims[c->order - 1] * (int64_t)comm_rank(c); }

// leverage scores of one mode on a sharded handle; the sharded factor gets the all-reduced Gram and the GLOBAL row count
int sharded_leverage(itcpd_ctx *c, int mode) {
    if (c->lev_ver[mode] == c->fver[mode] && c->lev_ver[mode] != 0) return ITCPD_OK;
    const int R = c->rank;
    const bool last = mode == c->order - 1;
    TRY(k_gram(c, c->A[mode].as<double>(), c->dims[mode], R, c->G[mode].as<double>()));
    if (last) TRY(comm_allreduce_sum(c, c->G[mode].as<double>(), (int64_t)R * R));
    TRY(k_leverage_rows(c, c->A[mode].as<double>(), c->G[mode].as<double>(), c->dims[mode], last ? sharded_last_rows(c) : c->dims[mode], R,
                        c->lev[mode].as<double>()));
    c->lev_ver[mode] = c->fver[mode];
    return ITCPD_OK;
}

// weighted row sampling for every mode but `skip_mode`; identical pivots on every rank
int sharded_sample_rows(itcpd_ctx *c, int skip_mode, int64_t nsamp, uint64_t seed, int64_t *piv_dev) {
    int col = 0;
    for (int m = 0; m < c->order; ++m) {
        if (m == skip_mode) continue;
        TRY(sharded_leverage(c, m));
        if (m == c->order - 1) {
            const int64_t tot = sharded_last_rows(c);
            TRY(c->lev_gather.reserve((size_t)tot * 8));
            TRY(comm_allgather(c, c->lev[m].as<double>(), c->lev_gather.as<double>(), c->dims[m]));  // rank-major == row order
            TRY(k_cdf_sample(c, c->lev_gather.as<double>(), tot, nsamp, seed, (uint64_t)m, piv_dev + nsamp * col));
        } else {
            TRY(k_cdf_sample(c, c->lev[m].as<double>(), c->dims[m], nsamp, seed, (uint64_t)m, piv_dev + nsamp * col));
        }
        ++col;
    }
    return ITCPD_OK;
}

// One sampled update of `mode` on a sharded handle.  piv_dev: nsamp x (N-1) global pivots on the device.
// Ts_cached: the rank's cached sampled unfolding (projected algorithms) or nullptr to gather it now into samp_T.
int sharded_sampled_update(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv_dev, const double *Ts_cached, double chol_tol,
                           int normal, bool refresh_leverage) {
    const int N = c->order, R = c->rank, last = N - 1;
    const int64_t I = c->dims[mode];  // local rows when mode == last
    ShFac fp;
    for (int m = 0; m < N; ++m) fp.a[m] = c->A[m].as<double>();
    TRY(c->samp_K.reserve((size_t)nsamp * R * 8));
    double *K = c->samp_K.as<double>();
    const double *Ts = Ts_cached;
    c->m_valid[mode] = false;  // M[mode] is reused for the sampled MTTKRP
    if (mode != last) {
        const int64_t total = nsamp * R;
        pivot_hadamard_owned_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 16), 256, 0, c->stream>>>(
            fp, shdims(c), mode, R, nsamp, piv_dev, sharded_last_offset(c), K);
        c->launches++;
        if (!Ts) {
            TRY(c->samp_T.reserve((size_t)nsamp * I * 8));
            TRY(sharded_gather_fibers(c, mode, nsamp, piv_dev, c->samp_T.as<double>()));
            Ts = c->samp_T.as<double>();
        }
        CUDA_TRY(cudaGetLastError());
        if (normal) {
            TRY(k_gram(c, K, nsamp, R, c->Gamma.as<double>()));
            TRY(comm_allreduce_sum(c, c->Gamma.as<double>(), (int64_t)R * R));
            TRY(k_small_gemm_nn(c, Ts, K, I, nsamp, R, c->M[mode].as<double>()));
            TRY(comm_allreduce_sum(c, c->M[mode].as<double>(), I * R));
            TRY(k_solve(c, c->Gamma.as<double>(), c->M[mode].as<double>(), I, R, chol_tol, c->X.as<double>(), c->status.as<int>()));
        } else {
            ARG_CHECK(nsamp >= R && nsamp < ((int64_t)1 << 31), "normal=false needs at least R samples");
            // the full K and T_s on every rank: the owned pieces are disjoint, so a sum is a gather
            TRY(comm_allreduce_sum(c, K, nsamp * R));
            TRY(c->work2.reserve((size_t)nsamp * I * 8));
            CUDA_TRY(cudaMemcpyAsync(c->work2.p, Ts, (size_t)nsamp * I * 8, cudaMemcpyDeviceToDevice, c->stream));
            TRY(comm_allreduce_sum(c, c->work2.as<double>(), nsamp * I));
            TRY(qrcp_ls_solve(c, K, (int)nsamp, R, c->work2.as<double>(), I, c->X.as<double>(), c->status.as<int>(), 1));
        }
        TRY(k_colnorm_scale(c, c->X.as<double>(), I, R, c->A[mode].as<double>(), c->lambda.as<double>(), false));
    } else {
        TRY(k_pivot_hadamard(c, mode, nsamp, piv_dev, K));  // every other factor is replicated
        if (!Ts) {
            TRY(c->samp_T.reserve((size_t)nsamp * I * 8));
            TRY(k_gather_fibers(c, mode, nsamp, piv_dev, c->samp_T.as<double>()));  // fibres along the slab: my rows
            Ts = c->samp_T.as<double>();
        }
        if (normal) {
            TRY(k_gram(c, K, nsamp, R, c->Gamma.as<double>()));
            TRY(k_small_gemm_nn(c, Ts, K, I, nsamp, R, c->M[mode].as<double>()));
            TRY(k_solve(c, c->Gamma.as<double>(), c->M[mode].as<double>(), I, R, chol_tol, c->X.as<double>(), c->status.as<int>()));
        } else {
            ARG_CHECK(nsamp >= R && nsamp < ((int64_t)1 << 31), "normal=false needs at least R samples");
            TRY(qrcp_ls_solve(c, K, (int)nsamp, R, Ts, I, c->X.as<double>(), c->status.as<int>(), 1));
        }
        TRY(k_colnorm_scale(c, c->X.as<double>(), I, R, c->A[mode].as<double>(), c->lambda.as<double>(), true));  // all-reduced norms
    }
    c->fver[mode]++;
    if (refresh_leverage) {
        TRY(sharded_leverage(c, mode));  // also refreshes G[mode] (all-reduced for the sharded factor)
    } else {
        TRY(k_gram(c, c->A[mode].as<double>(), c->dims[mode], R, c->G[mode].as<double>()));
        if (mode == last) TRY(comm_allreduce_sum(c, c->G[mode].as<double>(), (int64_t)R * R));
    }
    return ITCPD_OK;
}

int sharded_gather_fibers(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv_dev, double *out_dev) {
    if (nsamp == 0) return ITCPD_OK;
    if (mode == c->order - 1) return k_gather_fibers(c, mode, nsamp, piv_dev, out_dev);
    gather_fibers_owned_kernel<<<(unsigned)nsamp, 128, 0, c->stream>>>(c->T.as<double>(), shdims(c), mode, nsamp, piv_dev,
                                                                      sharded_last_offset(c), out_dev);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

}  // namespace itcpd
