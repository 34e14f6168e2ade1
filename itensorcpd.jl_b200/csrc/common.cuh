// Shared declarations of libitcpd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/itcpd_b200.h"

#define ITCPD_MAX_ORDER 8

namespace itcpd {

void set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            itcpd::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return ITCPD_ERR_CUDA;                                                              \
        }                                                                                       \
    } while (0)

#define ARG_CHECK(cond, msg)                                         \
    do {                                                             \
        if (!(cond)) {                                               \
            itcpd::set_error("%s:%d %s", __FILE__, __LINE__, msg);   \
            return ITCPD_ERR_ARG;                                    \
        }                                                            \
    } while (0)

#define TRY(expr)                    \
    do {                             \
        int _s = (expr);             \
        if (_s != ITCPD_OK) return _s; \
    } while (0)

// A device buffer that grows on demand (never shrinks): all scratch lives behind the handle.
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int reserve(size_t n);
    void release();
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// One cached partial contraction of the dimension tree.
//   kind A: P[(i_0..i_{s-1}), r] = sum_{i_s..i_{N-1}} T * prod_{m>=s} A_m   (modes [0,s) stay free)
//   kind B: Q[(i_s..i_{N-1}), r] = sum_{i_0..i_{s-1}} T * prod_{m<s}  A_m   (modes [s,N) stay free)
struct Partial {
    DevBuf buf;
    bool valid = false;
    int split = 0;
    uint64_t dep_version[ITCPD_MAX_ORDER];  // versions of the contracted factors when computed
};

// gemm_i8.cu: cached row exponents of one unfolding of the tensor
struct I8ExpCache {
    DevBuf buf;
    bool valid = false;
    int split = 0;
    int64_t tensor_epoch = -1;
    int64_t nofit_epoch = -1;   // i8_apack: the digit planes of this tensor did not fit in HBM (do not retry every pass)
    int nofit_split = 0;
};

struct Comm;  // comm.cu

#define ITCPD_MAX_PEERS 16
// sources of a (possibly peer-reduced) right-hand-side matrix: n buffers summed in order; flags/epoch = the
// system-scope publication flags to wait for (null: no wait); reduced_out = where to store the reduced matrix
#define ITCPD_PEER_TIMEOUT_NS 60000000000ull  // 60 s: far above any legitimate skew between ranks of one sweep
struct PeerSrc {
    const double *p[ITCPD_MAX_PEERS];
    int n;
    const volatile long long *flags;
    long long epoch;
    const long long *epoch_dev;   // non-null: the epoch to wait for lives in device memory (capturable sweeps, peer_graph.cu)
    double *reduced_out;
};

// stream-K split description of one GEMM shape (host-built, cached per kind)
struct StreamKTable {
    int64_t key[6] = {-1, -1, -1, -1, -1, -1};
    int nsplit = 0, off_ptr = 0, off_list = 0;
    DevBuf dev;
};

}  // namespace itcpd

struct itcpd_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t hbm_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side_stream = nullptr;   // Gram-Hadamard + factorisation run here underneath the MTTKRP
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int overlap_factor = 1;
    // option "early_pass_b" (off by default, not yet run on hardware): pass B of the dimension tree only depends on the factors
    // of modes < split_b, so it is launched on its own stream as soon as they are updated and the updates of the modes in
    // [split_b, split_a) (second-level contraction from P_A, solve, normalise, Gram) run underneath it
    int early_pass_b = 0;
    cudaStream_t gemm_stream = nullptr;
    cudaEvent_t ev_gemm_fork = nullptr, ev_gemm_done = nullptr;
    bool gemm_join_pending = false;
    // chol_alg: 0 block kernel (any n); 1 team kernel for n <= 128 (same arithmetic, bitwise); 2 right-looking kernels for n <= 128;
    // 3 (default) right-looking where the factorisation is EXPOSED (no GEMM runs in that mode's update), where R > 64 (no factorisation
    // kernel can share an SM with a GEMM CTA) and where the GEMM pass is short (< ~1.4 ms: a co-resident team kernel then costs the
    // GEMM more than the 56 us it hides -- measured on the 8-GPU slabs); the team kernel (40 registers: co-resident) under long passes
    int chol_alg = 3;
    int64_t chol_short_gflop = 50;   // option "chol_short_gflop": passes below this many GFLOP count as short (0: never)
    bool chol_exposed = true;   // set by the sweep driver before every factorisation
    int64_t launches = 0;

    // options
    int mttkrp_alg = ITCPD_MTTKRP_TREE;
    int swizzle = 1;
    int tile_warps = 8;
    int stream_k = 1;  // split the tiles of the last partial wave along k (hybrid stream-K)
    int tma3d = 1;  // kind-0 tiles as one 3-D TMA box when the row count is a multiple of 16
    int force_split_a = 0, force_split_b = 0;

    // tensor
    int order = 0;
    int64_t dims[ITCPD_MAX_ORDER] = {0};
    int64_t ld0 = 0;    // leading dimension of mode 0 in device storage (dims[0] rounded up to even)
    int64_t nelem = 0;  // logical element count
    int64_t nstore = 0; // stored element count (ld0 * prod(dims[1:]))
    itcpd::DevBuf T;
    bool has_tensor = false;        // shape known (dims valid)
    bool has_tensor_data = false;   // dense data resident (false after itcpd_drop_tensor)

    // CPD
    int rank = 0;
    itcpd::DevBuf A[ITCPD_MAX_ORDER];   // I_n x R col-major
    itcpd::DevBuf G[ITCPD_MAX_ORDER];   // R x R
    itcpd::DevBuf M[ITCPD_MAX_ORDER];   // last MTTKRP of each mode, I_n x R
    itcpd::DevBuf X;                    // solve output (max I x R)
    itcpd::DevBuf lambda, Gamma, lev[ITCPD_MAX_ORDER];
    itcpd::DevBuf prevA[ITCPD_MAX_ORDER], prev_lambda;   // snapshot of the CPD (CPDiffCheck / CPAngleCheck PrevCP)
    bool has_snapshot = false;
    int snapshot_rank = 0;
    uint64_t lev_ver[ITCPD_MAX_ORDER] = {0};  // fver value the leverage scores were computed for (0 = never)
    uint64_t fver[ITCPD_MAX_ORDER] = {0};  // factor versions (bumped on every change)
    bool m_valid[ITCPD_MAX_ORDER] = {false};
    int last_mttkrp_mode = -1;

    // dimension tree
    itcpd::Partial PA, PB;
    itcpd::StreamKTable sk_table[2];
    itcpd::DevBuf sk_slots;
    int split_a = 0, split_b = 0;

    // scratch
    itcpd::DevBuf packK, krp_scratch[2], work, work2, redux, solve_ws, ipiv, status, fit2, samp_piv, samp_K, samp_T, flush;
    cudaEvent_t user_events[16] = {nullptr};
    // pivot-projected solvers: cached projector (1-based int64, nsamp x (N-1)) and sampled target (I_n x nsamp) per mode
    itcpd::DevBuf proj_piv[ITCPD_MAX_ORDER], proj_T[ITCPD_MAX_ORDER], qr_A, qr_piv, qr_rdiag;
    int64_t proj_n[ITCPD_MAX_ORDER] = {0};
    // option "staged_upload" (default on): a pageable host tensor is uploaded through pinned staging buffers filled by host threads
    int staged_upload = 1;
    // option "seqrcs_use_omega": candidate columns of itcpd_seqrcs in increasing order, as SEQRCS(...; use_omega = true) lists them
    // (SEQRCS.jl:109-113); off = the order of the matrix-free variant (:159), the reference's default
    int seqrcs_use_omega = 0;
    // option "sketch_unfold" (default on): the SE-QRCS set-up sketches modes >= 1 from an explicit unfolding when HBM has room for it
    int sketch_unfold = 1;
    void *upload_stage = nullptr;
    cudaEvent_t upload_events[8] = {nullptr};
    itcpd::DevBuf unfolded;   // explicit unfolding of the mode being sketched (SE-QRCS set-up, modes >= 1; released when the set-up returns)
    // pinned homes of the SE-QRCS embeddings' CSR entries (two: the helper thread fills one while the other is in use)
    void *sketch_pin[2] = {nullptr, nullptr};
    size_t sketch_pin_bytes[2] = {0, 0};
    double *pinned = nullptr;   // pinned host staging (fit scalars, status words)
    size_t pinned_doubles = 0;

    // GEMM timing (CUDA events on `stream`)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gemm_events;
    size_t gemm_events_used = 0;
    bool time_gemm = false;
    // optional per-phase CUDA events of the sweep driver (option "time_phases"; itcpd_phase_timing)
    bool time_phases = false;
    std::vector<cudaEvent_t> phase_events;
    std::vector<int> phase_ids;
    size_t phase_used = 0;

    // whole-sweep CUDA graph + device-side result log
    int use_graph = 1;
    int64_t graph_epoch = 0;
    cudaGraphExec_t sweep_graph_exec = nullptr;
    int64_t sweep_graph_key[24] = {0};
    // device-resident sweeps of the leverage-score sampled solver (itcpd_sampled_sweep_async): their own captured graph, and the
    // device-side draw counter that seeds every weighted draw (so that a replayed graph draws fresh samples)
    cudaGraphExec_t sampled_graph_exec = nullptr;
    int64_t sampled_graph_key[32] = {0};
    int64_t sampled_graph_launches = 0;
    int64_t sampled_plain_key[32] = {0};
    bool sampled_plain_key_valid = false;
    itcpd::DevBuf draw_counter;
    itcpd::DevBuf lev_q;   // leverage scores: Q_1 = A P U^{-1} (rows x R), G_2 = Q_1^T Q_1 and the Neumann correction W (solve.cu)
    // option "graph_single" (off by default, not yet run on hardware): the reference-facing loop calls itcpd_sweep(1) once per
    // iteration (optimize.jl:15-31), which never reaches the nsweeps >= 3 rule; with the option on, the second single-sweep call
    // with an unchanged configuration captures the graph and later calls replay it
    int graph_single = 1;
    int64_t plain_sweep_key[24] = {0};
    bool plain_sweep_key_valid = false;
    int64_t sweep_graph_launches = 0;
    itcpd::DevBuf sweep_log;      // [0] = counter (u64), then inner[cap] | norm2[cap] | fallbacks[cap]
    int64_t sweep_log_cap = 0;
    bool sweep_log_reduced = false;

    // multi-GPU
    itcpd::Comm *comm = nullptr;
    // peer-memory exchange (CUDA IPC): [flags: 16 x int64][pad to 256 B][partial-M buffer 0][partial-M buffer 1]
    itcpd::DevBuf xchg;
    void *peer_base[ITCPD_MAX_PEERS] = {nullptr};
    int peer_n = 0, peer_rank = 0;
    bool peer_on = false;
    long long peer_epoch = 0;
    int64_t peer_slot_doubles = 0;
    itcpd::DevBuf lev_gather;     // all-gathered leverage scores of the sharded factor (sampled_sharded.cu)
    // peer_graph.cu (option "peer_graph", off by default): device-side exchange epochs [0] partial-M, [1] small all-reduce;
    // small slots live behind the partial-M slots of the exchange buffer
    int peer_graph = 1;
    itcpd::DevBuf peer_epochs;
    size_t peer_small_off = 0;
    int64_t peer_small_doubles = 0;
    int peer_slots = 2;
    // gemm_i8.cu (option "gemm_i8", off by default): INT8 tensor-core digit-split contraction
    int gemm_i8 = 0;
    int64_t i8_tensor_epoch = 0;      // bumped whenever the tensor contents change
    itcpd::I8ExpCache i8_exp[2];
    itcpd::I8ExpCache i8_apack[2];    // gemm_i8 = 2: pre-packed digit planes of T per unfolding kind
    int i8_spare_sms = 0;   // option "i8_spare_sms": SMs the persistent INT8 GEMM leaves free for the side-stream factorisation (a 132 KB Cholesky at R = 128 cannot share an SM with it)
    itcpd::DevBuf i8_eb, i8_bdig, i8_part;   // i8_part: split-K partial tiles (short-and-wide contractions)
};

namespace itcpd {

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- gemm_dmma.cu -------------------------------------------------------------------------
// Partial contraction of the dimension tree (the dominant kernel).
//  kind 0 ("A"): out[m, r] = sum_k T[m + M*k] K[k, r]; m in [0,M) = stored modes [0,split), k = modes [split,N)
//  kind 1 ("B"): out[n, r] = sum_m T[m + Mc*n] K[m, r]; m = stored modes [0,split), n = modes [split,N)
int launch_partial_gemm(itcpd_ctx *c, int kind, int split, double *out);
int launch_partial_gemm_i8(itcpd_ctx *c, int kind, int split, double *out);  // gemm_i8.cu (experimental)
int probe_dmma(itcpd_ctx *c, double *tflops);
int probe_dfma(itcpd_ctx *c, double *tflops);

// ---- kernels.cu ---------------------------------------------------------------------------
int k_gram(itcpd_ctx *c, const double *A, int64_t rows, int R, double *G);
int k_sum_slices(itcpd_ctx *c, const double *part, int64_t n, int slices, double *out);
int k_cross_gram(itcpd_ctx *c, const double *A, const double *B, int64_t rows, int R, double *C);
int k_cpd_diff_terms(itcpd_ctx *c, double *out2);
int k_gram_hadamard(itcpd_ctx *c, int skip_mode, double *Gamma);
int k_colnorm_scale(itcpd_ctx *c, const double *X, int64_t rows, int R, double *A, double *lambda, bool rows_are_slab);
int k_fit_terms(itcpd_ctx *c, double *out2 /*device: inner, norm2*/, bool reduce_now);
int k_partial_mttkrp(itcpd_ctx *c, const double *P, int gfirst, int glast, int64_t ld_first, int mode, double *out);
int k_direct_mttkrp(itcpd_ctx *c, int mode, double *out);
int k_generate(itcpd_ctx *c, uint64_t seed, int64_t elem_offset);
int k_add_noise(itcpd_ctx *c, uint64_t seed, double sigma);
int k_randn_matrix(itcpd_ctx *c, double *dst, int64_t n, uint64_t seed, uint64_t stream_offset);
int k_sumsq(itcpd_ctx *c, const double *x, int64_t n, double *out_dev);
int k_pad_copy_in(itcpd_ctx *c, const double *src_dense, double *dst_padded);   // dims[0] -> ld0
int k_pad_copy_out(itcpd_ctx *c, const double *src_padded, double *dst_dense);
int k_reconstruct(itcpd_ctx *c, double *out_dense_or_null, double *resid_sumsq_dev_or_null);

// ---- solve.cu -----------------------------------------------------------------------------
// X (rows x R) = (Gamma \ M^T)^T with the ldiv_solve.jl semantics. status_dev[0]=path, [1]=rank.
int k_solve(itcpd_ctx *c, const double *Gamma, const double *M, int64_t rows, int R, double tol, double *X, int *status_dev);
int k_solve_factor(itcpd_ctx *c, const double *Gamma, int R, double tol, int *status_dev);
int k_solve_apply(itcpd_ctx *c, const double *Gamma, const double *M, int64_t rows, int R, double *X, int *status_dev);
int qrcp_ls_solve(itcpd_ctx *c, const double *A, int m, int n, const double *Bt, int64_t rows, double *X, int *status_dev, int force);  // qrcp.cu
int k_solve_apply_peers(itcpd_ctx *c, const double *Gamma, const PeerSrc &src, int64_t rows, int R, double *X, int *status_dev);
int k_leverage(itcpd_ctx *c, const double *A, const double *G, int64_t rows, int R, double *lev_out);
int k_leverage_rows(itcpd_ctx *c, const double *A, const double *G, int64_t rows_local, int64_t rows_total, int R, double *lev_out);

// ---- sampled.cu ---------------------------------------------------------------------------
int k_sample_rows(itcpd_ctx *c, int skip_mode, int64_t nsamp, uint64_t seed, int64_t *piv_dev, unsigned long long *draw_counter_dev = nullptr);
int k_sampled_mttkrp(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv_dev, const double *Ts_dev, const double *K_dev, double *M_dev);  // T_s K
int k_pivot_hadamard(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv_dev, double *K_dev);
int k_gather_fibers(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv_dev, double *out_dev);
int k_sketch_csr(itcpd_ctx *c, int mode, int l, int64_t nnz, const int64_t *row_ptr_dev, int64_t *col_dev, const double *val_dev, double *out_dev,
                 const double *unfolded_dev = nullptr);   // col_dev is overwritten (element offsets)
int k_qrcp_wide(itcpd_ctx *c, double *A, int64_t m, int64_t n, int64_t steps, int64_t *jpvt_dev, double *rdiag_dev);  // qrcp_wide.cu
int k_unfold(itcpd_ctx *c, int mode, double *out);
int k_omega_hadamard(itcpd_ctx *c, int mode, int l, const int64_t *row_ptr_dev, const int64_t *col_dev, const double *val_dev, double *out_dev);
int k_pivot_hadamard_t(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv_dev, double *out_dev);
int k_small_gemm_nn(itcpd_ctx *c, const double *A, const double *B, int64_t m, int64_t k, int n, double *C); // C = A B
int k_cdf_sample(itcpd_ctx *c, const double *weights_dev, int64_t n, int64_t nsamp, uint64_t seed, uint64_t stream_id, int64_t *out_dev);

// ---- sampled_sharded.cu (slab-sharded sampled path) ----------------------------------------
int sharded_leverage(itcpd_ctx *c, int mode);
int sharded_sample_rows(itcpd_ctx *c, int skip_mode, int64_t nsamp, uint64_t seed, int64_t *piv_dev);
int sharded_gather_fibers(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv_dev, double *out_dev);
int sharded_sampled_update(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv_dev, const double *Ts_cached, double chol_tol,
                           int normal, bool refresh_leverage);
int64_t sharded_last_rows(const itcpd_ctx *c);
int64_t sharded_last_offset(const itcpd_ctx *c);

// ---- peer_graph.cu ------------------------------------------------------------------------
bool peer_graph_active(const itcpd_ctx *c);
int peer_graph_signal(itcpd_ctx *c);
int peer_graph_wait(itcpd_ctx *c);
int peer_allreduce_small(itcpd_ctx *c, double *buf, int64_t n);

// ---- comm.cu ------------------------------------------------------------------------------
int comm_allreduce_sum(itcpd_ctx *c, double *buf, int64_t n);
int comm_allgather(itcpd_ctx *c, const double *send, double *recv, int64_t n_per_rank);
bool comm_active(const itcpd_ctx *c);
int comm_rank(const itcpd_ctx *c);
int comm_size(const itcpd_ctx *c);

// ---- api.cu helpers -----------------------------------------------------------------------
int ensure_cpd_buffers(itcpd_ctx *c);
int64_t mode_rows(const itcpd_ctx *c, int mode);
void choose_splits(itcpd_ctx *c);

}  // namespace itcpd
