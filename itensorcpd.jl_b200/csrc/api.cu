// extern "C" entry points of libitcpd_b200 (see include/itcpd_b200.h for the contract and the
// reference interfaces each one replaces).
#include "common.cuh"
#include <chrono>
#include <cmath>
#include <functional>
#include <cstdarg>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <thread>

namespace itcpd {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int DevBuf::reserve(size_t n) {
    if (n <= bytes && p) return ITCPD_OK;
    release();
    if (n == 0) n = 256;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", n, cudaGetErrorString(e));
        cudaGetLastError();  // the failed allocation is also the runtime's "last error": clear it, or the next launch check reports it again
        p = nullptr;
        bytes = 0;
        return ITCPD_ERR_CUDA;
    }
    bytes = n;
    return ITCPD_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
}

int64_t mode_rows(const itcpd_ctx *c, int mode) { return c->dims[mode]; }

static int ensure_pinned(itcpd_ctx *c, size_t doubles) {
    if (doubles <= c->pinned_doubles) return ITCPD_OK;
    if (c->pinned) { CUDA_TRY(cudaStreamSynchronize(c->stream)); cudaFreeHost(c->pinned); c->pinned = nullptr; }
    size_t n = std::max<size_t>(doubles, 4096);
    CUDA_TRY(cudaHostAlloc((void **)&c->pinned, n * 8, cudaHostAllocDefault));
    c->pinned_doubles = n;
    return ITCPD_OK;
}

// Split points of the dimension tree (DESIGN.md "dimension tree"): pass A keeps modes [0,sa) free,
// pass B keeps modes [sb,N) free, sb <= sa; both GEMM outputs should have >= 16384 rows when possible.
void choose_splits(itcpd_ctx *c) {
    // Every (sa, sb) costs the same two tensor passes (2*R*P flops, 8*P bytes each); what differs is the HBM
    // traffic of the intermediates, in units of R doubles:
    //   pass-A operand (Khatri-Rao of modes >= sa, written by the pack kernel and read back): 2 * prod_{n>=sa} I_n
    //   partial P_A (written once, read once per mode updated from it):                       rows_A * (1 + sa)
    //   pass-B operand / partial P_B likewise with the modes < sb contracted and N - sa modes served.
    // A group of one mode needs no second level (its partial IS the MTTKRP): 2 * rows.  Ties go to the larger sa,
    // then the smaller sb.  With slab sharding the LOCAL extents are used, which is what moves an order-3 tensor
    // from (2,1) on one GPU to (1,1) once the last mode is split over several GPUs.
    const int N = c->order;
    double best = -1.0;
    int sa = N - 1, sb = 1;
    for (int a = 1; a <= N - 1; ++a) {
        for (int b = 1; b <= a; ++b) {
            double rowsA = (double)c->ld0, kA = 1.0, rowsB = 1.0, kB = (double)c->ld0;
            for (int n = 1; n < a; ++n) rowsA *= (double)c->dims[n];
            for (int n = a; n < N; ++n) kA *= (double)c->dims[n];
            for (int n = b; n < N; ++n) rowsB *= (double)c->dims[n];
            for (int n = 1; n < b; ++n) kB *= (double)c->dims[n];
            const double costA = 2.0 * kA + rowsA * (a == 1 ? 2.0 : 1.0 + a);
            const int servedB = N - a;
            const double costB = 2.0 * kB + rowsB * ((N - b) == 1 ? 2.0 : 1.0 + servedB);
            const double cost = costA + costB;
            if (best < 0.0 || cost < best * (1.0 - 1e-9) || (cost <= best * (1.0 + 1e-9) && (a > sa || (a == sa && b < sb)))) {
                best = cost; sa = a; sb = b;
            }
        }
    }
    if (c->force_split_a >= 1 && c->force_split_a <= N - 1) sa = c->force_split_a;
    if (c->force_split_b >= 1 && c->force_split_b <= sa) sb = c->force_split_b;
    if (sb > sa) sb = sa;
    c->split_a = sa;
    c->split_b = sb;
    c->PA.valid = c->PB.valid = false;
}

int ensure_cpd_buffers(itcpd_ctx *c) {
    ARG_CHECK(c->has_tensor && c->rank > 0, "set the tensor and the rank first");
    const int R = c->rank;
    int64_t maxrows = 0;
    for (int n = 0; n < c->order; ++n) {
        const int64_t rows = c->dims[n];
        maxrows = std::max(maxrows, rows);
        TRY(c->A[n].reserve((size_t)rows * R * 8));
        TRY(c->M[n].reserve((size_t)rows * R * 8));
        TRY(c->G[n].reserve((size_t)R * R * 8));
        TRY(c->lev[n].reserve((size_t)rows * 8));
    }
    TRY(c->X.reserve((size_t)maxrows * R * 8));
    TRY(c->lambda.reserve((size_t)R * 8));
    TRY(c->Gamma.reserve((size_t)R * R * 8));
    TRY(c->status.reserve(256));
    TRY(c->fit2.reserve(64));
    return ITCPD_OK;
}


// Host -> device upload of the tensor.  A pinned source (itcpd_host_alloc, cudaHostRegister'ed memory) is one DMA; a PAGEABLE source
// -- what a Julia Array or a numpy array is -- would make the driver stage it through its own small bounce buffer at ~12 GB/s.
// Here UPLOAD_THREADS host threads copy 32 MB chunks into their own pair of pinned staging buffers while the DMA engine drains the
// previous chunks, so the upload runs at min(parallel memcpy, PCIe) instead (bench.py `e2e.pageable`).  The copies are enqueued on the
// handle's stream (stream order protects every later kernel); the call returns when the last chunk has been ENQUEUED and staged.
constexpr size_t UPLOAD_CHUNK = (size_t)32 << 20;
constexpr int UPLOAD_THREADS = 4;

static int upload_to_device(itcpd_ctx *c, void *dst, const void *host, size_t bytes) {
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, host) == cudaSuccess && (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    if (pinned || !c->staged_upload || bytes < 4 * UPLOAD_CHUNK) {
        CUDA_TRY(cudaMemcpyAsync(dst, host, bytes, cudaMemcpyHostToDevice, c->stream));
        return ITCPD_OK;
    }
    const int nslots = 2 * UPLOAD_THREADS;
    if (!c->upload_stage) {
        CUDA_TRY(cudaHostAlloc(&c->upload_stage, (size_t)nslots * UPLOAD_CHUNK, cudaHostAllocDefault));
        for (int i = 0; i < nslots; ++i) CUDA_TRY(cudaEventCreateWithFlags(&c->upload_events[i], cudaEventDisableTiming));
    }
    const size_t nchunks = (bytes + UPLOAD_CHUNK - 1) / UPLOAD_CHUNK;
    std::atomic<int> failed{0};
    auto worker = [&](int t) {
        if (cudaSetDevice(c->device) != cudaSuccess) { failed = 1; return; }
        int use = 0;
        for (size_t ch = (size_t)t; ch < nchunks && !failed; ch += UPLOAD_THREADS, ++use) {
            const int slot = 2 * t + (use & 1);
            char *stage = (char *)c->upload_stage + (size_t)slot * UPLOAD_CHUNK;
            if (use >= 2 && cudaEventSynchronize(c->upload_events[slot]) != cudaSuccess) { failed = 1; return; }   // the slot's previous DMA is done
            const size_t off = ch * UPLOAD_CHUNK, n = std::min(UPLOAD_CHUNK, bytes - off);
            memcpy(stage, (const char *)host + off, n);
            if (cudaMemcpyAsync((char *)dst + off, stage, n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                cudaEventRecord(c->upload_events[slot], c->stream) != cudaSuccess) { failed = 1; return; }
        }
    };
    std::thread th[UPLOAD_THREADS];
    for (int t = 0; t < UPLOAD_THREADS; ++t) th[t] = std::thread(worker, t);
    for (int t = 0; t < UPLOAD_THREADS; ++t) th[t].join();
    if (failed) { set_error("staged upload failed: %s", cudaGetErrorString(cudaGetLastError())); return ITCPD_ERR_CUDA; }
    return ITCPD_OK;
}

static void invalidate_all(itcpd_ctx *c) {
    c->graph_epoch++;
    c->i8_tensor_epoch++;
    c->PA.valid = c->PB.valid = false;
    for (int n = 0; n < ITCPD_MAX_ORDER; ++n) { c->m_valid[n] = false; c->fver[n]++; }
}

static int set_shape(itcpd_ctx *c, int order, const int64_t *dims, bool allocate = true) {
    ARG_CHECK(order >= 2 && order <= ITCPD_MAX_ORDER, "tensor order must be in [2,8]");
    int64_t n = 1;
    for (int i = 0; i < order; ++i) {
        ARG_CHECK(dims[i] >= 1, "tensor dimensions must be positive");
        n *= dims[i];
    }
    c->order = order;
    for (int i = 0; i < order; ++i) c->dims[i] = dims[i];
    c->ld0 = dims[0] + (dims[0] & 1);
    c->nelem = n;
    c->nstore = n / dims[0] * c->ld0;
    if (allocate) TRY(c->T.reserve((size_t)c->nstore * 8 + 256));
    c->has_tensor = true;
    c->has_tensor_data = allocate;
    for (int n = 0; n < ITCPD_MAX_ORDER; ++n) c->proj_n[n] = 0;  // cached projectors belong to the previous tensor
    choose_splits(c);
    invalidate_all(c);
    if (c->rank > 0) TRY(ensure_cpd_buffers(c));
    return ITCPD_OK;
}

// is the cached partial of this kind still the contraction of the CURRENT factors?
static bool partial_is_current(const itcpd_ctx *c, int kind) {
    const Partial &P = (kind == 0) ? c->PA : c->PB;
    const int split = (kind == 0) ? c->split_a : c->split_b;
    const int d0 = (kind == 0) ? split : 0, d1 = (kind == 0) ? c->order : split;
    bool ok = P.valid && P.split == split;
    for (int n = d0; ok && n < d1; ++n) ok = (P.dep_version[n] == c->fver[n]);
    return ok;
}

// the partial contraction feeding `mode`, recomputed only when a contracted factor changed
static int ensure_partial(itcpd_ctx *c, int kind) {
    Partial &P = (kind == 0) ? c->PA : c->PB;
    const int split = (kind == 0) ? c->split_a : c->split_b;
    const int d0 = (kind == 0) ? split : 0, d1 = (kind == 0) ? c->order : split;
    if (partial_is_current(c, kind)) return ITCPD_OK;
    int64_t rows = 1;
    if (kind == 0) { rows = c->ld0; for (int n = 1; n < split; ++n) rows *= c->dims[n]; }
    else { for (int n = split; n < c->order; ++n) rows *= c->dims[n]; }
    TRY(P.buf.reserve((size_t)rows * c->rank * 8));
    int st = ITCPD_ERR_UNSUPPORTED;
    if (c->gemm_i8) {  // experimental INT8 tensor-core path; shapes outside its envelope fall through to the DMMA kernel
        st = launch_partial_gemm_i8(c, kind, split, P.buf.as<double>());
        if (st != ITCPD_OK && st != ITCPD_ERR_UNSUPPORTED) return st;
    }
    if (st != ITCPD_OK) TRY(launch_partial_gemm(c, kind, split, P.buf.as<double>()));
    P.valid = true;
    P.split = split;
    for (int n = d0; n < d1; ++n) P.dep_version[n] = c->fver[n];
    return ITCPD_OK;
}

static int mttkrp_device(itcpd_ctx *c, int mode, double *out_override = nullptr, bool reduce = true) {
    double *out = out_override ? out_override : c->M[mode].as<double>();
    if (c->mttkrp_alg == ITCPD_MTTKRP_DIRECT) {
        TRY(k_direct_mttkrp(c, mode, out));
    } else if (mode < c->split_a) {
        TRY(ensure_partial(c, 0));
        TRY(k_partial_mttkrp(c, c->PA.buf.as<double>(), 0, c->split_a - 1, c->ld0, mode, out));
    } else {
        TRY(ensure_partial(c, 1));
        TRY(k_partial_mttkrp(c, c->PB.buf.as<double>(), c->split_b, c->order - 1, c->dims[c->split_b], mode, out));
    }
    // slab sharding: T is a slab of the last mode, so every other mode's MTTKRP is a partial sum
    if (reduce && comm_active(c) && mode != c->order - 1) TRY(comm_allreduce_sum(c, out, c->dims[mode] * c->rank));
    c->m_valid[mode] = true;
    c->last_mttkrp_mode = mode;
    return ITCPD_OK;
}

static int gram_device(itcpd_ctx *c, int mode) {
    TRY(k_gram(c, c->A[mode].as<double>(), c->dims[mode], c->rank, c->G[mode].as<double>()));
    if (comm_active(c) && mode == c->order - 1) TRY(comm_allreduce_sum(c, c->G[mode].as<double>(), (int64_t)c->rank * c->rank));
    return ITCPD_OK;
}

// publish "my partial of exchange `epoch` is complete" into every peer's flag array (system scope)
struct PeerFlags { long long *dst[ITCPD_MAX_PEERS]; int n; int rank; };
__global__ void peer_signal_kernel(PeerFlags f, long long epoch) {
    __threadfence_system();
    if ((int)threadIdx.x < f.n) {
        volatile long long *d = f.dst[threadIdx.x] + f.rank;
        *d = epoch;
    }
    __threadfence_system();
}
static int peer_signal(itcpd_ctx *c, long long epoch) {
    PeerFlags f;
    memset(&f, 0, sizeof(f));
    f.n = c->peer_n;
    f.rank = c->peer_rank;
    for (int q = 0; q < c->peer_n; ++q) f.dst[q] = reinterpret_cast<long long *>(c->peer_base[q]);
    peer_signal_kernel<<<1, 32, 0, c->stream>>>(f, epoch);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

// phase ids of itcpd_phase_timing: time between two marks is charged to the later mark
enum { PH_MODE_BEGIN = 0, PH_MTTKRP = 1, PH_SIGNAL = 2, PH_SOLVE = 3, PH_NORMALIZE = 4, PH_GRAM = 5, PH_FIT = 6, PH_COUNT = 7 };
static int phase_mark(itcpd_ctx *c, int id) {
    if (!c->time_phases) return ITCPD_OK;
    if (c->phase_used == c->phase_events.size()) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        c->phase_events.push_back(e);
        c->phase_ids.push_back(0);
    }
    c->phase_ids[c->phase_used] = id;
    CUDA_TRY(cudaEventRecord(c->phase_events[c->phase_used], c->stream));
    c->phase_used++;
    return ITCPD_OK;
}

static int mode_update_device(itcpd_ctx *c, int mode, double tol, int *status_dev) {
    TRY(phase_mark(c, PH_MODE_BEGIN));
    // fork: Gram-Hadamard + pivoted Cholesky depend only on the Grams, so they run on the side stream
    // underneath the MTTKRP (the persistent GEMM leaves room for one small CTA); join before the row solves
    cudaStream_t main_stream = c->stream;
    if (c->overlap_factor) {
        CUDA_TRY(cudaEventRecord(c->ev_fork, main_stream));
        CUDA_TRY(cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
        c->stream = c->side_stream;
    }
    // chol_alg = 3: the factorisation hides under a GEMM if one runs in this mode's update (or pass B is in flight: early_pass_b)
    // a factorisation that would hide under a SHORT pass is treated as exposed: 2 R P_local flops at ~35 TFLOP/s below ~1.4 ms
    const bool short_pass = 2.0 * (double)c->rank * (double)c->nstore < 1e9 * (double)c->chol_short_gflop;
    c->chol_exposed = short_pass || !(c->overlap_factor && (c->gemm_join_pending ||
                                              (c->mttkrp_alg != ITCPD_MTTKRP_DIRECT && !partial_is_current(c, mode < c->split_a ? 0 : 1))));
    int st = k_gram_hadamard(c, mode, c->Gamma.as<double>());
    if (st == ITCPD_OK) st = k_solve_factor(c, c->Gamma.as<double>(), c->rank, tol, status_dev);
    // every factorisation outside this driver (itcpd_solve, the sampled solves, the leverage refresh) has nothing to hide under: leave
    // the flag at its default, or those callers inherit whatever the last mode of the last dense sweep decided (after a full-size
    // sweep: "hidden" -> the 96 us team kernel instead of the 56 us right-looking one, 0.26 ms per sampled sweep)
    c->chol_exposed = true;
    c->stream = main_stream;
    TRY(st);
    if (c->overlap_factor) CUDA_TRY(cudaEventRecord(c->ev_join, c->side_stream));
    if (c->gemm_join_pending && mode >= c->split_a) {  // early_pass_b: P_B is produced on the GEMM stream
        CUDA_TRY(cudaStreamWaitEvent(main_stream, c->ev_gemm_done, 0));
        c->gemm_join_pending = false;
    }
    const bool fused_peers = c->peer_on && comm_active(c) && mode != c->order - 1;
    if (fused_peers) {
        // fused all-reduce + solve over NVLink peer memory: the partial MTTKRP lands in this rank's exchange slot, is
        // published with a system-scope flag, and every rank's row-solve kernel sums the peers' slots while loading
        ARG_CHECK(c->dims[mode] * (int64_t)c->rank <= c->peer_slot_doubles,
                  "the peer exchange buffer was exported for a smaller shape/rank: call itcpd_peer_export/import again");
        const bool dev_epoch = peer_graph_active(c);  // capturable variant: epoch in device memory, slot = exchange index
        const int64_t epoch = dev_epoch ? 0 : ++c->peer_epoch;
        const size_t slot_off = 256 + (size_t)(dev_epoch ? mode : (epoch & 1)) * (size_t)c->peer_slot_doubles * 8;
        TRY(mttkrp_device(c, mode, reinterpret_cast<double *>((char *)c->xchg.p + slot_off), false));
        TRY(phase_mark(c, PH_MTTKRP));
        if (dev_epoch) TRY(peer_graph_signal(c));
        else TRY(peer_signal(c, epoch));
        TRY(phase_mark(c, PH_SIGNAL));
        if (c->overlap_factor) CUDA_TRY(cudaStreamWaitEvent(main_stream, c->ev_join, 0));
        PeerSrc src;
        memset(&src, 0, sizeof(src));
        src.n = c->peer_n;
        for (int q = 0; q < c->peer_n; ++q) src.p[q] = reinterpret_cast<const double *>((const char *)c->peer_base[q] + slot_off);
        src.flags = reinterpret_cast<const volatile long long *>(c->xchg.p);
        src.epoch = epoch;
        src.epoch_dev = dev_epoch ? c->peer_epochs.as<long long>() : nullptr;   // capturable sweeps: the signal kernel advanced it
        src.reduced_out = c->M[mode].as<double>();
        TRY(k_solve_apply_peers(c, c->Gamma.as<double>(), src, c->dims[mode], c->rank, c->X.as<double>(), status_dev));
    } else {
        TRY(mttkrp_device(c, mode));
        TRY(phase_mark(c, PH_MTTKRP));
        if (c->overlap_factor) CUDA_TRY(cudaStreamWaitEvent(main_stream, c->ev_join, 0));
        TRY(k_solve_apply(c, c->Gamma.as<double>(), c->M[mode].as<double>(), c->dims[mode], c->rank, c->X.as<double>(), status_dev));
    }
    TRY(phase_mark(c, PH_SOLVE));
    TRY(k_colnorm_scale(c, c->X.as<double>(), c->dims[mode], c->rank, c->A[mode].as<double>(), c->lambda.as<double>(), mode == c->order - 1));
    TRY(phase_mark(c, PH_NORMALIZE));
    c->fver[mode]++;
    TRY(gram_device(c, mode));
    TRY(phase_mark(c, PH_GRAM));
    return ITCPD_OK;
}

}  // namespace itcpd

using namespace itcpd;

#define CHECK_CTX(c) ARG_CHECK((c) != nullptr, "null context")
#define CHECK_MODE(c, m) ARG_CHECK((m) >= 0 && (m) < (c)->order, "mode out of range (0-based)")
#define USE_DEVICE(c) CUDA_TRY(cudaSetDevice((c)->device))
#define NEED_T(c) ARG_CHECK((c)->has_tensor && (c)->has_tensor_data, "the dense tensor is not resident (never set, or released by itcpd_drop_tensor)")

extern "C" {

int itcpd_version(void) { return 100; }
const char *itcpd_last_error(void) { return g_err; }

int itcpd_create(itcpd_ctx **out, int device) {
    ARG_CHECK(out != nullptr, "null out pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device visible (%s); libitcpd_b200 has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
        return ITCPD_ERR_NO_DEVICE;
    }
    ARG_CHECK(device >= 0 && device < ndev, "device index out of range");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libitcpd_b200 is built for sm_100a only and has no fallback path", device, prop.major, prop.minor);
        return ITCPD_ERR_NO_DEVICE;
    }
    CUDA_TRY(cudaSetDevice(device));
    itcpd_ctx *c = new itcpd_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->hbm_bytes = prop.totalGlobalMem;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    {   // early_pass_b: the pass-B GEMM must win the SMs over the small, many-CTA kernels of the modes updated underneath it,
        // or it only starts once they have drained (measured: no gain without the priority)
        int lo = 0, hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&c->gemm_stream, cudaStreamNonBlocking, hi));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_gemm_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_gemm_done, cudaEventDisableTiming));
    if (const char *s = getenv("ITCPD_EARLY_B")) c->early_pass_b = atoi(s) != 0;   // experimental
    if (const char *s = getenv("ITCPD_GRAPH_SINGLE")) c->graph_single = atoi(s) != 0;   // experimental
    if (const char *s = getenv("ITCPD_I8_SPARE_SMS")) c->i8_spare_sms = std::min(63, std::max(0, atoi(s)));
    if (const char *s = getenv("ITCPD_NO_SWIZZLE")) c->swizzle = (atoi(s) != 0) ? 0 : 1;
    if (const char *s = getenv("ITCPD_CHOL")) c->chol_alg = std::min(3, std::max(0, atoi(s)));
    if (const char *s = getenv("ITCPD_NO_GRAPH")) c->use_graph = atoi(s) == 0;
    if (const char *s = getenv("ITCPD_GEMM_I8")) c->gemm_i8 = std::min(2, std::max(0, atoi(s)));   // experimental (csrc/gemm_i8.cu)
    int st = ensure_pinned(c, 4096);
    if (st != ITCPD_OK) { delete c; return st; }
    *out = c;
    return ITCPD_OK;
}

int itcpd_destroy(itcpd_ctx *c) {
    if (!c) return ITCPD_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    itcpd_peer_disable(c);
    itcpd_comm_destroy(c);
    c->xchg.release();
    DevBuf *bufs[] = {&c->T, &c->X, &c->lambda, &c->Gamma, &c->PA.buf, &c->PB.buf, &c->packK, &c->krp_scratch[0], &c->krp_scratch[1],
                      &c->work, &c->work2, &c->redux, &c->solve_ws, &c->ipiv, &c->status, &c->fit2, &c->samp_piv, &c->samp_K, &c->samp_T, &c->flush, &c->sk_slots, &c->sk_table[0].dev, &c->sk_table[1].dev, &c->qr_A, &c->qr_piv, &c->qr_rdiag, &c->sweep_log};
    for (DevBuf *b : bufs) b->release();
    for (int n = 0; n < ITCPD_MAX_ORDER; ++n) { c->A[n].release(); c->G[n].release(); c->M[n].release(); c->lev[n].release(); c->proj_piv[n].release(); c->proj_T[n].release(); c->prevA[n].release(); }
    c->prev_lambda.release();
    c->lev_gather.release();
    c->peer_epochs.release();
    c->i8_exp[0].buf.release(); c->i8_exp[1].buf.release(); c->i8_eb.release(); c->i8_bdig.release(); c->i8_part.release();
    c->i8_apack[0].buf.release(); c->i8_apack[1].buf.release();
    for (auto &ev : c->gemm_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (auto &ev : c->phase_events) cudaEventDestroy(ev);
    for (auto &ev : c->user_events) if (ev) cudaEventDestroy(ev);
    if (c->pinned) cudaFreeHost(c->pinned);
    for (auto &pin : c->sketch_pin) if (pin) cudaFreeHost(pin);
    c->unfolded.release();
    if (c->upload_stage) {
        cudaFreeHost(c->upload_stage);
        for (auto &ev : c->upload_events) if (ev) cudaEventDestroy(ev);
    }
    if (c->sweep_graph_exec) { cudaGraphExecDestroy(c->sweep_graph_exec); c->sweep_graph_exec = nullptr; }
    if (c->sampled_graph_exec) { cudaGraphExecDestroy(c->sampled_graph_exec); c->sampled_graph_exec = nullptr; }
    c->draw_counter.release();
    c->lev_q.release();
    cudaStreamSynchronize(c->side_stream);
    cudaEventDestroy(c->ev_fork);
    cudaEventDestroy(c->ev_join);
    cudaStreamDestroy(c->side_stream);
    if (c->gemm_stream) { cudaStreamSynchronize(c->gemm_stream); cudaStreamDestroy(c->gemm_stream); }
    if (c->ev_gemm_fork) cudaEventDestroy(c->ev_gemm_fork);
    if (c->ev_gemm_done) cudaEventDestroy(c->ev_gemm_done);
    cudaStreamDestroy(c->stream);
    delete c;
    return ITCPD_OK;
}

int itcpd_device_info(itcpd_ctx *c, int *sm_count, int *cc_major, int *cc_minor, int64_t *hbm_bytes) {
    CHECK_CTX(c);
    if (sm_count) *sm_count = c->sm_count;
    if (cc_major) *cc_major = c->cc_major;
    if (cc_minor) *cc_minor = c->cc_minor;
    if (hbm_bytes) *hbm_bytes = (int64_t)c->hbm_bytes;
    return ITCPD_OK;
}

int itcpd_synchronize(itcpd_ctx *c) {
    CHECK_CTX(c);
    USE_DEVICE(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int64_t itcpd_launch_count(itcpd_ctx *c) { return c ? c->launches : -1; }

int itcpd_set_option(itcpd_ctx *c, const char *name, int64_t value) {
    CHECK_CTX(c);
    ARG_CHECK(name != nullptr, "null option name");
    std::string n(name);
    if (n == "mttkrp_alg") { ARG_CHECK(value == 0 || value == 1, "mttkrp_alg must be 0 or 1"); c->mttkrp_alg = (int)value; }
    else if (n == "swizzle") c->swizzle = value != 0;
    else if (n == "tile_warps") { ARG_CHECK(value == 4 || value == 8, "tile_warps must be 4 or 8"); c->tile_warps = (int)value; }
    else if (n == "split_a") { c->force_split_a = (int)value; if (c->has_tensor) choose_splits(c); }
    else if (n == "split_b") { c->force_split_b = (int)value; if (c->has_tensor) choose_splits(c); }
    else if (n == "time_gemm") c->time_gemm = value != 0;
    else if (n == "time_phases") { c->time_phases = value != 0; c->phase_used = 0; }
    else if (n == "tma3d") c->tma3d = value != 0;
    else if (n == "overlap_factor") c->overlap_factor = value != 0;
    else if (n == "staged_upload") c->staged_upload = value != 0;
    else if (n == "seqrcs_use_omega") c->seqrcs_use_omega = value != 0;
    else if (n == "sketch_unfold") { ARG_CHECK(value >= 0 && value <= 2, "sketch_unfold must be 0, 1 or 2"); c->sketch_unfold = (int)value; }
    else if (n == "early_pass_b") c->early_pass_b = value != 0;
    else if (n == "graph_single") c->graph_single = value != 0;
    else if (n == "i8_spare_sms") { ARG_CHECK(value >= 0 && value < 64, "i8_spare_sms must be in [0, 64)"); c->i8_spare_sms = (int)value; }
    else if (n == "use_graph") c->use_graph = value != 0;
    else if (n == "gemm_i8") { ARG_CHECK(value >= 0 && value <= 2, "gemm_i8 must be 0, 1 (convert on the fly) or 2 (pre-packed digits)"); c->gemm_i8 = (int)value; }
    else if (n == "peer_graph") {
        ARG_CHECK(!c->peer_on, "set peer_graph before itcpd_peer_export / itcpd_peer_import");
        c->peer_graph = value != 0;
    }
    else if (n == "chol_alg") { ARG_CHECK(value >= 0 && value <= 3, "chol_alg must be 0, 1, 2 or 3"); c->chol_alg = (int)value; }
    else if (n == "chol_short_gflop") { ARG_CHECK(value >= 0, "chol_short_gflop must be non-negative"); c->chol_short_gflop = value; }
    else if (n == "stream_k") { ARG_CHECK(value >= 0 && value <= 2, "stream_k must be 0, 1 or 2"); c->stream_k = (int)value; }
    else { set_error("unknown option '%s'", name); return ITCPD_ERR_ARG; }
    c->graph_epoch++;
    return ITCPD_OK;
}

// ---- tensor -----------------------------------------------------------------------------------
int itcpd_set_tensor(itcpd_ctx *c, int order, const int64_t *dims, const double *host) {
    CHECK_CTX(c);
    ARG_CHECK(dims && host, "null argument");
    USE_DEVICE(c);
    TRY(set_shape(c, order, dims));
    if (c->ld0 == c->dims[0]) {
        TRY(upload_to_device(c, c->T.p, host, (size_t)c->nelem * 8));
    } else {
        TRY(c->work.reserve((size_t)c->nelem * 8));
        TRY(upload_to_device(c, c->work.p, host, (size_t)c->nelem * 8));
        TRY(k_pad_copy_in(c, c->work.as<double>(), c->T.as<double>()));
    }
    CUDA_TRY(cudaMemsetAsync((char *)c->T.p + (size_t)c->nstore * 8, 0, 256, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_set_shape(itcpd_ctx *c, int order, const int64_t *dims) {
    CHECK_CTX(c);
    ARG_CHECK(dims != nullptr, "null dims");
    USE_DEVICE(c);
    return set_shape(c, order, dims, false);
}

int itcpd_generate_tensor(itcpd_ctx *c, int order, const int64_t *dims, uint64_t seed, int64_t elem_offset) {
    CHECK_CTX(c);
    ARG_CHECK(dims != nullptr, "null dims");
    USE_DEVICE(c);
    TRY(set_shape(c, order, dims));
    TRY(k_generate(c, seed, elem_offset));
    CUDA_TRY(cudaMemsetAsync((char *)c->T.p + (size_t)c->nstore * 8, 0, 256, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_generate_lowrank_tensor(itcpd_ctx *c, int order, const int64_t *dims, int rank, uint64_t seed, double noise) {
    CHECK_CTX(c);
    ARG_CHECK(dims != nullptr && rank >= 1, "null dims / bad rank");
    USE_DEVICE(c);
    TRY(set_shape(c, order, dims));
    c->rank = rank;
    TRY(ensure_cpd_buffers(c));
    TRY(itcpd_random_cpd(c, seed));
    {   // lambda = 1: the planted tensor is the plain sum of the rank-one terms
        std::vector<double> ones((size_t)rank, 1.0);
        CUDA_TRY(cudaMemcpyAsync(c->lambda.p, ones.data(), (size_t)rank * 8, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    if (c->ld0 == c->dims[0]) {
        TRY(k_reconstruct(c, c->T.as<double>(), nullptr));
    } else {
        TRY(c->work.reserve((size_t)c->nelem * 8));
        TRY(k_reconstruct(c, c->work.as<double>(), nullptr));
        TRY(k_pad_copy_in(c, c->work.as<double>(), c->T.as<double>()));
    }
    if (noise != 0.0) TRY(k_add_noise(c, seed ^ 0x9E3779B97F4A7C15ull, noise));
    CUDA_TRY(cudaMemsetAsync((char *)c->T.p + (size_t)c->nstore * 8, 0, 256, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_get_tensor(itcpd_ctx *c, double *host) {
    CHECK_CTX(c);
    NEED_T(c);
    ARG_CHECK(c->has_tensor && host, "no tensor / null host pointer");
    USE_DEVICE(c);
    if (c->ld0 == c->dims[0]) {
        CUDA_TRY(cudaMemcpyAsync(host, c->T.p, (size_t)c->nelem * 8, cudaMemcpyDeviceToHost, c->stream));
    } else {
        TRY(c->work.reserve((size_t)c->nelem * 8));
        TRY(k_pad_copy_out(c, c->T.as<double>(), c->work.as<double>()));
        CUDA_TRY(cudaMemcpyAsync(host, c->work.p, (size_t)c->nelem * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_tensor_norm(itcpd_ctx *c, double *fro) {
    CHECK_CTX(c);
    NEED_T(c);
    ARG_CHECK(c->has_tensor && fro, "no tensor / null out pointer");
    USE_DEVICE(c);
    TRY(c->fit2.reserve(64));
    TRY(k_sumsq(c, c->T.as<double>(), c->nstore, c->fit2.as<double>()));
    if (comm_active(c)) TRY(comm_allreduce_sum(c, c->fit2.as<double>(), 1));
    CUDA_TRY(cudaMemcpyAsync(c->pinned, c->fit2.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *fro = std::sqrt(c->pinned[0]);
    return ITCPD_OK;
}

// ---- CPD state --------------------------------------------------------------------------------
int itcpd_set_rank(itcpd_ctx *c, int rank) {
    CHECK_CTX(c);
    ARG_CHECK(rank >= 1, "rank must be positive");
    USE_DEVICE(c);
    c->rank = rank;
    invalidate_all(c);
    if (c->has_tensor) TRY(ensure_cpd_buffers(c));
    return ITCPD_OK;
}

int itcpd_set_factor(itcpd_ctx *c, int mode, const double *host) {
    CHECK_CTX(c);
    CHECK_MODE(c, mode);
    ARG_CHECK(host && c->rank > 0, "null host pointer / rank not set");
    USE_DEVICE(c);
    TRY(ensure_cpd_buffers(c));
    CUDA_TRY(cudaMemcpyAsync(c->A[mode].p, host, (size_t)c->dims[mode] * c->rank * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->fver[mode]++;
    return ITCPD_OK;
}

static int d2h(itcpd_ctx *c, double *host, const void *dev, size_t bytes) {
    CUDA_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_get_factor(itcpd_ctx *c, int mode, double *host) {
    CHECK_CTX(c);
    CHECK_MODE(c, mode);
    ARG_CHECK(host && c->A[mode].p, "null host pointer / factor not set");
    USE_DEVICE(c);
    return d2h(c, host, c->A[mode].p, (size_t)c->dims[mode] * c->rank * 8);
}

int itcpd_set_lambda(itcpd_ctx *c, const double *host) {
    CHECK_CTX(c);
    ARG_CHECK(host && c->rank > 0, "null host pointer / rank not set");
    USE_DEVICE(c);
    TRY(ensure_cpd_buffers(c));
    CUDA_TRY(cudaMemcpyAsync(c->lambda.p, host, (size_t)c->rank * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_get_lambda(itcpd_ctx *c, double *host) {
    CHECK_CTX(c);
    ARG_CHECK(host && c->lambda.p, "null host pointer / lambda not set");
    USE_DEVICE(c);
    return d2h(c, host, c->lambda.p, (size_t)c->rank * 8);
}

int itcpd_get_gram(itcpd_ctx *c, int mode, double *host) {
    CHECK_CTX(c);
    CHECK_MODE(c, mode);
    ARG_CHECK(host && c->G[mode].p, "null host pointer / gram not computed");
    USE_DEVICE(c);
    return d2h(c, host, c->G[mode].p, (size_t)c->rank * c->rank * 8);
}

int itcpd_random_cpd(itcpd_ctx *c, uint64_t seed) {
    CHECK_CTX(c);
    USE_DEVICE(c);
    TRY(ensure_cpd_buffers(c));
    ARG_CHECK(!comm_active(c), "random_cpd is single-GPU; set the factor slabs explicitly when sharded");
    uint64_t off = 0;
    for (int n = 0; n < c->order; ++n) {
        const int64_t cnt = c->dims[n] * c->rank;
        TRY(k_randn_matrix(c, c->X.as<double>(), cnt, seed, off));
        off += (uint64_t)cnt + (cnt & 1);
        TRY(k_colnorm_scale(c, c->X.as<double>(), c->dims[n], c->rank, c->A[n].as<double>(), c->lambda.as<double>(), false));
        c->fver[n]++;
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

// ---- per-mode hooks ---------------------------------------------------------------------------
int itcpd_compute_grams(itcpd_ctx *c) {
    CHECK_CTX(c);
    USE_DEVICE(c);
    TRY(ensure_cpd_buffers(c));
    for (int n = 0; n < c->order; ++n) TRY(gram_device(c, n));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_gram_hadamard(itcpd_ctx *c, int mode, double *host_out) {
    CHECK_CTX(c);
    CHECK_MODE(c, mode);
    USE_DEVICE(c);
    TRY(ensure_cpd_buffers(c));
    TRY(k_gram_hadamard(c, mode, c->Gamma.as<double>()));
    if (host_out) return d2h(c, host_out, c->Gamma.p, (size_t)c->rank * c->rank * 8);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_mttkrp(itcpd_ctx *c, int mode, double *host_out) {
    CHECK_CTX(c);
    NEED_T(c);
    CHECK_MODE(c, mode);
    USE_DEVICE(c);
    TRY(ensure_cpd_buffers(c));
    TRY(mttkrp_device(c, mode));
    if (host_out) return d2h(c, host_out, c->M[mode].p, (size_t)c->dims[mode] * c->rank * 8);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_solve(itcpd_ctx *c, int mode, double chol_tol, int *path_out, int *rank_out) {
    CHECK_CTX(c);
    CHECK_MODE(c, mode);
    USE_DEVICE(c);
    ARG_CHECK(c->m_valid[mode], "call itcpd_mttkrp(mode) before itcpd_solve(mode)");
    TRY(k_solve(c, c->Gamma.as<double>(), c->M[mode].as<double>(), c->dims[mode], c->rank, chol_tol, c->X.as<double>(), c->status.as<int>()));
    int *h = reinterpret_cast<int *>(c->pinned);
    CUDA_TRY(cudaMemcpyAsync(h, c->status.p, 12, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (path_out) *path_out = h[0];
    if (rank_out) *rank_out = h[1];
    return ITCPD_OK;
}

int itcpd_last_solve_status(itcpd_ctx *c, int mode_slot, int *path_out, int *rank_out) {
    CHECK_CTX(c);
    ARG_CHECK(mode_slot >= 0 && mode_slot < ITCPD_MAX_ORDER && c->status.p, "bad slot / no solve has run (slot 0 outside whole sweeps, the mode index after a sweep)");
    USE_DEVICE(c);
    int *h = reinterpret_cast<int *>(c->pinned);
    CUDA_TRY(cudaMemcpyAsync(h, c->status.as<int>() + 3 * mode_slot, 12, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (path_out) *path_out = h[0];
    if (rank_out) *rank_out = h[1];
    return ITCPD_OK;
}

int itcpd_normalize(itcpd_ctx *c, int mode) {
    CHECK_CTX(c);
    CHECK_MODE(c, mode);
    USE_DEVICE(c);
    TRY(k_colnorm_scale(c, c->X.as<double>(), c->dims[mode], c->rank, c->A[mode].as<double>(), c->lambda.as<double>(), mode == c->order - 1));
    c->fver[mode]++;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_post_solve(itcpd_ctx *c, int mode) {
    CHECK_CTX(c);
    CHECK_MODE(c, mode);
    USE_DEVICE(c);
    TRY(gram_device(c, mode));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_fit_terms(itcpd_ctx *c, double *inner, double *model_norm2) {
    CHECK_CTX(c);
    USE_DEVICE(c);
    ARG_CHECK(c->m_valid[c->order - 1], "the last mode's MTTKRP has not been computed");
    TRY(k_fit_terms(c, c->fit2.as<double>(), true));
    CUDA_TRY(cudaMemcpyAsync(c->pinned, c->fit2.p, 16, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (inner) *inner = c->pinned[0];
    if (model_norm2) *model_norm2 = c->pinned[1];
    return ITCPD_OK;
}

// ---- CPDiffCheck / CPAngleCheck scalars (cp_diff_check.jl:20-71, cp_angle_check.jl:20-73) -----------------------
// PrevCP snapshot and the two scalars the host state machines need: <T_prev, T_curr> and ||T_curr||^2, from the factor
// matrices only (cp_cp_contract + norm_factors).  The sharded last factor's cross-Grams are all-reduced.
int itcpd_cpd_snapshot(itcpd_ctx *c) {
    CHECK_CTX(c);
    USE_DEVICE(c);
    TRY(ensure_cpd_buffers(c));
    for (int n = 0; n < c->order; ++n) {
        const size_t b = (size_t)c->dims[n] * c->rank * 8;
        TRY(c->prevA[n].reserve(b));
        CUDA_TRY(cudaMemcpyAsync(c->prevA[n].p, c->A[n].p, b, cudaMemcpyDeviceToDevice, c->stream));
    }
    TRY(c->prev_lambda.reserve((size_t)c->rank * 8));
    CUDA_TRY(cudaMemcpyAsync(c->prev_lambda.p, c->lambda.p, (size_t)c->rank * 8, cudaMemcpyDeviceToDevice, c->stream));
    c->has_snapshot = true;
    c->snapshot_rank = c->rank;
    return ITCPD_OK;
}

int itcpd_cpd_diff_terms(itcpd_ctx *c, double *inner_prev_curr, double *norm2_curr) {
    CHECK_CTX(c);
    USE_DEVICE(c);
    ARG_CHECK(c->has_snapshot && c->snapshot_rank == c->rank, "call itcpd_cpd_snapshot first (same rank)");
    TRY(k_cpd_diff_terms(c, c->fit2.as<double>()));
    CUDA_TRY(cudaMemcpyAsync(c->pinned, c->fit2.p, 16, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (inner_prev_curr) *inner_prev_curr = c->pinned[0];
    if (norm2_curr) *norm2_curr = c->pinned[1];
    return ITCPD_OK;
}

// ---- whole sweeps -----------------------------------------------------------------------------
// Per-sweep results (inner, model_norm2, #QRCP fallbacks) are appended to a device-side log by a one-thread
// kernel, so a sweep has no host-visible side effect until itcpd_sweep_results copies the log back.  That makes
// the sweep body a fixed kernel sequence: after one plain sweep (which sizes every scratch buffer) it is captured
// into a CUDA graph -- side-stream fork/join and NCCL calls included -- and the remaining sweeps are graph
// launches (no per-kernel launch gaps; the inner loop is launch-bound once the tensor is sharded 8 ways).
__global__ void log_sweep_kernel(const double *__restrict__ fit2, const int *__restrict__ status, int nmodes, double *__restrict__ log,
                                 unsigned long long *__restrict__ counter, unsigned long long cap) {
    const unsigned long long idx = *counter;
    if (idx < cap) {
        int fb = 0;
        for (int m = 0; m < nmodes; ++m) fb += (status[3 * m] == ITCPD_SOLVE_QRCP);
        log[idx] = fit2[0];              // <T, That> (slab-partial when sharded: reduced once in itcpd_sweep_results)
        log[cap + idx] = fit2[1];        // ||That||^2
        log[2 * cap + idx] = (double)fb;
    }
    *counter = idx + 1;
}

static int one_sweep_device(itcpd_ctx *c, double chol_tol) {
    const int N = c->order;
    const bool early_b = c->early_pass_b && c->mttkrp_alg != ITCPD_MTTKRP_DIRECT && c->split_b < c->split_a && !comm_active(c) && !c->time_phases;
    for (int mode = 0; mode < N; ++mode) {
        TRY(mode_update_device(c, mode, chol_tol, c->status.as<int>() + 3 * mode));
        if (early_b && mode == c->split_b - 1) {
            // every factor pass B contracts is final for this sweep: start it now on the GEMM stream; mode split_a joins
            cudaStream_t main_stream = c->stream;
            CUDA_TRY(cudaEventRecord(c->ev_gemm_fork, main_stream));
            CUDA_TRY(cudaStreamWaitEvent(c->gemm_stream, c->ev_gemm_fork, 0));
            c->stream = c->gemm_stream;
            const int st = ensure_partial(c, 1);
            c->stream = main_stream;
            TRY(st);
            CUDA_TRY(cudaEventRecord(c->ev_gemm_done, c->gemm_stream));
            c->gemm_join_pending = true;
        }
    }
    if (c->gemm_join_pending) {  // cannot happen (mode split_a always follows), but a captured fork must never be left open
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_gemm_done, 0));
        c->gemm_join_pending = false;
    }
    TRY(k_fit_terms(c, c->fit2.as<double>(), false));
    TRY(phase_mark(c, PH_FIT));
    log_sweep_kernel<<<1, 1, 0, c->stream>>>(c->fit2.as<double>(), c->status.as<int>(), N, c->sweep_log.as<double>() + 1,
                                            reinterpret_cast<unsigned long long *>(c->sweep_log.p), (unsigned long long)c->sweep_log_cap);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

static void graph_key(const itcpd_ctx *c, double tol, int64_t key[24]) {
    int k = 0;
    key[k++] = (int64_t)(intptr_t)c->T.p; key[k++] = (int64_t)(intptr_t)c->A[0].p; key[k++] = (int64_t)(intptr_t)c->sweep_log.p;
    key[k++] = c->order; key[k++] = c->rank; key[k++] = c->split_a; key[k++] = c->split_b; key[k++] = c->mttkrp_alg;
    key[k++] = c->swizzle; key[k++] = c->tile_warps; key[k++] = c->stream_k; key[k++] = c->tma3d; key[k++] = c->overlap_factor;
    key[k++] = (int64_t)(intptr_t)c->comm ^ ((int64_t)c->peer_graph << 1) ^ (int64_t)c->peer_on; key[k++] = c->graph_epoch;
    memcpy(&key[k++], &tol, 8);
    for (int n = 0; n < ITCPD_MAX_ORDER; ++n) key[k++] = n < c->order ? c->dims[n] : 0;
}

static void drop_graph(itcpd_ctx *c) {
    if (c->sweep_graph_exec) { cudaGraphExecDestroy(c->sweep_graph_exec); c->sweep_graph_exec = nullptr; }
}

int itcpd_sweep_async(itcpd_ctx *c, int nsweeps, double chol_tol) {
    CHECK_CTX(c);
    NEED_T(c);
    ARG_CHECK(nsweeps >= 1, "nsweeps must be positive");
    USE_DEVICE(c);
    TRY(ensure_cpd_buffers(c));
    const int N = c->order;
    TRY(c->status.reserve(256 + (size_t)N * 12));
    if ((int64_t)nsweeps > c->sweep_log_cap || !c->sweep_log.p) {
        c->sweep_log_cap = std::max<int64_t>(4096, 2 * (int64_t)nsweeps);
        TRY(c->sweep_log.reserve((size_t)(1 + 3 * c->sweep_log_cap) * 8));
    }
    CUDA_TRY(cudaMemsetAsync(c->sweep_log.p, 0, 8, c->stream));
    c->sweep_log_reduced = false;
    int done = 0;
    // NCCL collectives are not captured: a sharded sweep is launched kernel by kernel unless it is NCCL-free (peer_graph)
    const bool graph_ok = c->use_graph && !c->time_gemm && !c->time_phases && (!comm_active(c) || peer_graph_active(c));
    int64_t key[24];
    graph_key(c, chol_tol, key);
    const bool have_exec = c->sweep_graph_exec && memcmp(key, c->sweep_graph_key, sizeof(key)) == 0;
    // short calls (the per-iteration loop of the reference API): replay an existing graph, or capture once a plain sweep with
    // this very configuration has sized every buffer
    const bool warm = c->graph_single && c->plain_sweep_key_valid && memcmp(key, c->plain_sweep_key, sizeof(key)) == 0;   // never without the option
    const bool want_graph = graph_ok && (nsweeps >= 3 || (c->graph_single && (have_exec || warm)));
    if (want_graph) {
        if (!have_exec) {
            drop_graph(c);
            if (!warm) {
                TRY(one_sweep_device(c, chol_tol));  // plain sweep: sizes every buffer, sets function attributes, builds tables
                done = 1;
            }
            graph_key(c, chol_tol, key);          // buffers may have been (re)allocated by the plain sweep
            const int64_t l0 = c->launches;
            cudaGraph_t graph = nullptr;
            // The capture records whatever ensure_partial decides from the host state: a partial that happens to be current now
            // (e.g. after itcpd_mttkrp in the per-hook path) would leave its GEMM out of every replayed sweep.  Capture from the
            // state every replay starts in -- both partials stale -- and put the host state back if the capture fails.
            uint64_t fver0[ITCPD_MAX_ORDER];
            bool mvalid0[ITCPD_MAX_ORDER];
            memcpy(fver0, c->fver, sizeof(fver0));
            memcpy(mvalid0, c->m_valid, sizeof(mvalid0));
            const int last0 = c->last_mttkrp_mode;
            c->PA.valid = c->PB.valid = false;
            CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
            int st = one_sweep_device(c, chol_tol);
            cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            // nothing ran: the device still holds the pre-capture factors, so the captured sweep's bookkeeping is undone either way
            memcpy(c->fver, fver0, sizeof(fver0));
            memcpy(c->m_valid, mvalid0, sizeof(mvalid0));
            c->last_mttkrp_mode = last0;
            c->PA.valid = c->PB.valid = false;
            c->gemm_join_pending = false;
            if (st != ITCPD_OK || e != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                if (st != ITCPD_OK) return st;
                set_error("CUDA graph capture of the sweep failed: %s", cudaGetErrorString(e));
                return ITCPD_ERR_CUDA;
            }
            c->launches = l0;  // nothing ran during the capture
            c->sweep_graph_launches = 0;
            {
                size_t nn = 0;
                if (cudaGraphGetNodes(graph, nullptr, &nn) == cudaSuccess) c->sweep_graph_launches = (int64_t)nn;
            }
            e = cudaGraphInstantiate(&c->sweep_graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { c->sweep_graph_exec = nullptr; set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); return ITCPD_ERR_CUDA; }
            memcpy(c->sweep_graph_key, key, sizeof(key));
        }
        for (; done < nsweeps; ++done) {
            CUDA_TRY(cudaGraphLaunch(c->sweep_graph_exec, c->stream));
            c->launches += c->sweep_graph_launches;
            for (int n = 0; n < N; ++n) { c->fver[n]++; c->m_valid[n] = true; }   // what one_sweep_device does to the host state
            c->last_mttkrp_mode = N - 1;
        }
    }
    if (done < nsweeps) {
        for (; done < nsweeps; ++done) TRY(one_sweep_device(c, chol_tol));
        graph_key(c, chol_tol, c->plain_sweep_key);   // a plain sweep with this configuration has run: scratch is sized
        c->plain_sweep_key_valid = true;
    }
    return ITCPD_OK;
}

int itcpd_sweep_results(itcpd_ctx *c, int nsweeps, double *inner, double *model_norm2, int *qrcp_fallbacks) {
    CHECK_CTX(c);
    USE_DEVICE(c);
    ARG_CHECK(nsweeps >= 1 && (int64_t)nsweeps <= c->sweep_log_cap && c->sweep_log.p, "no results logged for that many sweeps");
    TRY(ensure_pinned(c, 3 * (size_t)nsweeps + 8));
    double *log = c->sweep_log.as<double>() + 1;
    const size_t cap = (size_t)c->sweep_log_cap, nb = (size_t)nsweeps * 8;
    if (comm_active(c) && !c->sweep_log_reduced) {  // the only per-sweep collective that can wait
        TRY(comm_allreduce_sum(c, log, nsweeps));
        c->sweep_log_reduced = true;
    }
    CUDA_TRY(cudaMemcpyAsync(c->pinned, log, nb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->pinned + nsweeps, log + cap, nb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->pinned + 2 * (size_t)nsweeps, log + 2 * cap, nb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    int fb = 0;
    for (int s = 0; s < nsweeps; ++s) {
        if (inner) inner[s] = c->pinned[s];
        if (model_norm2) model_norm2[s] = c->pinned[nsweeps + s];
        fb += (int)c->pinned[2 * (size_t)nsweeps + s];
    }
    if (qrcp_fallbacks) *qrcp_fallbacks = fb;
    return ITCPD_OK;
}

int itcpd_sweep(itcpd_ctx *c, int nsweeps, double chol_tol, double *inner, double *model_norm2) {
    TRY(itcpd_sweep_async(c, nsweeps, chol_tol));
    return itcpd_sweep_results(c, nsweeps, inner, model_norm2, nullptr);
}

int itcpd_als_from_host(itcpd_ctx *c, int order, const int64_t *dims, const double *host_T, int rank,
                        const double *const *fin, int nsweeps, double chol_tol, double *const *fout, double *lambda_out,
                        double *inner, double *model_norm2) {
    CHECK_CTX(c);
    ARG_CHECK(dims && host_T && fin && fout, "null argument");
    USE_DEVICE(c);
    TRY(set_shape(c, order, dims));
    if (c->rank != rank) { c->rank = rank; invalidate_all(c); }
    TRY(ensure_cpd_buffers(c));
    // uploads are enqueued back to back on the handle's stream; no host sync until the results are read
    if (c->ld0 == c->dims[0]) {
        TRY(upload_to_device(c, c->T.p, host_T, (size_t)c->nelem * 8));
    } else {
        TRY(c->work.reserve((size_t)c->nelem * 8));
        TRY(upload_to_device(c, c->work.p, host_T, (size_t)c->nelem * 8));
        TRY(k_pad_copy_in(c, c->work.as<double>(), c->T.as<double>()));
    }
    CUDA_TRY(cudaMemsetAsync((char *)c->T.p + (size_t)c->nstore * 8, 0, 256, c->stream));
    for (int n = 0; n < order; ++n) {
        CUDA_TRY(cudaMemcpyAsync(c->A[n].p, fin[n], (size_t)dims[n] * rank * 8, cudaMemcpyHostToDevice, c->stream));
        c->fver[n]++;
    }
    for (int n = 0; n < order; ++n) TRY(gram_device(c, n));
    TRY(itcpd_sweep_async(c, nsweeps, chol_tol));
    for (int n = 0; n < order; ++n)
        CUDA_TRY(cudaMemcpyAsync(fout[n], c->A[n].p, (size_t)dims[n] * rank * 8, cudaMemcpyDeviceToHost, c->stream));
    if (lambda_out) CUDA_TRY(cudaMemcpyAsync(lambda_out, c->lambda.p, (size_t)rank * 8, cudaMemcpyDeviceToHost, c->stream));
    return itcpd_sweep_results(c, nsweeps, inner, model_norm2, nullptr);
}

// ---- reconstruct / residual ---------------------------------------------------------------------
int itcpd_reconstruct(itcpd_ctx *c, double *host) {
    CHECK_CTX(c);
    ARG_CHECK(host && c->has_tensor && c->rank > 0, "null host pointer / no state");
    USE_DEVICE(c);
    TRY(c->work.reserve((size_t)c->nelem * 8));
    TRY(k_reconstruct(c, c->work.as<double>(), nullptr));
    return d2h(c, host, c->work.p, (size_t)c->nelem * 8);
}

int itcpd_residual_norm(itcpd_ctx *c, double *fro) {
    CHECK_CTX(c);
    NEED_T(c);
    ARG_CHECK(fro && c->has_tensor && c->rank > 0, "null out pointer / no state");
    USE_DEVICE(c);
    TRY(k_reconstruct(c, nullptr, c->fit2.as<double>()));
    if (comm_active(c)) TRY(comm_allreduce_sum(c, c->fit2.as<double>(), 1));
    CUDA_TRY(cudaMemcpyAsync(c->pinned, c->fit2.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *fro = std::sqrt(c->pinned[0]);
    return ITCPD_OK;
}

// ---- sampled path -------------------------------------------------------------------------------
static int ensure_leverage(itcpd_ctx *c, int mode) {
    if (c->lev_ver[mode] == c->fver[mode] && c->lev_ver[mode] != 0) return ITCPD_OK;
    TRY(k_gram(c, c->A[mode].as<double>(), c->dims[mode], c->rank, c->G[mode].as<double>()));
    TRY(k_leverage(c, c->A[mode].as<double>(), c->G[mode].as<double>(), c->dims[mode], c->rank, c->lev[mode].as<double>()));
    c->lev_ver[mode] = c->fver[mode];
    return ITCPD_OK;
}

int itcpd_leverage_scores(itcpd_ctx *c, int mode, double *host_out) {
    CHECK_CTX(c);
    CHECK_MODE(c, mode);
    USE_DEVICE(c);
    TRY(ensure_cpd_buffers(c));
    if (comm_active(c)) TRY(sharded_leverage(c, mode));  // the sharded factor returns the scores of the local rows
    else TRY(ensure_leverage(c, mode));
    if (host_out) return d2h(c, host_out, c->lev[mode].p, (size_t)c->dims[mode] * 8);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_sample_factor_matrices(itcpd_ctx *c, int skip_mode, int64_t nsamp, uint64_t seed, int64_t *host_out) {
    CHECK_CTX(c);
    CHECK_MODE(c, skip_mode);
    ARG_CHECK(nsamp >= 1 && host_out, "bad nsamp / null out pointer");
    USE_DEVICE(c);
    TRY(ensure_cpd_buffers(c));
    const size_t bytes = (size_t)nsamp * (c->order - 1) * 8;
    TRY(c->samp_piv.reserve(bytes));
    if (comm_active(c)) {
        TRY(sharded_sample_rows(c, skip_mode, nsamp, seed, c->samp_piv.as<int64_t>()));
    } else {
        for (int m = 0; m < c->order; ++m)
            if (m != skip_mode) TRY(ensure_leverage(c, m));
        TRY(k_sample_rows(c, skip_mode, nsamp, seed, c->samp_piv.as<int64_t>()));
    }
    CUDA_TRY(cudaMemcpyAsync(host_out, c->samp_piv.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

// the sampled least-squares problem of one mode (ProjectionAlgorithm.jl:57-68): K is nsamp x R, Ts is I x nsamp
// Ts == null (normal equations only): the fibres are read straight from the tensor through the device pivots `piv`
static int sampled_ls(itcpd_ctx *c, int mode, const double *K, const double *Ts, const int64_t *piv, int64_t nsamp, double chol_tol, int normal) {
    const int R = c->rank;
    const int64_t I = c->dims[mode];
    c->m_valid[mode] = false;  // M[mode] is reused for the *sampled* MTTKRP
    if (normal) {
        // (K'K) X' = (T_s K)'
        TRY(k_gram(c, K, nsamp, R, c->Gamma.as<double>()));
        TRY(k_sampled_mttkrp(c, mode, nsamp, piv, Ts, K, c->M[mode].as<double>()));
        return k_solve(c, c->Gamma.as<double>(), c->M[mode].as<double>(), I, R, chol_tol, c->X.as<double>(), c->status.as<int>());
    }
    ARG_CHECK(Ts != nullptr, "normal=false needs the gathered unfolding");
    ARG_CHECK(nsamp >= R && nsamp < ((int64_t)1 << 31), "normal=false needs at least R samples");
    return qrcp_ls_solve(c, K, (int)nsamp, R, Ts, I, c->X.as<double>(), c->status.as<int>(), 1);
}

static int check_pivots(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv) {
    int col = 0;
    for (int m = 0; m < c->order; ++m) {
        if (m == mode) continue;
        for (int64_t s = 0; s < nsamp; ++s) {
            const int64_t v = piv[s + nsamp * col];
            // pivots are global coordinates: on a sharded handle the last mode spans all the slabs
            const int64_t ext = (comm_active(c) && m == c->order - 1) ? sharded_last_rows(c) : c->dims[m];
            if (v < 1 || v > ext) { set_error("pivot (%lld,%d) = %lld out of range [1,%lld]", (long long)s, col, (long long)v, (long long)ext); return ITCPD_ERR_ARG; }
        }
        ++col;
    }
    return ITCPD_OK;
}

static int upload_pivots(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *host_pivots) {
    TRY(check_pivots(c, mode, nsamp, host_pivots));
    const size_t bytes = (size_t)nsamp * (c->order - 1) * 8;
    TRY(c->samp_piv.reserve(bytes));
    CUDA_TRY(cudaMemcpyAsync(c->samp_piv.p, host_pivots, bytes, cudaMemcpyHostToDevice, c->stream));
    return ITCPD_OK;
}

int itcpd_pivot_hadamard(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *host_pivots, double *host_out) {
    CHECK_CTX(c);
    CHECK_MODE(c, mode);
    ARG_CHECK(nsamp >= 1 && host_pivots && host_out, "bad nsamp / null pointer");
    ARG_CHECK(!comm_active(c), "itcpd_pivot_hadamard is a single-GPU diagnostic; sharded handles go through itcpd_sampled_update");
    USE_DEVICE(c);
    TRY(ensure_cpd_buffers(c));
    TRY(upload_pivots(c, mode, nsamp, host_pivots));
    TRY(c->samp_K.reserve((size_t)nsamp * c->rank * 8));
    TRY(k_pivot_hadamard(c, mode, nsamp, c->samp_piv.as<int64_t>(), c->samp_K.as<double>()));
    return d2h(c, host_out, c->samp_K.p, (size_t)nsamp * c->rank * 8);
}

int itcpd_gather_fibers(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *host_pivots, double *host_out) {
    CHECK_CTX(c);
    NEED_T(c);
    CHECK_MODE(c, mode);
    ARG_CHECK(nsamp >= 1 && host_pivots && host_out && c->has_tensor, "bad nsamp / null pointer / no tensor");
    USE_DEVICE(c);
    TRY(upload_pivots(c, mode, nsamp, host_pivots));
    TRY(c->samp_T.reserve((size_t)nsamp * c->dims[mode] * 8));
    if (comm_active(c)) {  // the rank's piece: owned fibres (zeros elsewhere), or its rows of the fibres along the sharded mode
        TRY(sharded_gather_fibers(c, mode, nsamp, c->samp_piv.as<int64_t>(), c->samp_T.as<double>()));
        return d2h(c, host_out, c->samp_T.p, (size_t)nsamp * c->dims[mode] * 8);
    }
    TRY(k_gather_fibers(c, mode, nsamp, c->samp_piv.as<int64_t>(), c->samp_T.as<double>()));
    return d2h(c, host_out, c->samp_T.p, (size_t)nsamp * c->dims[mode] * 8);
}

int itcpd_column_to_multi_coords(int64_t ncols, const int64_t *cols, int ndims, const int64_t *dims, int64_t *out) {
    ARG_CHECK(cols && dims && out && ndims >= 1, "null argument");
    for (int64_t s = 0; s < ncols; ++s) {
        int64_t rem = cols[s] - 1;
        for (int d = 0; d < ndims; ++d) {
            out[s + ncols * d] = rem % dims[d] + 1;
            rem /= dims[d];
        }
    }
    return ITCPD_OK;
}

int itcpd_multi_coords_to_column(int64_t ncols, const int64_t *coords, int ndims, const int64_t *dims, int64_t *out) {
    ARG_CHECK(coords && dims && out && ndims >= 1, "null argument");
    for (int64_t s = 0; s < ncols; ++s) {
        int64_t col = 0, stride = 1;
        for (int d = 0; d < ndims; ++d) {
            col += (coords[s + ncols * d] - 1) * stride;
            stride *= dims[d];
        }
        out[s] = col + 1;
    }
    return ITCPD_OK;
}

// CSR by sketch row with a stable counting sort: entries of a row stay in increasing non-zero order
// (the order the reference's dict_rows visits them, pivot_mapping.jl:127-137); uploaded into c->work.
struct SketchCsr {
    std::vector<int64_t> row_ptr;
    std::vector<int64_t> col_own;   // pageable storage of the single-shot entry points; the SE-QRCS set-up points col / val at the
    std::vector<double> val_own;    // handle's pinned slots instead (c->sketch_pin)
    int64_t *col = nullptr;
    double *val = nullptr;
    int64_t nnz = 0;
    const int64_t *d_ptr = nullptr;
    int64_t *d_col = nullptr;       // the sketch kernel turns the column numbers into element offsets in place
    const double *d_val = nullptr;
};

// Pure host work (no CUDA calls: it also runs on the set-up's helper thread).  Stable and parallel: the non-zeros are cut into one
// chunk per thread, every thread counts its chunk per sketch row, a serial pass over (row, thread) hands each thread its first slot
// in every row, and the threads place their chunks -- thread w's entries of a row land behind those of threads < w, in order.
static int sketch_csr_fill(int l, int s_eff, int64_t ncols, const int *rows0, const double *vals, SketchCsr &k) {
    const int64_t nnz = ncols * s_eff;
    k.nnz = nnz;
    k.row_ptr.assign((size_t)l + 1, 0);
    const int hw = (int)std::thread::hardware_concurrency();
    int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(8, hw > 0 ? hw : 1), nnz / 262144));
    if (const char *f = getenv("ITCPD_SKETCH_THREADS")) nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(64, atoi(f)), ncols));   // tests: threads on small inputs
    std::vector<int64_t> first((size_t)nt * l, 0);          // counts, then first slots: first[w * l + row]
    std::atomic<int> bad(0);
    auto chunk_lo = [&](int w) { return (nnz * w / nt) / s_eff * s_eff; };   // chunk borders on column borders
    auto count = [&](int w) {
        int64_t *cnt = first.data() + (size_t)w * l;
        const int64_t hi = w + 1 == nt ? nnz : chunk_lo(w + 1);
        for (int64_t q = chunk_lo(w); q < hi; ++q) {
            const int r = rows0[q];
            if (r < 0 || r >= l) { bad = 1; return; }
            cnt[r]++;
        }
    };
    auto place = [&](int w) {
        int64_t *slot = first.data() + (size_t)w * l;
        const int64_t lo = chunk_lo(w), hi = w + 1 == nt ? nnz : chunk_lo(w + 1);
        int64_t column = lo / s_eff;
        int within = 0;
        for (int64_t q = lo; q < hi; ++q) {
            const int64_t pos = slot[rows0[q]]++;
            k.col[pos] = column;
            k.val[pos] = vals[q];
            if (++within == s_eff) { within = 0; ++column; }
        }
    };
    auto run = [&](auto &&fn) {
        std::vector<std::thread> th;
        for (int w = 1; w < nt; ++w) th.emplace_back(fn, w);
        fn(0);
        for (auto &t : th) t.join();
    };
    run(count);
    if (bad) return ITCPD_ERR_ARG;
    int64_t running = 0;
    for (int r = 0; r < l; ++r) {
        for (int w = 0; w < nt; ++w) {
            const int64_t n_here = first[(size_t)w * l + r];
            first[(size_t)w * l + r] = running;
            running += n_here;
        }
        k.row_ptr[(size_t)r + 1] = running;
    }
    run(place);
    return ITCPD_OK;
}

static int sketch_csr_upload(itcpd_ctx *c, int l, SketchCsr &k) {
    const size_t b_ptr = ((size_t)l + 1) * 8, b_col = (size_t)k.nnz * 8;
    TRY(c->work.reserve(b_ptr + 2 * b_col + 64));
    char *base = (char *)c->work.p;
    CUDA_TRY(cudaMemcpyAsync(base, k.row_ptr.data(), b_ptr, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + b_ptr, k.col, b_col, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + b_ptr + b_col, k.val, b_col, cudaMemcpyHostToDevice, c->stream));
    k.d_ptr = (const int64_t *)base;
    k.d_col = (int64_t *)(base + b_ptr);
    k.d_val = (const double *)(base + b_ptr + b_col);
    return ITCPD_OK;
}

static int build_sketch_csr(itcpd_ctx *c, int l, int s_eff, int64_t ncols, const int *rows0, const double *vals, SketchCsr &k) {
    k.col_own.resize((size_t)(ncols * s_eff));
    k.val_own.resize((size_t)(ncols * s_eff));
    k.col = k.col_own.data();
    k.val = k.val_own.data();
    if (sketch_csr_fill(l, s_eff, ncols, rows0, vals, k) != ITCPD_OK) {
        set_error("sketch row index out of range");
        return ITCPD_ERR_ARG;
    }
    return sketch_csr_upload(c, l, k);
}

// every mode but the first reads its unfolding's columns with a stride: make the unfolding explicit once (one read + one write of
// the tensor) when HBM has room for a second copy, and sketch from contiguous columns.  *out stays nullptr when it does not apply.
static int unfold_for_sketch(itcpd_ctx *c, int mode, const double **out) {
    *out = nullptr;
    if (mode == 0 || !c->sketch_unfold) return ITCPD_OK;
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const size_t need = (size_t)c->nelem * 8;
    if (c->unfolded.bytes < need && free_b < need + ((size_t)4 << 30)) return ITCPD_OK;
    TRY(c->unfolded.reserve(need));
    TRY(k_unfold(c, mode, c->unfolded.as<double>()));
    *out = c->unfolded.as<double>();
    return ITCPD_OK;
}

int itcpd_sketch_unfolding(itcpd_ctx *c, int mode, int l, int s, const int *rows0, const double *vals, double *host_out) {
    CHECK_CTX(c);
    NEED_T(c);
    CHECK_MODE(c, mode);
    ARG_CHECK(l >= 1 && s >= 1 && rows0 && vals && host_out && c->has_tensor, "bad argument");
    ARG_CHECK(!comm_active(c), "itcpd_sketch_unfolding needs the whole unfolding: run the pivot setup on an unsharded handle (replicas), then shard");
    USE_DEVICE(c);
    const int s_eff = std::min(s, l);
    const int64_t ncols = c->nelem / c->dims[mode];
    SketchCsr k;
    TRY(build_sketch_csr(c, l, s_eff, ncols, rows0, vals, k));
    TRY(c->samp_T.reserve((size_t)l * c->dims[mode] * 8));
    const double *unfolded = nullptr;
    if (c->sketch_unfold == 2) TRY(unfold_for_sketch(c, mode, &unfolded));   // the set-up's path, reachable for the parity tests
    TRY(k_sketch_csr(c, mode, l, k.nnz, k.d_ptr, k.d_col, k.d_val, c->samp_T.as<double>(), unfolded));
    const int rc = d2h(c, host_out, c->samp_T.p, (size_t)l * c->dims[mode] * 8);
    c->unfolded.release();
    return rc;
}

// sparse-matrix variant of the sketch (pivot_mapping.jl:90-104): Omega (l x ncols) in compressed-sparse-column form exactly as Julia's
// SparseMatrixCSC stores it (colptr: ncols + 1 entries, rowval, both 1-based; nzval).  A_sk = T_(mode) Omega^T, entries of a sketch row
// visited in increasing column order (the order of `omega[j, :].nzind`), so the result is bitwise the matrix-free variant's when Omega
// came from sparse_sign_matrix(...; omega = true).
int itcpd_sketch_unfolding_csc(itcpd_ctx *c, int mode, int l, int64_t ncols, const int64_t *colptr, const int64_t *rowval, const double *nzval,
                               double *host_out) {
    CHECK_CTX(c);
    NEED_T(c);
    CHECK_MODE(c, mode);
    ARG_CHECK(l >= 1 && colptr && rowval && nzval && host_out, "bad argument");
    ARG_CHECK(!comm_active(c), "itcpd_sketch_unfolding_csc needs the whole unfolding: run the pivot setup on an unsharded handle (replicas), then shard");
    ARG_CHECK(ncols == c->nelem / c->dims[mode], "Omega must have one column per column of the mode's unfolding");
    USE_DEVICE(c);
    const int64_t nnz = colptr[ncols] - 1;
    ARG_CHECK(colptr[0] == 1 && nnz >= 0, "colptr must be 1-based (SparseMatrixCSC)");
    for (int64_t col = 0; col < ncols; ++col) ARG_CHECK(colptr[col + 1] >= colptr[col], "colptr must be non-decreasing");
    SketchCsr k;
    k.nnz = nnz;
    k.row_ptr.assign((size_t)l + 1, 0);
    k.col_own.resize((size_t)nnz);
    k.val_own.resize((size_t)nnz);
    k.col = k.col_own.data();
    k.val = k.val_own.data();
    for (int64_t q = 0; q < nnz; ++q) {
        ARG_CHECK(rowval[q] >= 1 && rowval[q] <= l, "sketch row index out of range");
        k.row_ptr[(size_t)rowval[q]]++;
    }
    for (int j = 0; j < l; ++j) k.row_ptr[(size_t)j + 1] += k.row_ptr[j];
    {
        std::vector<int64_t> fill(k.row_ptr.begin(), k.row_ptr.end() - 1);
        for (int64_t col = 0; col < ncols; ++col) {   // increasing column order inside every sketch row
            for (int64_t q = colptr[col] - 1; q < colptr[col + 1] - 1; ++q) {
                const int64_t pos = fill[(size_t)rowval[q] - 1]++;
                k.col[pos] = col;
                k.val[pos] = nzval[q];
            }
        }
    }
    TRY(sketch_csr_upload(c, l, k));
    TRY(c->samp_T.reserve((size_t)l * c->dims[mode] * 8));
    const double *unfolded = nullptr;
    if (c->sketch_unfold == 2) TRY(unfold_for_sketch(c, mode, &unfolded));   // the set-up's path, reachable for the parity tests
    TRY(k_sketch_csr(c, mode, l, k.nnz, k.d_ptr, k.d_col, k.d_val, c->samp_T.as<double>(), unfolded));
    const int rc = d2h(c, host_out, c->samp_T.p, (size_t)l * c->dims[mode] * 8);
    c->unfolded.release();
    return rc;
}

// ---- column-pivoted QR on the device (pivot-projected setup) ----------------------------------------
static int qrcp_fetch(itcpd_ctx *c, int64_t n, int64_t nr, int64_t *piv_out, double *rdiag_out) {
    std::vector<int64_t> tmp((size_t)n);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), c->qr_piv.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (rdiag_out) CUDA_TRY(cudaMemcpyAsync(rdiag_out, c->qr_rdiag.p, (size_t)nr * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < n; ++i) piv_out[i] = tmp[(size_t)i] + 1;  // 1-based, like LinearAlgebra's QRPivoted.p
    return ITCPD_OK;
}

int itcpd_qrcp_matrix(itcpd_ctx *c, int64_t m, int64_t n, const double *host_A, int64_t steps, int64_t *piv_out, double *rdiag_out) {
    CHECK_CTX(c);
    ARG_CHECK(m >= 1 && n >= 1 && host_A && piv_out && steps >= 1, "bad argument");
    USE_DEVICE(c);
    const int64_t nr = std::min<int64_t>(std::min(m, n), steps);
    TRY(c->qr_A.reserve((size_t)m * n * 8));
    TRY(c->qr_piv.reserve((size_t)n * 8));
    TRY(c->qr_rdiag.reserve((size_t)std::min(m, n) * 8));
    CUDA_TRY(cudaMemcpyAsync(c->qr_A.p, host_A, (size_t)m * n * 8, cudaMemcpyHostToDevice, c->stream));
    TRY(k_qrcp_wide(c, c->qr_A.as<double>(), m, n, nr, c->qr_piv.as<int64_t>(), c->qr_rdiag.as<double>()));
    return qrcp_fetch(c, n, nr, piv_out, rdiag_out);
}

int itcpd_qrcp_unfolding(itcpd_ctx *c, int mode, int64_t *piv_out, double *rdiag_out) {
    CHECK_CTX(c);
    NEED_T(c);
    CHECK_MODE(c, mode);
    ARG_CHECK(piv_out && c->has_tensor, "null out pointer / no tensor");
    ARG_CHECK(!comm_active(c), "itcpd_qrcp_unfolding needs the whole unfolding: run the pivot setup on an unsharded handle (replicas), then shard");
    USE_DEVICE(c);
    const int64_t m = c->dims[mode], n = c->nelem / m, nr = std::min(m, n);
    TRY(c->qr_A.reserve((size_t)m * n * 8));
    TRY(c->qr_piv.reserve((size_t)n * 8));
    TRY(c->qr_rdiag.reserve((size_t)nr * 8));
    TRY(k_unfold(c, mode, c->qr_A.as<double>()));
    TRY(k_qrcp_wide(c, c->qr_A.as<double>(), m, n, nr, c->qr_piv.as<int64_t>(), c->qr_rdiag.as<double>()));
    return qrcp_fetch(c, n, nr, piv_out, rdiag_out);
}

// One SE-QRCS embedding on the host: the reference's generator (same libc rand() stream, SEQRCS.jl:144-146) and the sketch in
// CSR-by-sketch-row order.  No CUDA calls: the set-up runs it for mode n+1 on a helper thread while the device factorises mode n.
struct Embedding {
    std::vector<double> vals;
    std::vector<int> rows, colstarts;
    SketchCsr k;
    int l = 0, s_eff = 0;
    int64_t ncols = 0;
    int status = ITCPD_OK;
    double gen_ms = 0.0, csr_ms = 0.0;
};

static void host_embedding(int l, int s, int64_t ncols, int injective, int64_t seed, Embedding &e) {
    const auto t0 = std::chrono::steady_clock::now();
    e.l = l;
    e.s_eff = std::min(s, l);
    e.ncols = ncols;
    const size_t nnz = (size_t)ncols * e.s_eff;
    if (e.vals.size() < nnz) { e.vals.resize(nnz); e.rows.resize(nnz); }
    if (e.colstarts.size() < (size_t)ncols + 1) e.colstarts.resize((size_t)ncols + 1);
    if (seed >= 0) srand((unsigned)seed);
    if (injective) itcpd_sparsestack(l, (int)ncols, s, e.vals.data(), e.rows.data(), e.colstarts.data());
    else itcpd_sparse_sign(l, (int)ncols, s, e.vals.data(), e.rows.data(), e.colstarts.data());
    const auto t1 = std::chrono::steady_clock::now();
    e.status = sketch_csr_fill(l, e.s_eff, ncols, e.rows.data(), e.vals.data(), e.k);
    const auto t2 = std::chrono::steady_clock::now();
    e.gen_ms = 1e3 * std::chrono::duration<double>(t1 - t0).count();
    e.csr_ms = 1e3 * std::chrono::duration<double>(t2 - t1).count();
}

// pinned home of an embedding's CSR entries (slot 0 / 1 alternate between consecutive modes); grows, never shrinks
static int sketch_pin_reserve(itcpd_ctx *c, int slot, int64_t nnz, SketchCsr &k) {
    const size_t need = (size_t)nnz * 16;
    if (c->sketch_pin_bytes[slot] < need) {
        if (c->sketch_pin[slot]) cudaFreeHost(c->sketch_pin[slot]);
        c->sketch_pin[slot] = nullptr;
        c->sketch_pin_bytes[slot] = 0;
        CUDA_TRY(cudaHostAlloc(&c->sketch_pin[slot], need, cudaHostAllocDefault));
        c->sketch_pin_bytes[slot] = need;
    }
    k.col = (int64_t *)c->sketch_pin[slot];
    k.val = (double *)((char *)c->sketch_pin[slot] + (size_t)nnz * 8);
    return ITCPD_OK;
}

// ITCPD_TRACE_SETUP=1: host wall-clock marks (after a stream synchronize) between the steps of the SE-QRCS set-up, to stderr
struct SetupTrace {
    bool on;
    cudaStream_t st;
    std::chrono::steady_clock::time_point t0;
    explicit SetupTrace(cudaStream_t s) : on(getenv("ITCPD_TRACE_SETUP") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[itcpd setup] %-28s %8.2f ms\n", what, 1e3 * std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    }
};

// device half of SE-QRCS for one mode, given its embedding (SEQRCS.jl:149-168)
static int seqrcs_device_part(itcpd_ctx *c, int mode, int t, Embedding &e, int64_t *piv_out, double *rdiag_out, int64_t *nrdiag_out,
                              int64_t *ncand_out) {
    const int N = c->order, l = e.l;
    const int64_t m = c->dims[mode], n = e.ncols;
    SketchCsr &k = e.k;
    SetupTrace tr(c->stream);
    // 2. sketch A_sk = T_(mode) Omega^T (I x l) (SEQRCS.jl:149)
    TRY(sketch_csr_upload(c, l, k));
    TRY(c->qr_A.reserve((size_t)m * std::max<int64_t>(l, 1) * 8));
    tr.mark("csr upload");
    const double *unfolded = nullptr;
    if (mode > 0 && c->sketch_unfold && tr.on) {   // trace only: time the allocation apart from the transpose
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        tr.mark("cudaMemGetInfo");
        if (free_b > (size_t)c->nelem * 8 + ((size_t)4 << 30)) { TRY(c->unfolded.reserve((size_t)c->nelem * 8)); tr.mark("allocate the unfolding"); }
    }
    TRY(unfold_for_sketch(c, mode, &unfolded));
    if (unfolded) tr.mark("explicit unfolding");
    TRY(k_sketch_csr(c, mode, l, k.nnz, k.d_ptr, k.d_col, k.d_val, c->qr_A.as<double>(), unfolded));
    tr.mark("sketch kernel");
    // 3. QRCP of the sketch, first t pivots (SEQRCS.jl:152-158)
    TRY(c->qr_piv.reserve((size_t)std::max<int64_t>(n, l) * 8));
    TRY(c->qr_rdiag.reserve((size_t)std::max<int64_t>(m, 1) * 8 + (size_t)std::min<int64_t>(m, l) * 8));
    TRY(k_qrcp_wide(c, c->qr_A.as<double>(), m, l, std::min<int64_t>(t, std::min<int64_t>(m, l)), c->qr_piv.as<int64_t>(), c->qr_rdiag.as<double>()));
    std::vector<int64_t> p_sk((size_t)t);
    CUDA_TRY(cudaMemcpyAsync(p_sk.data(), c->qr_piv.p, (size_t)t * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    tr.mark("qrcp of the sketch");
    // 4. candidate columns: for each selected sketch row (pivot order) the columns hashed to it, unique'd (SEQRCS.jl:159)
    std::vector<char> seen((size_t)n, 0);
    std::vector<int64_t> cand;
    for (int64_t q = 0; q < t; ++q) {
        const int64_t r = p_sk[(size_t)q];
        for (int64_t en = k.row_ptr[(size_t)r]; en < k.row_ptr[(size_t)r + 1]; ++en) {
            const int64_t col = k.col[en];
            if (!seen[(size_t)col]) { seen[(size_t)col] = 1; cand.push_back(col); }
        }
    }
    // SEQRCS(Val(true), ...) (SEQRCS.jl:109-113) finds the same columns with findall over eachcol(omega[p_sk, :]): increasing column order
    if (c->seqrcs_use_omega) std::sort(cand.begin(), cand.end());
    const int64_t nc = (int64_t)cand.size();
    ARG_CHECK(nc >= 1, "SE-QRCS selected no candidate column");
    // 5. gather the candidate columns (fused_flatten_sample, SEQRCS.jl:166) and QRCP them
    std::vector<int64_t> coords((size_t)nc * (N - 1));
    {
        int col = 0;
        std::vector<int64_t> rem(cand);
        for (int q = 0; q < N; ++q) {
            if (q == mode) continue;
            for (int64_t i = 0; i < nc; ++i) { coords[(size_t)(i + nc * col)] = rem[(size_t)i] % c->dims[q] + 1; rem[(size_t)i] /= c->dims[q]; }
            ++col;
        }
    }
    TRY(c->samp_piv.reserve((size_t)nc * (N - 1) * 8));
    CUDA_TRY(cudaMemcpyAsync(c->samp_piv.p, coords.data(), (size_t)nc * (N - 1) * 8, cudaMemcpyHostToDevice, c->stream));
    tr.mark("candidates + coords");
    TRY(c->qr_A.reserve((size_t)m * nc * 8));
    TRY(k_gather_fibers(c, mode, nc, c->samp_piv.as<int64_t>(), c->qr_A.as<double>()));
    tr.mark("gather candidates");
    const int64_t nr = std::min(m, nc);
    TRY(k_qrcp_wide(c, c->qr_A.as<double>(), m, nc, nr, c->qr_piv.as<int64_t>(), c->qr_rdiag.as<double>()));
    std::vector<int64_t> p_sub((size_t)nc);
    CUDA_TRY(cudaMemcpyAsync(p_sub.data(), c->qr_piv.p, (size_t)nc * 8, cudaMemcpyDeviceToHost, c->stream));
    if (rdiag_out) CUDA_TRY(cudaMemcpyAsync(rdiag_out, c->qr_rdiag.p, (size_t)nr * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    tr.mark("qrcp of the candidates");
    // 6. p = [candidates[p_subset]; setdiff(1:n, candidates)] (SEQRCS.jl:167-168), 1-based
    int64_t w = 0;
    for (int64_t i = 0; i < nc; ++i) piv_out[w++] = cand[(size_t)p_sub[(size_t)i]] + 1;
    for (int64_t col = 0; col < n; ++col)
        if (!seen[(size_t)col]) piv_out[w++] = col + 1;
    tr.mark("pivot list");
    if (nrdiag_out) *nrdiag_out = nr;
    if (ncand_out) *ncand_out = nc;
    return ITCPD_OK;
}

// SE-QRCS of several modes in one call (the loop of optimizers/.../qr_lev_score_sampled.jl:80-176 over its random modes).  The
// embeddings are generated in the order of `modes`, so the libc rand() stream is consumed exactly as by the per-mode calls; the
// host half of mode i+1 (generator + CSR) runs on a helper thread while the device works on mode i.
int itcpd_seqrcs_modes(itcpd_ctx *c, int nmodes, const int *modes, const int *l, const int *s, const int *t, int injective,
                       const int64_t *seeds, int64_t *const *piv_out, double *const *rdiag_out, int64_t *nrdiag_out, int64_t *ncand_out) {
    CHECK_CTX(c);
    NEED_T(c);
    ARG_CHECK(nmodes >= 1 && modes && l && s && t && piv_out && c->has_tensor, "bad argument");
    ARG_CHECK(!comm_active(c), "itcpd_seqrcs needs the whole unfolding: run the pivot setup on an unsharded handle (replicas), then shard");
    for (int i = 0; i < nmodes; ++i) {
        CHECK_MODE(c, modes[i]);
        ARG_CHECK(l[i] >= 1 && s[i] >= 1 && t[i] >= 1 && t[i] <= l[i] && piv_out[i], "bad argument (need 1 <= t <= l)");
        const int64_t n = c->nelem / c->dims[modes[i]];
        ARG_CHECK(n < (int64_t)1 << 31 && (int64_t)n * std::min(s[i], l[i]) < (int64_t)1 << 31, "the reference's C generators index with 32-bit ints");
    }
    USE_DEVICE(c);
    Embedding emb[2];
    auto ncols_of = [&](int i) { return c->nelem / c->dims[modes[i]]; };
    auto start_host = [&](int i, std::thread &th) -> int {
        Embedding &e = emb[i & 1];
        TRY(sketch_pin_reserve(c, i & 1, ncols_of(i) * std::min(s[i], l[i]), e.k));   // pinned slots are allocated on the calling thread
        th = std::thread(host_embedding, l[i], s[i], ncols_of(i), injective, seeds ? seeds[i] : (int64_t)-1, std::ref(e));
        return ITCPD_OK;
    };
    std::thread th;
    TRY(start_host(0, th));
    int rc = ITCPD_OK;
    for (int i = 0; i < nmodes && rc == ITCPD_OK; ++i) {
        th.join();
        Embedding &e = emb[i & 1];
        if (getenv("ITCPD_TRACE_SETUP")) fprintf(stderr, "[itcpd setup] mode %d: generator %.2f ms, csr %.2f ms (host%s)\n", modes[i], e.gen_ms, e.csr_ms, i ? ", overlapped" : "");
        if (e.status != ITCPD_OK) { set_error("sketch row index out of range"); rc = e.status; break; }
        if (i + 1 < nmodes) rc = start_host(i + 1, th);
        if (rc == ITCPD_OK)
            rc = seqrcs_device_part(c, modes[i], t[i], e, piv_out[i], rdiag_out ? rdiag_out[i] : nullptr, nrdiag_out ? nrdiag_out + i : nullptr,
                                    ncand_out ? ncand_out + i : nullptr);
    }
    if (th.joinable()) th.join();
    {
        SetupTrace tr(c->stream);
        c->unfolded.release();
        tr.mark("release the unfolding");
    }
    return rc;
}

int itcpd_seqrcs(itcpd_ctx *c, int mode, int l, int s, int t, int injective, int64_t *piv_out, double *rdiag_out,
                 int64_t *nrdiag_out, int64_t *ncand_out) {
    int64_t *pivs[1] = {piv_out};
    double *rds[1] = {rdiag_out};
    return itcpd_seqrcs_modes(c, 1, &mode, &l, &s, &t, injective, nullptr, pivs, rds, nrdiag_out, ncand_out);
}

int itcpd_seqrcs_krp(itcpd_ctx *c, int mode, int l, int s, int t, int injective, int64_t *piv_out, double *rdiag_out,
                     int64_t *nrdiag_out, int64_t *ncand_out) {
    CHECK_CTX(c);
    CHECK_MODE(c, mode);
    ARG_CHECK(l >= 1 && s >= 1 && t >= 1 && t <= l && piv_out && c->has_tensor && c->rank > 0 && c->A[0].p, "bad argument (need 1 <= t <= l and a CPD state)");
    ARG_CHECK(!comm_active(c), "itcpd_seqrcs_krp needs the whole unfolding: run the pivot setup on an unsharded handle (replicas), then shard");
    USE_DEVICE(c);
    const int N = c->order, R = c->rank;
    int64_t n = 1;
    for (int q = 0; q < N; ++q) if (q != mode) n *= c->dims[q];
    ARG_CHECK(n < (int64_t)1 << 31 && (int64_t)n * std::min(s, l) < (int64_t)1 << 31, "the reference's C generators index with 32-bit ints");
    const int s_eff = std::min(s, l);
    std::vector<double> vals((size_t)n * s_eff);
    std::vector<int> rows((size_t)n * s_eff), colstarts((size_t)n + 1);
    if (injective) itcpd_sparsestack(l, (int)n, s, vals.data(), rows.data(), colstarts.data());
    else itcpd_sparse_sign(l, (int)n, s, vals.data(), rows.data(), colstarts.data());
    // sketch of the KRP (SEQRCS.jl:195): A_sk is l x R in the reference; we hold its transpose R x l, which is what is QR'd (:198)
    SketchCsr k;
    TRY(build_sketch_csr(c, l, s_eff, n, rows.data(), vals.data(), k));
    TRY(c->qr_A.reserve((size_t)R * l * 8));
    TRY(k_omega_hadamard(c, mode, l, k.d_ptr, k.d_col, k.d_val, c->qr_A.as<double>()));
    TRY(c->qr_piv.reserve((size_t)std::max<int64_t>(n, l) * 8));
    TRY(c->qr_rdiag.reserve((size_t)std::max<int64_t>(R, 1) * 8 * 2));
    TRY(k_qrcp_wide(c, c->qr_A.as<double>(), R, l, std::min<int64_t>(t, std::min<int64_t>(R, l)), c->qr_piv.as<int64_t>(), c->qr_rdiag.as<double>()));
    std::vector<int64_t> p_sk((size_t)t);
    CUDA_TRY(cudaMemcpyAsync(p_sk.data(), c->qr_piv.p, (size_t)t * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    // candidates: every column with a non-zero in one of the selected sketch rows, in ASCENDING column order
    // (rows_sel = omega[p_sk, :]; findall over eachcol, SEQRCS.jl:205-207)
    std::vector<char> seen((size_t)n, 0);
    for (int64_t q = 0; q < t; ++q) {
        const int64_t r = p_sk[(size_t)q];
        for (int64_t e = k.row_ptr[(size_t)r]; e < k.row_ptr[(size_t)r + 1]; ++e) seen[(size_t)k.col[e]] = 1;
    }
    std::vector<int64_t> cand;
    for (int64_t col = 0; col < n; ++col) if (seen[(size_t)col]) cand.push_back(col);
    const int64_t nc = (int64_t)cand.size();
    ARG_CHECK(nc >= 1, "SE-QRCS selected no candidate column");
    std::vector<int64_t> coords((size_t)nc * (N - 1));
    {
        int col = 0;
        std::vector<int64_t> rem(cand);
        for (int q = 0; q < N; ++q) {
            if (q == mode) continue;
            for (int64_t i = 0; i < nc; ++i) { coords[(size_t)(i + nc * col)] = rem[(size_t)i] % c->dims[q] + 1; rem[(size_t)i] /= c->dims[q]; }
            ++col;
        }
    }
    TRY(c->samp_piv.reserve((size_t)nc * (N - 1) * 8));
    CUDA_TRY(cudaMemcpyAsync(c->samp_piv.p, coords.data(), (size_t)nc * (N - 1) * 8, cudaMemcpyHostToDevice, c->stream));
    TRY(c->qr_A.reserve((size_t)R * nc * 8));
    TRY(k_pivot_hadamard_t(c, mode, nc, c->samp_piv.as<int64_t>(), c->qr_A.as<double>()));   // ffkrpn' (SEQRCS.jl:220-221)
    const int64_t nr = std::min<int64_t>(R, nc);
    TRY(k_qrcp_wide(c, c->qr_A.as<double>(), R, nc, nr, c->qr_piv.as<int64_t>(), c->qr_rdiag.as<double>()));
    std::vector<int64_t> p_sub((size_t)nc);
    CUDA_TRY(cudaMemcpyAsync(p_sub.data(), c->qr_piv.p, (size_t)nc * 8, cudaMemcpyDeviceToHost, c->stream));
    if (rdiag_out) CUDA_TRY(cudaMemcpyAsync(rdiag_out, c->qr_rdiag.p, (size_t)nr * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    int64_t w = 0;
    for (int64_t i = 0; i < nc; ++i) piv_out[w++] = cand[(size_t)p_sub[(size_t)i]] + 1;
    for (int64_t col = 0; col < n; ++col)
        if (!seen[(size_t)col]) piv_out[w++] = col + 1;
    if (nrdiag_out) *nrdiag_out = nr;
    if (ncand_out) *ncand_out = nc;
    return ITCPD_OK;
}

int itcpd_set_projector(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *host_pivots) {
    CHECK_CTX(c);
    NEED_T(c);
    CHECK_MODE(c, mode);
    ARG_CHECK(nsamp >= 1 && host_pivots && c->has_tensor, "bad nsamp / null pointer / no tensor");
    USE_DEVICE(c);
    TRY(check_pivots(c, mode, nsamp, host_pivots));
    const size_t pb = (size_t)nsamp * (c->order - 1) * 8;
    TRY(c->proj_piv[mode].reserve(pb));
    TRY(c->proj_T[mode].reserve((size_t)nsamp * c->dims[mode] * 8));
    CUDA_TRY(cudaMemcpyAsync(c->proj_piv[mode].p, host_pivots, pb, cudaMemcpyHostToDevice, c->stream));
    if (comm_active(c)) TRY(sharded_gather_fibers(c, mode, nsamp, c->proj_piv[mode].as<int64_t>(), c->proj_T[mode].as<double>()));
    else TRY(k_gather_fibers(c, mode, nsamp, c->proj_piv[mode].as<int64_t>(), c->proj_T[mode].as<double>()));
    c->proj_n[mode] = nsamp;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_projected_update(itcpd_ctx *c, int mode, double chol_tol, int normal) {
    CHECK_CTX(c);
    CHECK_MODE(c, mode);
    USE_DEVICE(c);
    ARG_CHECK(c->proj_n[mode] > 0, "call itcpd_set_projector(mode) first");
    ARG_CHECK(c->rank > 0 && c->A[mode].p, "CPD state not set");
    const int R = c->rank;
    const int64_t I = c->dims[mode], ns = c->proj_n[mode];
    if (comm_active(c)) {
        TRY(sharded_sampled_update(c, mode, ns, c->proj_piv[mode].as<int64_t>(), c->proj_T[mode].as<double>(), chol_tol, normal, false));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return ITCPD_OK;
    }
    TRY(c->samp_K.reserve((size_t)ns * R * 8));
    TRY(k_pivot_hadamard(c, mode, ns, c->proj_piv[mode].as<int64_t>(), c->samp_K.as<double>()));
    TRY(sampled_ls(c, mode, c->samp_K.as<double>(), c->proj_T[mode].as<double>(), c->proj_piv[mode].as<int64_t>(), ns, chol_tol, normal));
    TRY(k_colnorm_scale(c, c->X.as<double>(), I, R, c->A[mode].as<double>(), c->lambda.as<double>(), false));
    c->fver[mode]++;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

int itcpd_drop_tensor(itcpd_ctx *c) {
    CHECK_CTX(c);
    USE_DEVICE(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->T.release();
    c->PA.buf.release();
    c->PB.buf.release();
    c->qr_A.release();
    c->i8_apack[0].buf.release(); c->i8_apack[1].buf.release();
    c->i8_apack[0].valid = c->i8_apack[1].valid = false;
    c->PA.valid = c->PB.valid = false;
    c->has_tensor_data = false;
    return ITCPD_OK;
}

// one mode of the leverage-score sampled solver from DEVICE pivots: sampled KRP rows, sampled least squares, row_norm, leverage refresh
static int sampled_mode_update_device(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *piv, double chol_tol, int normal) {
    const int R = c->rank;
    const int64_t I = c->dims[mode];
    TRY(c->samp_K.reserve((size_t)nsamp * R * 8));
    TRY(k_pivot_hadamard(c, mode, nsamp, piv, c->samp_K.as<double>()));
    const double *Ts = nullptr;
    if (!normal) {   // the QR least squares works on the gathered unfolding itself
        TRY(c->samp_T.reserve((size_t)nsamp * I * 8));
        TRY(k_gather_fibers(c, mode, nsamp, piv, c->samp_T.as<double>()));
        Ts = c->samp_T.as<double>();
    }
    TRY(sampled_ls(c, mode, c->samp_K.as<double>(), Ts, piv, nsamp, chol_tol, normal));
    TRY(k_colnorm_scale(c, c->X.as<double>(), I, R, c->A[mode].as<double>(), c->lambda.as<double>(), false));
    c->fver[mode]++;
    return ensure_leverage(c, mode);  // also refreshes G[mode] (post_solve of LevScoreSampled, krp_lev...:55-58)
}

int itcpd_sampled_update(itcpd_ctx *c, int mode, int64_t nsamp, const int64_t *host_pivots, double chol_tol, int normal) {
    CHECK_CTX(c);
    NEED_T(c);
    CHECK_MODE(c, mode);
    ARG_CHECK(nsamp >= 1 && host_pivots && c->has_tensor, "bad nsamp / null pointer / no tensor");
    USE_DEVICE(c);
    TRY(ensure_cpd_buffers(c));
    TRY(upload_pivots(c, mode, nsamp, host_pivots));
    if (comm_active(c)) {
        TRY(sharded_sampled_update(c, mode, nsamp, c->samp_piv.as<int64_t>(), nullptr, chol_tol, normal, true));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return ITCPD_OK;
    }
    TRY(sampled_mode_update_device(c, mode, nsamp, c->samp_piv.as<int64_t>(), chol_tol, normal));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITCPD_OK;
}

// ---- device-resident sweeps of the leverage-score sampled solver -----------------------------------------------------------
// One sweep = for every mode: weighted draw (seeded by the device-side draw counter), sampled Khatri-Rao rows, sampled normal
// equations or QR least squares, row_norm, leverage refresh -- the ProjectionAlgorithm hooks of optimize.jl:17-28 for
// LevScoreSampled (krp_lev_score_sampled.jl:9-58) with no host round trip.  Same kernels and seeds as the per-mode entry points
// (itcpd_sample_factor_matrices + itcpd_sampled_update): the two drivers are bitwise equal.  After one plain sweep has sized every
// buffer the sweep body is captured into a CUDA graph; the draw counter lives in device memory, so a replay draws fresh samples.
__global__ void set_counter_kernel(unsigned long long *ctr, unsigned long long v) { *ctr = v; }

static int one_sampled_sweep_device(itcpd_ctx *c, const int64_t *nsamp, double chol_tol, int normal) {
    for (int mode = 0; mode < c->order; ++mode) {
        TRY(k_sample_rows(c, mode, nsamp[mode], 0, c->samp_piv.as<int64_t>(), c->draw_counter.as<unsigned long long>()));
        TRY(sampled_mode_update_device(c, mode, nsamp[mode], c->samp_piv.as<int64_t>(), chol_tol, normal));
    }
    return ITCPD_OK;
}

int itcpd_sampled_sweep_async(itcpd_ctx *c, int nsweeps, const int64_t *nsamp, uint64_t draw_counter, double chol_tol, int normal) {
    CHECK_CTX(c);
    NEED_T(c);
    ARG_CHECK(nsweeps >= 1 && nsamp != nullptr, "bad nsweeps / null nsamp");
    ARG_CHECK(!comm_active(c), "the device-resident sampled sweep is single-GPU; a sharded handle goes through itcpd_sampled_update per mode");
    USE_DEVICE(c);
    TRY(ensure_cpd_buffers(c));
    const int N = c->order, R = c->rank;
    int64_t ns_max = 0, it_max = 0;
    for (int m = 0; m < N; ++m) {
        ARG_CHECK(nsamp[m] >= 1, "nsamp must be positive for every mode");
        ARG_CHECK(normal || nsamp[m] >= R, "normal=false needs at least R samples");
        ns_max = std::max(ns_max, nsamp[m]);
        it_max = std::max(it_max, nsamp[m] * c->dims[m]);
    }
    TRY(c->samp_piv.reserve((size_t)ns_max * (N - 1) * 8));
    TRY(c->samp_K.reserve((size_t)ns_max * R * 8));
    if (!normal) TRY(c->samp_T.reserve((size_t)it_max * 8));
    TRY(c->draw_counter.reserve(64));
    for (int m = 0; m < N; ++m) TRY(ensure_leverage(c, m));
    set_counter_kernel<<<1, 1, 0, c->stream>>>(c->draw_counter.as<unsigned long long>(), (unsigned long long)draw_counter);
    c->launches++;
    CUDA_TRY(cudaGetLastError());

    auto make_key = [&](int64_t key[32]) {
        int k = 0;
        memset(key, 0, 32 * sizeof(int64_t));
        key[k++] = (int64_t)(intptr_t)c->T.p; key[k++] = (int64_t)(intptr_t)c->A[0].p; key[k++] = (int64_t)(intptr_t)c->samp_K.p;
        key[k++] = (int64_t)(intptr_t)c->samp_piv.p; key[k++] = (int64_t)(intptr_t)c->work2.p; key[k++] = (int64_t)(intptr_t)c->work.p;
        key[k++] = (int64_t)(intptr_t)c->redux.p; key[k++] = (int64_t)(intptr_t)c->samp_T.p;
        key[k++] = N; key[k++] = R; key[k++] = normal; key[k++] = c->chol_alg; key[k++] = c->graph_epoch;
        memcpy(&key[k++], &chol_tol, 8);
        for (int n = 0; n < ITCPD_MAX_ORDER; ++n) { key[k++] = n < N ? c->dims[n] : 0; key[k++] = n < N ? nsamp[n] : 0; }
    };
    auto host_bookkeeping = [&]() {   // what one_sampled_sweep_device does to the host state
        for (int n = 0; n < N; ++n) { c->fver[n]++; c->lev_ver[n] = c->fver[n]; c->m_valid[n] = false; }
    };
    int done = 0;
    int64_t key[32];
    make_key(key);
    const bool have_exec = c->sampled_graph_exec && memcmp(key, c->sampled_graph_key, sizeof(key)) == 0;
    const bool warm = c->sampled_plain_key_valid && memcmp(key, c->sampled_plain_key, sizeof(key)) == 0;
    if (c->use_graph && (nsweeps >= 3 || have_exec || warm)) {
        if (!have_exec) {
            if (c->sampled_graph_exec) { cudaGraphExecDestroy(c->sampled_graph_exec); c->sampled_graph_exec = nullptr; }
            if (!warm) {
                TRY(one_sampled_sweep_device(c, nsamp, chol_tol, normal));   // plain sweep: sizes every scratch buffer
                done = 1;
            }
            make_key(key);
            const int64_t l0 = c->launches;
            uint64_t fver0[ITCPD_MAX_ORDER], lev0[ITCPD_MAX_ORDER];
            memcpy(fver0, c->fver, sizeof(fver0));
            memcpy(lev0, c->lev_ver, sizeof(lev0));
            cudaGraph_t graph = nullptr;
            CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
            const int st = one_sampled_sweep_device(c, nsamp, chol_tol, normal);
            const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            memcpy(c->fver, fver0, sizeof(fver0));     // nothing ran during the capture
            memcpy(c->lev_ver, lev0, sizeof(lev0));
            c->sampled_graph_launches = c->launches - l0;
            c->launches = l0;
            if (st != ITCPD_OK || e != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                if (st != ITCPD_OK) return st;
                set_error("CUDA graph capture of the sampled sweep failed: %s", cudaGetErrorString(e));
                return ITCPD_ERR_CUDA;
            }
            const cudaError_t e2 = cudaGraphInstantiate(&c->sampled_graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e2 != cudaSuccess) { c->sampled_graph_exec = nullptr; set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e2)); return ITCPD_ERR_CUDA; }
            memcpy(c->sampled_graph_key, key, sizeof(key));
        }
        for (; done < nsweeps; ++done) {
            CUDA_TRY(cudaGraphLaunch(c->sampled_graph_exec, c->stream));
            c->launches += c->sampled_graph_launches;
            host_bookkeeping();
        }
    }
    if (done < nsweeps) {
        for (; done < nsweeps; ++done) TRY(one_sampled_sweep_device(c, nsamp, chol_tol, normal));
        make_key(c->sampled_plain_key);
        c->sampled_plain_key_valid = true;
    }
    return ITCPD_OK;
}

int itcpd_allgather_factor(itcpd_ctx *c, int mode, int64_t rows_total, double *host_out) {
    CHECK_CTX(c);
    CHECK_MODE(c, mode);
    ARG_CHECK(host_out != nullptr, "null out pointer");
    USE_DEVICE(c);
    const int R = c->rank;
    if (!comm_active(c) || mode != c->order - 1) {
        ARG_CHECK(rows_total == c->dims[mode], "rows_total must equal the local row count for a replicated factor");
        return d2h(c, host_out, c->A[mode].p, (size_t)rows_total * R * 8);
    }
    const int nr = comm_size(c);
    const int64_t loc = c->dims[mode];
    ARG_CHECK(loc * nr == rows_total, "slabs must be equal-sized for the all-gather");
    // gather rank-major [rank][loc x R] then interleave into the rows_total x R column-major factor on the host
    TRY(c->work.reserve((size_t)rows_total * R * 8));
    TRY(comm_allgather(c, c->A[mode].as<double>(), c->work.as<double>(), loc * R));
    std::vector<double> tmp((size_t)rows_total * R);
    TRY(d2h(c, tmp.data(), c->work.p, (size_t)rows_total * R * 8));
    for (int g = 0; g < nr; ++g)
        for (int r = 0; r < R; ++r)
            memcpy(host_out + (size_t)g * loc + (size_t)rows_total * r, tmp.data() + ((size_t)g * R + r) * loc, (size_t)loc * 8);
    return ITCPD_OK;
}

// ---- peer-memory exchange (CUDA IPC over NVLink) ----------------------------------------------------
int itcpd_peer_export(itcpd_ctx *c, void *handle64_out) {
    CHECK_CTX(c);
    ARG_CHECK(handle64_out && c->has_tensor && c->rank > 0, "set the tensor slab and the rank before exporting");
    USE_DEVICE(c);
    int64_t maxrows = 0;
    for (int n = 0; n + 1 < c->order; ++n) maxrows = std::max(maxrows, c->dims[n]);
    c->peer_slot_doubles = maxrows * c->rank;
    c->peer_slots = std::max(2, c->order - 1);  // peer_graph uses one partial-M slot per exchange of a sweep
    c->peer_small_doubles = (int64_t)c->rank * c->rank + 64;
    c->peer_small_off = 256 + (size_t)c->peer_slots * (size_t)c->peer_slot_doubles * 8;
    c->xchg.release();
    TRY(c->xchg.reserve(c->peer_small_off + 2 * (size_t)c->peer_small_doubles * 8));
    CUDA_TRY(cudaMemset(c->xchg.p, 0, 256));
    TRY(c->peer_epochs.reserve(64));
    CUDA_TRY(cudaMemset(c->peer_epochs.p, 0, 64));
    CUDA_TRY(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, c->xchg.p));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle64_out, &h, 64);
    c->peer_on = false;
    return ITCPD_OK;
}

int itcpd_peer_import(itcpd_ctx *c, int nranks, int rank, const void *handles) {
    CHECK_CTX(c);
    ARG_CHECK(handles && nranks >= 2 && nranks <= ITCPD_MAX_PEERS && rank >= 0 && rank < nranks && c->xchg.p, "bad peer_import arguments (export first)");
    USE_DEVICE(c);
    for (int q = 0; q < nranks; ++q) {
        if (q == rank) { c->peer_base[q] = c->xchg.p; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * (size_t)q, 64);
        void *ptr = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_base[q] = ptr;
    }
    c->peer_n = nranks;
    c->peer_rank = rank;
    c->peer_epoch = 0;
    c->peer_on = true;
    c->graph_epoch++;
    return ITCPD_OK;
}

int itcpd_peer_disable(itcpd_ctx *c) {
    CHECK_CTX(c);
    USE_DEVICE(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int q = 0; q < c->peer_n; ++q)
        if (q != c->peer_rank && c->peer_base[q]) cudaIpcCloseMemHandle(c->peer_base[q]);
    for (int q = 0; q < ITCPD_MAX_PEERS; ++q) c->peer_base[q] = nullptr;
    c->peer_on = false;
    c->peer_n = 0;
    return ITCPD_OK;
}

// ---- measurement ----------------------------------------------------------------------------------
int itcpd_gemm_timing(itcpd_ctx *c, int reset, double *avg_ms, int64_t *launches) {
    CHECK_CTX(c);
    USE_DEVICE(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    double tot = 0.0;
    for (size_t i = 0; i < c->gemm_events_used; ++i) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, c->gemm_events[i].first, c->gemm_events[i].second));
        tot += ms;
    }
    if (avg_ms) *avg_ms = c->gemm_events_used ? tot / (double)c->gemm_events_used : 0.0;
    if (launches) *launches = (int64_t)c->gemm_events_used;
    if (reset) c->gemm_events_used = 0;
    return ITCPD_OK;
}

int itcpd_phase_timing(itcpd_ctx *c, int reset, int nphases, double *ms_by_phase, int64_t *marks) {
    CHECK_CTX(c);
    ARG_CHECK(nphases >= 1 && ms_by_phase, "bad phase_timing arguments");
    USE_DEVICE(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < nphases; ++i) ms_by_phase[i] = 0.0;
    for (size_t i = 1; i < c->phase_used; ++i) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, c->phase_events[i - 1], c->phase_events[i]));
        const int id = c->phase_ids[i];
        if (id >= 0 && id < nphases) ms_by_phase[id] += ms;
    }
    if (marks) *marks = (int64_t)c->phase_used;
    if (reset) c->phase_used = 0;
    return ITCPD_OK;
}

int itcpd_probe_dmma_peak(itcpd_ctx *c, double *tflops) {
    CHECK_CTX(c);
    ARG_CHECK(tflops != nullptr, "null out pointer");
    USE_DEVICE(c);
    return probe_dmma(c, tflops);
}

int itcpd_probe_dfma_peak(itcpd_ctx *c, double *tflops) {
    CHECK_CTX(c);
    ARG_CHECK(tflops != nullptr, "null out pointer");
    USE_DEVICE(c);
    return probe_dfma(c, tflops);
}

int itcpd_event_record(itcpd_ctx *c, int slot) {
    CHECK_CTX(c);
    ARG_CHECK(slot >= 0 && slot < 16, "event slot must be in [0,16)");
    USE_DEVICE(c);
    if (!c->user_events[slot]) CUDA_TRY(cudaEventCreate(&c->user_events[slot]));
    CUDA_TRY(cudaEventRecord(c->user_events[slot], c->stream));
    return ITCPD_OK;
}

int itcpd_event_elapsed_ms(itcpd_ctx *c, int s0, int s1, double *ms) {
    CHECK_CTX(c);
    ARG_CHECK(s0 >= 0 && s0 < 16 && s1 >= 0 && s1 < 16 && ms && c->user_events[s0] && c->user_events[s1], "bad event slots");
    USE_DEVICE(c);
    CUDA_TRY(cudaEventSynchronize(c->user_events[s1]));
    float f = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&f, c->user_events[s0], c->user_events[s1]));
    *ms = f;
    return ITCPD_OK;
}

int itcpd_host_alloc(int64_t bytes, void **out) {
    ARG_CHECK(bytes > 0 && out, "bad host_alloc arguments");
    CUDA_TRY(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return ITCPD_OK;
}

int itcpd_host_free(void *p) {
    if (p) CUDA_TRY(cudaFreeHost(p));
    return ITCPD_OK;
}

int itcpd_flush_l2(itcpd_ctx *c, int64_t bytes) {
    CHECK_CTX(c);
    ARG_CHECK(bytes > 0, "bytes must be positive");
    USE_DEVICE(c);
    TRY(c->flush.reserve((size_t)bytes));
    CUDA_TRY(cudaMemsetAsync(c->flush.p, 1, (size_t)bytes, c->stream));
    c->launches++;
    return ITCPD_OK;
}

}  // extern "C"
