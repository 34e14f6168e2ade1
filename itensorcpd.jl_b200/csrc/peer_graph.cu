// NCCL-free sharded sweeps that can be captured in a CUDA graph (option "peer_graph", OFF by default).
// The default sharded sweep passes the exchange epoch and the slot pointer as kernel ARGUMENTS and all-reduces the small
// last-mode quantities (column sums of squares, Gram) with NCCL, so it cannot be replayed from a graph: a replay would
// reuse the captured epoch, and capturing NCCL calls dead-locked in round 1.  Here
//   * the epoch lives in DEVICE memory: peer_signal_dev_kernel advances and publishes it, peer_wait_dev_kernel spins
//     on the peers' flags against it, and the (unchanged) row-solve kernel then sums the peers' slots while loading;
//   * the partial-M slot is the exchange's index within the sweep (constant across replays; with >= 2 exchanges per
//     sweep a slot is rewritten only after every peer has passed a later exchange, so the in-order streams make it safe);
//   * the small all-reduces are one kernel over the same peer-mapped exchange buffer: write my partial, publish, wait,
//     sum in rank order (bitwise identical on every rank), double buffered by the device epoch's parity.
// STATUS: compiled; every GPU-validated kernel is untouched (tools/sass_guard.py).  Not yet run on hardware.
#include "common.cuh"
#include "peer.cuh"

namespace itcpd {

__global__ void peer_signal_dev_kernel(PeerPtrs f, long long *epoch_dev) {
    __shared__ long long s_e;
    __threadfence_system();
    if (threadIdx.x == 0) { s_e = *epoch_dev + 1; *epoch_dev = s_e; }
    __syncthreads();
    if ((int)threadIdx.x < f.n) {
        volatile long long *d = reinterpret_cast<long long *>(f.base[threadIdx.x]) + f.rank;
        *d = s_e;
    }
    __threadfence_system();
}

__global__ void peer_wait_dev_kernel(const volatile long long *flags, int n, const long long *epoch_dev) {
    if ((int)threadIdx.x < n) bounded_wait(flags + threadIdx.x, *epoch_dev, (int)threadIdx.x);
    __syncthreads();
    __threadfence_system();
}

__global__ void __launch_bounds__(1024) peer_allreduce_small_kernel(PeerPtrs f, long long *epoch_dev, size_t small_off, int64_t slot_doubles,
                                                                   double *buf, int n) {
    __shared__ long long s_e;
    if (threadIdx.x == 0) { s_e = *epoch_dev + 1; *epoch_dev = s_e; }
    __syncthreads();
    const long long e = s_e;
    const size_t off = small_off + (size_t)(e & 1) * (size_t)slot_doubles * 8;
    double *mine = reinterpret_cast<double *>(f.base[f.rank] + off);
    for (int i = threadIdx.x; i < n; i += blockDim.x) mine[i] = buf[i];
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < f.n) {
        volatile long long *d = reinterpret_cast<long long *>(f.base[threadIdx.x]) + 16 + f.rank;
        *d = e;
        bounded_wait(reinterpret_cast<const volatile long long *>(f.base[f.rank]) + 16 + threadIdx.x, e, (int)threadIdx.x);
    }
    __syncthreads();
    __threadfence_system();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {   // 1024 threads: a handful of remote elements per thread (was 16 x G at 256)
        double v = 0.0;
        for (int q = 0; q < f.n; ++q) v += reinterpret_cast<const volatile double *>(f.base[q] + off)[i];
        buf[i] = v;
    }
}

bool peer_graph_active(const itcpd_ctx *c) { return c->peer_on && c->peer_graph && comm_active(c) && c->order >= 3; }

int peer_graph_signal(itcpd_ctx *c) {
    peer_signal_dev_kernel<<<1, 32, 0, c->stream>>>(peer_ptrs(c), c->peer_epochs.as<long long>());
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

int peer_graph_wait(itcpd_ctx *c) {
    peer_wait_dev_kernel<<<1, 32, 0, c->stream>>>(reinterpret_cast<const volatile long long *>(c->xchg.p), c->peer_n, c->peer_epochs.as<long long>());
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

int peer_allreduce_small(itcpd_ctx *c, double *buf, int64_t n) {
    ARG_CHECK(n <= c->peer_small_doubles, "small all-reduce larger than the exchange buffer's small slots");
    peer_allreduce_small_kernel<<<1, 1024, 0, c->stream>>>(peer_ptrs(c), c->peer_epochs.as<long long>() + 1, c->peer_small_off, c->peer_small_doubles, buf, (int)n);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ITCPD_OK;
}

}  // namespace itcpd
