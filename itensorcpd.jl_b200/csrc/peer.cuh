// Peer-memory helpers of peer_graph.cu: the table of peer-mapped exchange buffers and the bounded
// system-scope flag wait (a peer that died must surface as a launch failure on this rank, not as a hung box).
#pragma once
#include "common.cuh"

namespace itcpd {

struct PeerPtrs { char *base[ITCPD_MAX_PEERS]; int n; int rank; };

static inline PeerPtrs peer_ptrs(const itcpd_ctx *c) {
    PeerPtrs f;
    memset(&f, 0, sizeof(f));
    f.n = c->peer_n;
    f.rank = c->peer_rank;
    for (int q = 0; q < c->peer_n; ++q) f.base[q] = reinterpret_cast<char *>(c->peer_base[q]);
    return f;
}

__device__ __forceinline__ void bounded_wait(const volatile long long *flag, long long epoch, int who) {
    unsigned long long t0 = 0, spins = 0;
    while (*flag < epoch) {
        if ((++spins & 0xfffff) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > ITCPD_PEER_TIMEOUT_NS) {
                printf("itcpd: peer %d never published exchange %lld\n", who, epoch);
                __trap();
            }
        }
    }
}


}  // namespace itcpd
