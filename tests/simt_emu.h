// Test infrastructure: a minimal SIMT emulator for single-CTA CUDA kernels.  Every CUDA thread is an OS thread;
// __syncthreads / bar.sync / __syncwarp are sense-reversing barriers built from C++ atomics (ThreadSanitizer models their
// acquire/release edges exactly; with pthread_barrier_t it missed every race in a negative control), warp collectives
// exchange through a per-warp scratch row between two warp barriers.  Kernels are compiled from their verbatim .cu text behind these macros, so ThreadSanitizer
// sees every shared-memory access with exactly the happens-before edges the CUDA barriers provide.
// Valid for kernels whose collectives are executed by full, convergent warps (true for the solve.cu kernels).
#pragma once
#include <pthread.h>
#include <sched.h>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(x) __attribute__((aligned(x)))

struct dim3_ { unsigned x, y, z; };
static thread_local dim3_ threadIdx, blockIdx;
static dim3_ blockDim, gridDim;
struct double2 { double x, y; };

namespace emu {
struct SpinBarrier {
    std::atomic<int> count{0}, sense{0};
    int n = 0;
    void init(int participants) { count.store(0); sense.store(0); n = participants; }
    void wait() {
        const int s = sense.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
            count.store(0, std::memory_order_relaxed);
            sense.store(s ^ 1, std::memory_order_release);
        } else {
            while (sense.load(std::memory_order_acquire) == s) sched_yield();
        }
    }
};
// Warp collectives (__shfl*_sync, __reduce_*_sync, __any_sync) converge the warp but give NO memory ordering on the
// hardware, so their rendezvous uses relaxed atomics only: ThreadSanitizer then derives no happens-before edge from a
// collective, and a kernel that leans on one to order shared-memory traffic is reported as racy.
struct RelaxedBarrier {
    std::atomic<int> count{0}, sense{0};
    void init() { count.store(0); sense.store(0); }
    void wait() {
        const int s = sense.load(std::memory_order_relaxed);
        if (count.fetch_add(1, std::memory_order_relaxed) == 31) {
            count.store(0, std::memory_order_relaxed);
            sense.store(s ^ 1, std::memory_order_relaxed);
        } else {
            while (sense.load(std::memory_order_relaxed) == s) sched_yield();
        }
    }
};
static SpinBarrier cta_bar, named_bar, warp_bar[32];
static RelaxedBarrier coll_bar[32];
static std::atomic<unsigned long long> scratch[32][32];
inline int warp() { return threadIdx.x >> 5; }
inline int lane() { return threadIdx.x & 31; }
inline void wbar() { warp_bar[warp()].wait(); }
template <typename T, typename F>
inline T collective(T v, F combine) {
    static_assert(sizeof(T) <= 8, "scratch cell");
    unsigned long long cell = 0;
    memcpy(&cell, &v, sizeof(T));
    scratch[warp()][lane()].store(cell, std::memory_order_relaxed);
    coll_bar[warp()].wait();
    T vals[32];
    for (int i = 0; i < 32; ++i) {
        const unsigned long long c = scratch[warp()][i].load(std::memory_order_relaxed);
        memcpy(&vals[i], &c, sizeof(T));
    }
    coll_bar[warp()].wait();
    return combine(vals);
}
}  // namespace emu

inline void __syncthreads() { emu::cta_bar.wait(); }
inline void __syncwarp() { emu::wbar(); }
inline void emu_named_barrier() { emu::named_bar.wait(); }
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    return emu::collective<unsigned>(v, [](unsigned *a) { unsigned r = a[0]; for (int i = 1; i < 32; ++i) r = a[i] > r ? a[i] : r; return r; });
}
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    return emu::collective<unsigned>(v, [](unsigned *a) { unsigned r = a[0]; for (int i = 1; i < 32; ++i) r = a[i] < r ? a[i] : r; return r; });
}
inline int __any_sync(unsigned, int p) {
    return emu::collective<int>(p != 0, [](int *a) { int r = 0; for (int i = 0; i < 32; ++i) r |= a[i]; return r; });
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, int o) {
    const int l = emu::lane();
    return emu::collective<T>(v, [l, o](T *a) { return (l + o < 32) ? a[l + o] : a[l]; });
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) {
    return emu::collective<T>(v, [src](T *a) { return a[src & 31]; });
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m) {
    const int l = emu::lane();
    return emu::collective<T>(v, [l, m](T *a) { return a[(l ^ m) & 31]; });
}
inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }

// run `body` as one CTA of `threads` CUDA threads; `named` = participant count of `bar.sync 1, named`
inline void emu_launch(int threads, int named, const std::function<void()> &body, unsigned bx = 0, unsigned by = 0, unsigned gx = 1, unsigned gy = 1) {
    blockDim = {(unsigned)threads, 1, 1};
    gridDim = {gx, gy, 1};
    emu::cta_bar.init(threads);
    emu::named_bar.init(named > 0 ? named : threads);
    for (int w = 0; w < (threads + 31) / 32; ++w) { emu::warp_bar[w].init(32); emu::coll_bar[w].init(); }
    struct Arg { int t; unsigned bx, by; const std::function<void()> *f; };
    std::vector<pthread_t> th(threads);
    std::vector<Arg> args(threads);
    for (int t = 0; t < threads; ++t) {
        args[t] = {t, bx, by, &body};
        pthread_create(&th[t], nullptr, [](void *p) -> void * {
            Arg *a = (Arg *)p;
            threadIdx = {(unsigned)a->t, 0, 0};
            blockIdx = {a->bx, a->by, 0};
            (*a->f)();
            return nullptr;
        }, &args[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], nullptr);
}
