// Test infrastructure (CPU only): a validating stand-in for libcudart.so.12 (and libnccl.so.2) behind which the HOST side of
// libitcpd_b200 runs without a GPU ("dry run").  Kernels are never executed: a launch is checked against the limits that make
// real launches fail (grid / block dimensions, registers per block, static + dynamic shared memory against the function
// attribute), device memory is lazily mapped host memory with an allocation table (copies and memsets are range-checked,
// large ones are not performed), TMA descriptors are checked against cuTensorMapEncodeTiled's documented constraints, and
// stream capture is tracked as a dependency graph: work that synchronises a capturing stream, NCCL calls inside a capture and
// forked streams that never re-join the origin are reported.  tests/test_dry_run_cpu.py drives the real C-ABI through it over
// the BASELINE shapes, the sharded paths and the experimental options and requires an empty violation list.
#include <cuda.h>
#include <cuda_runtime_api.h>
#include <sys/mman.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

namespace {
std::mutex g_mu;
std::vector<std::string> g_violations;
std::map<std::string, long> g_launch_count;
cudaError_t g_last_error = cudaSuccess;

void violation(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_violations.push_back(buf);
}

// ---- device memory -------------------------------------------------------------------------------------------------
struct Alloc { size_t size; bool host; };
std::map<uintptr_t, Alloc> g_allocs;
size_t g_device_bytes = 0;
size_t g_device_cap = (size_t)180 * 1000 * 1000 * 1000;
size_t REAL_COPY_LIMIT = (size_t)1 << 20;   // larger copies / memsets are only range-checked (fakecuda_set_real_copy_limit)
// Optional launch hook (tests/test_i8_host_device_cpu.py): a plug-in that EXECUTES selected kernels on the host -- the verbatim
// gemm_i8.cu kernels on the functional tcgen05 / TMEM / TMA model -- so that the real host code drives real (emulated) device code.
typedef int (*launch_hook_t)(const char *name, void **args, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, size_t smem);
launch_hook_t g_launch_hook = nullptr;

const std::pair<const uintptr_t, Alloc> *find_alloc(const void *p) {
    const uintptr_t a = (uintptr_t)p;
    auto it = g_allocs.upper_bound(a);
    if (it == g_allocs.begin()) return nullptr;
    --it;
    if (a >= it->first && a < it->first + it->second.size) return &*it;
    return nullptr;
}
// a device range must lie inside one live device allocation
bool check_device_range(const void *p, size_t n, const char *what) {
    if (n == 0) return true;
    auto *al = find_alloc(p);
    if (!al || al->second.host) { violation("%s: %p (+%zu) is not inside a live device allocation", what, p, n); return false; }
    if ((uintptr_t)p + n > al->first + al->second.size) {
        violation("%s: range %p +%zu overruns its allocation (%zu bytes at %p)", what, p, n, al->second.size, (void *)al->first);
        return false;
    }
    return true;
}
bool is_device_ptr(const void *p) { auto *al = find_alloc(p); return al && !al->second.host; }
bool is_pinned_ptr(const void *p) { auto *al = find_alloc(p); return al && al->second.host; }

// ---- kernels -------------------------------------------------------------------------------------------------------
struct KernelInfo { std::string name; int regs = 0, static_smem = 0; int max_dyn_smem = 48 * 1024; };
std::map<const void *, KernelInfo> g_kernels;
std::map<std::string, std::pair<int, int>> g_res_table;   // mangled name -> (registers, static shared memory)
bool g_table_loaded = false;
void load_table() {
    if (g_table_loaded) return;
    g_table_loaded = true;
    const char *path = getenv("FAKECUDA_KERNEL_TABLE");
    if (!path) return;
    FILE *f = fopen(path, "r");
    if (!f) return;
    char name[2048];
    int regs, sh;
    while (fscanf(f, "%2047s %d %d", name, &regs, &sh) == 3) g_res_table[name] = {regs, sh};
    fclose(f);
}

// ---- streams, events, capture ----------------------------------------------------------------------------------------
struct Node { std::vector<int> deps; };
struct Capture {
    std::vector<Node> nodes;
    std::set<void *> streams;   // streams currently part of the capture
    void *origin = nullptr;
    int nccl_calls = 0;
};
struct Stream { int id; Capture *cap = nullptr; int last_node = -1; };
struct Event { bool recorded = false; Capture *cap = nullptr; int node = -1; };
std::set<Stream *> g_streams;
std::set<Event *> g_events;
int g_next_stream = 1;
Stream g_default_stream{0};

Stream *as_stream(cudaStream_t s, const char *what) {
    if (s == nullptr) return &g_default_stream;
    Stream *st = reinterpret_cast<Stream *>(s);
    if (!g_streams.count(st)) { violation("%s: invalid stream handle %p", what, (void *)s); return nullptr; }
    return st;
}
Event *as_event(cudaEvent_t e, const char *what) {
    Event *ev = reinterpret_cast<Event *>(e);
    if (!ev || !g_events.count(ev)) { violation("%s: invalid event handle %p", what, (void *)e); return nullptr; }
    return ev;
}
int add_node(Stream *st) {
    Capture *c = st->cap;
    Node n;
    if (st->last_node >= 0) n.deps.push_back(st->last_node);
    c->nodes.push_back(n);
    st->last_node = (int)c->nodes.size() - 1;
    return st->last_node;
}
void stream_work(Stream *st) { if (st && st->cap) add_node(st); }
bool reaches(const Capture *c, int from, int target, std::vector<char> &seen) {
    if (from == target) return true;
    if (from < 0 || seen[from]) return false;
    seen[from] = 1;
    for (int d : c->nodes[from].deps)
        if (reaches(c, d, target, seen)) return true;
    return false;
}
struct Graph { int nodes; };
}  // namespace

// exported to the Python test (not part of any CUDA API)
extern "C" {
int fakecuda_violation_count() { std::lock_guard<std::mutex> g(g_mu); return (int)g_violations.size(); }
int fakecuda_violation(int i, char *buf, int n) {
    std::lock_guard<std::mutex> g(g_mu);
    if (i < 0 || i >= (int)g_violations.size()) return -1;
    snprintf(buf, n, "%s", g_violations[i].c_str());
    return 0;
}
void fakecuda_clear() { std::lock_guard<std::mutex> g(g_mu); g_violations.clear(); g_launch_count.clear(); }
long fakecuda_launches(const char *substr) {
    std::lock_guard<std::mutex> g(g_mu);
    long n = 0;
    for (auto &kv : g_launch_count)
        if (!substr || !*substr || kv.first.find(substr) != std::string::npos) n += kv.second;
    return n;
}
void fakecuda_set_real_copy_limit(unsigned long long bytes) { std::lock_guard<std::mutex> g(g_mu); REAL_COPY_LIMIT = bytes; }
void fakecuda_set_launch_hook(void *fn) { std::lock_guard<std::mutex> g(g_mu); g_launch_hook = (launch_hook_t)fn; }
void fakecuda_set_device_cap(unsigned long long bytes) { std::lock_guard<std::mutex> g(g_mu); g_device_cap = bytes; }
unsigned long long fakecuda_device_bytes() { std::lock_guard<std::mutex> g(g_mu); return g_device_bytes; }
int fakecuda_live_device_allocs() {
    std::lock_guard<std::mutex> g(g_mu);
    int n = 0;
    for (auto &kv : g_allocs) n += !kv.second.host;
    return n;
}

// ---- registration hooks emitted by nvcc --------------------------------------------------------------------------------
void **__cudaRegisterFatBinary(void *) { static void *h = nullptr; return &h; }
void __cudaRegisterFatBinaryEnd(void **) {}
void __cudaUnregisterFatBinary(void **) {}
void __cudaRegisterFunction(void **, const char *hostFun, char *, const char *deviceName, int, uint3 *, uint3 *, dim3 *, dim3 *, int *) {
    std::lock_guard<std::mutex> g(g_mu);
    load_table();
    KernelInfo k;
    k.name = deviceName ? deviceName : "?";
    auto it = g_res_table.find(k.name);
    if (it != g_res_table.end()) { k.regs = it->second.first; k.static_smem = it->second.second; }
    g_kernels[(const void *)hostFun] = k;
}
struct CallConfig { dim3 grid, block; size_t smem; void *stream; };
static thread_local std::vector<CallConfig> t_cfg;
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t smem, struct CUstream_st *stream) {
    t_cfg.push_back({grid, block, smem, (void *)stream});
    return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3 *grid, dim3 *block, size_t *smem, void *stream) {
    if (t_cfg.empty()) return cudaErrorInvalidConfiguration;
    CallConfig c = t_cfg.back();
    t_cfg.pop_back();
    *grid = c.grid; *block = c.block; *smem = c.smem; *(void **)stream = c.stream;
    return cudaSuccess;
}

cudaError_t cudaLaunchKernel(const void *func, dim3 grid, dim3 block, void **args, size_t smem, cudaStream_t stream) {
    std::unique_lock<std::mutex> g(g_mu);
    auto it = g_kernels.find(func);
    const KernelInfo unknown{"<unregistered kernel>"};
    const KernelInfo &k = it == g_kernels.end() ? unknown : it->second;
    g_launch_count[k.name]++;
    bool bad = false;
    const unsigned long long threads = (unsigned long long)block.x * block.y * block.z;
    if (grid.x < 1 || grid.y < 1 || grid.z < 1 || grid.x > 2147483647u || grid.y > 65535u || grid.z > 65535u) {
        violation("launch %s: invalid grid (%u, %u, %u)", k.name.c_str(), grid.x, grid.y, grid.z); bad = true;
    }
    if (threads < 1 || threads > 1024 || block.z > 64) { violation("launch %s: invalid block (%u, %u, %u)", k.name.c_str(), block.x, block.y, block.z); bad = true; }
    if (smem > (size_t)k.max_dyn_smem) {
        violation("launch %s: %zu bytes of dynamic shared memory above the function's limit %d (cudaFuncSetAttribute missing?)", k.name.c_str(), smem, k.max_dyn_smem);
        bad = true;
    }
    if (smem + k.static_smem > 227 * 1024) { violation("launch %s: %zu + %d bytes of shared memory exceed 227 KB", k.name.c_str(), smem, k.static_smem); bad = true; }
    if (k.regs > 0) {
        const unsigned long long warps = (threads + 31) / 32, per_warp = (unsigned long long)((k.regs + 7) / 8 * 8) * 32;
        if (warps * per_warp > 65536) { violation("launch %s: %llu threads x %d registers exceed the register file", k.name.c_str(), threads, k.regs); bad = true; }
    }
    Stream *st = as_stream(stream, "cudaLaunchKernel");
    if (!st) bad = true;
    if (bad) { g_last_error = cudaErrorInvalidConfiguration; return g_last_error; }
    stream_work(st);
    if (g_launch_hook && !st->cap) {
        launch_hook_t hook = g_launch_hook;
        const std::string name = k.name;
        g.unlock();   // the emulated kernel runs thousands of host threads; it never calls back into the runtime
        if (hook(name.c_str(), args, grid.x, grid.y, grid.z, block.x, block.y, block.z, smem)) {
            g.lock();
            g_launch_count["<executed> " + name]++;
        }
    }
    return cudaSuccess;
}

cudaError_t cudaFuncSetAttribute(const void *func, cudaFuncAttribute attr, int value) {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_kernels.find(func);
    if (it == g_kernels.end()) { violation("cudaFuncSetAttribute on an unregistered function"); return cudaErrorInvalidDeviceFunction; }
    if (attr == cudaFuncAttributeMaxDynamicSharedMemorySize) {
        if (value + it->second.static_smem > 227 * 1024) {
            violation("cudaFuncSetAttribute(%s): %d + %d static bytes exceed the 227 KB opt-in limit", it->second.name.c_str(), value, it->second.static_smem);
            return cudaErrorInvalidValue;
        }
        it->second.max_dyn_smem = value;
    }
    return cudaSuccess;
}

// ---- device management ---------------------------------------------------------------------------------------------------
cudaError_t cudaGetDeviceCount(int *n) { *n = getenv("FAKECUDA_NO_DEVICE") ? 0 : 8; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return (d >= 0 && d < 8) ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "fake B200 (dry run)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
    p->totalGlobalMem = g_device_cap;
    p->sharedMemPerBlockOptin = 227 * 1024; p->sharedMemPerMultiprocessor = 228 * 1024; p->regsPerMultiprocessor = 65536;
    p->warpSize = 32; p->maxThreadsPerBlock = 1024; p->l2CacheSize = 126 * 1024 * 1024;
    return cudaSuccess;
}
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory (fake)" : "fake CUDA error"; }
cudaError_t cudaGetLastError() { std::lock_guard<std::mutex> g(g_mu); cudaError_t e = g_last_error; g_last_error = cudaSuccess; return e; }
cudaError_t cudaDeviceSynchronize() {
    std::lock_guard<std::mutex> g(g_mu);
    for (Stream *s : g_streams)
        if (s->cap) { violation("cudaDeviceSynchronize while stream %d is capturing", s->id); return cudaErrorStreamCaptureUnsupported; }
    return cudaSuccess;
}

// ---- memory -----------------------------------------------------------------------------------------------------------
static cudaError_t alloc_common(void **p, size_t n, bool host) {
    std::lock_guard<std::mutex> g(g_mu);
    *p = nullptr;
    if (n == 0) n = 1;
    if (!host && g_device_bytes + n > g_device_cap) { g_last_error = cudaErrorMemoryAllocation; return cudaErrorMemoryAllocation; }
    const size_t len = (n + 4095) & ~(size_t)4095;
    void *m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) { g_last_error = cudaErrorMemoryAllocation; return cudaErrorMemoryAllocation; }
    g_allocs[(uintptr_t)m] = {n, host};
    if (!host) g_device_bytes += n;
    *p = m;
    return cudaSuccess;
}
static cudaError_t free_common(void *p, bool host, const char *what) {
    std::lock_guard<std::mutex> g(g_mu);
    if (!p) return cudaSuccess;
    auto it = g_allocs.find((uintptr_t)p);
    if (it == g_allocs.end() || it->second.host != host) { violation("%s(%p): not the base of a live %s allocation", what, p, host ? "pinned" : "device"); return cudaErrorInvalidValue; }
    for (Stream *s : g_streams)
        if (s->cap) { violation("%s while stream %d is capturing", what, s->id); break; }
    if (!host) g_device_bytes -= it->second.size;
    munmap(p, (it->second.size + 4095) & ~(size_t)4095);
    g_allocs.erase(it);
    return cudaSuccess;
}
cudaError_t cudaMalloc(void **p, size_t n) { return alloc_common(p, n, false); }
cudaError_t cudaFree(void *p) { return free_common(p, false, "cudaFree"); }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return alloc_common(p, n, true); }
cudaError_t cudaFreeHost(void *p) { return free_common(p, true, "cudaFreeHost"); }

cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b) {   // a roomy device: the set-up's "is there room for a second copy" test says yes
    if (total_b) *total_b = (size_t)180 << 30;
    if (free_b) *free_b = (size_t)160 << 30;
    return cudaSuccess;
}

cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) {
    std::lock_guard<std::mutex> g(g_mu);
    memset(a, 0, sizeof(*a));
    a->type = is_device_ptr(p) ? cudaMemoryTypeDevice : (is_pinned_ptr(p) ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered);
    return cudaSuccess;
}

static cudaError_t copy_common(void *dst, const void *src, size_t n, cudaMemcpyKind kind, Stream *st, bool sync, const char *what) {
    bool ok = true;
    if (kind == cudaMemcpyHostToDevice || kind == cudaMemcpyDeviceToDevice) ok &= check_device_range(dst, n, what);
    if (kind == cudaMemcpyDeviceToHost || kind == cudaMemcpyDeviceToDevice) ok &= check_device_range(src, n, what);
    if (kind == cudaMemcpyHostToDevice && is_device_ptr(src)) { violation("%s: host source %p is device memory", what, src); ok = false; }
    if (kind == cudaMemcpyDeviceToHost && is_device_ptr(dst)) { violation("%s: host destination %p is device memory", what, dst); ok = false; }
    if (n && (!dst || !src)) { violation("%s: null pointer", what); ok = false; }
    if (sync)
        for (Stream *s : g_streams)
            if (s->cap) { violation("%s (synchronous) while stream %d is capturing", what, s->id); ok = false; break; }
    if (st && st->cap && kind != cudaMemcpyDeviceToDevice) {
        // pageable / pinned host copies can be captured, but this library promises a side-effect-free sweep body
        violation("%s: host <-> device copy recorded into a stream capture", what); ok = false;
    }
    if (!ok) { g_last_error = cudaErrorInvalidValue; return cudaErrorInvalidValue; }
    if (n <= REAL_COPY_LIMIT) memmove(dst, src, n);
    stream_work(st);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind kind) {
    std::lock_guard<std::mutex> g(g_mu);
    return copy_common(dst, src, n, kind, nullptr, true, "cudaMemcpy");
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind kind, cudaStream_t s) {
    std::lock_guard<std::mutex> g(g_mu);
    Stream *st = as_stream(s, "cudaMemcpyAsync");
    if (!st) return cudaErrorInvalidResourceHandle;
    return copy_common(dst, src, n, kind, st, false, "cudaMemcpyAsync");
}
static cudaError_t memset_common(void *p, int v, size_t n, Stream *st, const char *what) {
    if (!check_device_range(p, n, what)) { g_last_error = cudaErrorInvalidValue; return cudaErrorInvalidValue; }
    if (n <= REAL_COPY_LIMIT) memset(p, v, n);
    stream_work(st);
    return cudaSuccess;
}
cudaError_t cudaMemset(void *p, int v, size_t n) { std::lock_guard<std::mutex> g(g_mu); return memset_common(p, v, n, nullptr, "cudaMemset"); }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t s) {
    std::lock_guard<std::mutex> g(g_mu);
    Stream *st = as_stream(s, "cudaMemsetAsync");
    if (!st) return cudaErrorInvalidResourceHandle;
    return memset_common(p, v, n, st, "cudaMemsetAsync");
}

// ---- streams and events ----------------------------------------------------------------------------------------------------
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
    std::lock_guard<std::mutex> g(g_mu);
    Stream *st = new Stream{g_next_stream++};
    g_streams.insert(st);
    *s = reinterpret_cast<cudaStream_t>(st);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned flags, int priority) {
    if (priority > 0 || priority < -5) violation("cudaStreamCreateWithPriority: priority %d outside the device's range [-5, 0]", priority);
    return cudaStreamCreateWithFlags(s, flags);
}
cudaError_t cudaDeviceGetStreamPriorityRange(int *least, int *greatest) {
    if (least) *least = 0;
    if (greatest) *greatest = -5;
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) {
    std::lock_guard<std::mutex> g(g_mu);
    Stream *st = as_stream(s, "cudaStreamDestroy");
    if (!st || st == &g_default_stream) return cudaErrorInvalidResourceHandle;
    if (st->cap) violation("cudaStreamDestroy of capturing stream %d", st->id);
    g_streams.erase(st);
    delete st;
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) {
    std::lock_guard<std::mutex> g(g_mu);
    Stream *st = as_stream(s, "cudaStreamSynchronize");
    if (!st) return cudaErrorInvalidResourceHandle;
    if (st->cap) { violation("cudaStreamSynchronize on capturing stream %d", st->id); return cudaErrorStreamCaptureUnsupported; }
    return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) {
    std::lock_guard<std::mutex> g(g_mu);
    Event *ev = new Event();
    g_events.insert(ev);
    *e = reinterpret_cast<cudaEvent_t>(ev);
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t *e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventDestroy(cudaEvent_t e) {
    std::lock_guard<std::mutex> g(g_mu);
    Event *ev = as_event(e, "cudaEventDestroy");
    if (!ev) return cudaErrorInvalidResourceHandle;
    g_events.erase(ev);
    delete ev;
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
    std::lock_guard<std::mutex> g(g_mu);
    Event *ev = as_event(e, "cudaEventRecord");
    Stream *st = as_stream(s, "cudaEventRecord");
    if (!ev || !st) return cudaErrorInvalidResourceHandle;
    ev->recorded = true;
    ev->cap = st->cap;
    ev->node = st->cap ? st->last_node : -1;
    return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
    std::lock_guard<std::mutex> g(g_mu);
    Event *ev = as_event(e, "cudaStreamWaitEvent");
    Stream *st = as_stream(s, "cudaStreamWaitEvent");
    if (!ev || !st) return cudaErrorInvalidResourceHandle;
    if (!ev->recorded) return cudaSuccess;   // waiting on a never-recorded event is a no-op
    if (ev->cap) {
        if (st->cap && st->cap != ev->cap) { violation("cudaStreamWaitEvent: stream %d belongs to another capture", st->id); return cudaErrorStreamCaptureIsolation; }
        if (!st->cap) {  // fork: the waiting stream joins the capture
            st->cap = ev->cap;
            st->last_node = -1;
            ev->cap->streams.insert(st);
        }
        // an empty node carrying both dependencies
        Node n;
        if (st->last_node >= 0) n.deps.push_back(st->last_node);
        if (ev->node >= 0) n.deps.push_back(ev->node);
        st->cap->nodes.push_back(n);
        st->last_node = (int)st->cap->nodes.size() - 1;
    } else if (st->cap) {
        violation("cudaStreamWaitEvent: capturing stream %d waits on an event recorded outside the capture", st->id);
        return cudaErrorStreamCaptureIsolation;
    }
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e) {
    std::lock_guard<std::mutex> g(g_mu);
    Event *ev = as_event(e, "cudaEventSynchronize");
    if (!ev) return cudaErrorInvalidResourceHandle;
    if (ev->cap) { violation("cudaEventSynchronize on an event recorded inside a stream capture"); return cudaErrorStreamCaptureUnsupported; }
    return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    std::lock_guard<std::mutex> g(g_mu);
    Event *ea = as_event(a, "cudaEventElapsedTime"), *eb = as_event(b, "cudaEventElapsedTime");
    if (!ea || !eb) return cudaErrorInvalidResourceHandle;
    if (!ea->recorded || !eb->recorded) { violation("cudaEventElapsedTime on an event that was never recorded"); return cudaErrorInvalidResourceHandle; }
    if (ea->cap || eb->cap) { violation("cudaEventElapsedTime on events recorded inside a stream capture"); return cudaErrorInvalidResourceHandle; }
    *ms = 1.0f;
    return cudaSuccess;
}

// ---- stream capture and graphs ----------------------------------------------------------------------------------------------
cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode) {
    std::lock_guard<std::mutex> g(g_mu);
    Stream *st = as_stream(s, "cudaStreamBeginCapture");
    if (!st) return cudaErrorInvalidResourceHandle;
    if (st->cap) { violation("cudaStreamBeginCapture on a stream that is already capturing"); return cudaErrorIllegalState; }
    Capture *c = new Capture();
    c->origin = st;
    c->streams.insert(st);
    st->cap = c;
    st->last_node = -1;
    return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t *graph) {
    std::lock_guard<std::mutex> g(g_mu);
    Stream *st = as_stream(s, "cudaStreamEndCapture");
    *graph = nullptr;
    if (!st || !st->cap) { violation("cudaStreamEndCapture on a stream that is not capturing"); return cudaErrorIllegalState; }
    Capture *c = st->cap;
    if (c->origin != st) { violation("cudaStreamEndCapture on a forked stream (not the origin)"); return cudaErrorStreamCaptureUnmatched; }
    bool ok = true;
    for (void *vs : c->streams) {
        Stream *f = (Stream *)vs;
        if (f == st) continue;
        std::vector<char> seen(c->nodes.size(), 0);
        if (f->last_node >= 0 && !reaches(c, st->last_node, f->last_node, seen)) {
            violation("cudaStreamEndCapture: forked stream %d did not re-join the origin (cudaErrorStreamCaptureUnjoined)", f->id);
            ok = false;
        }
    }
    if (c->nccl_calls) { violation("%d NCCL call(s) were issued inside a stream capture", c->nccl_calls); ok = false; }
    const int nn = (int)c->nodes.size();
    for (void *vs : c->streams) { ((Stream *)vs)->cap = nullptr; ((Stream *)vs)->last_node = -1; }
    for (Event *ev : g_events)
        if (ev->cap == c) { ev->cap = nullptr; ev->recorded = false; ev->node = -1; }
    delete c;
    if (!ok) return cudaErrorStreamCaptureUnjoined;
    *graph = reinterpret_cast<cudaGraph_t>(new Graph{nn});
    return cudaSuccess;
}
cudaError_t cudaGraphGetNodes(cudaGraph_t gr, cudaGraphNode_t *, size_t *n) { *n = gr ? ((Graph *)gr)->nodes : 0; return cudaSuccess; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *ex, cudaGraph_t gr, unsigned long long) {
    if (!gr) return cudaErrorInvalidValue;
    *ex = reinterpret_cast<cudaGraphExec_t>(new Graph{((Graph *)gr)->nodes});
    return cudaSuccess;
}
cudaError_t cudaGraphLaunch(cudaGraphExec_t ex, cudaStream_t s) {
    std::lock_guard<std::mutex> g(g_mu);
    Stream *st = as_stream(s, "cudaGraphLaunch");
    if (!ex || !st) return cudaErrorInvalidResourceHandle;
    g_launch_count["<graph launch>"]++;
    stream_work(st);
    return cudaSuccess;
}
cudaError_t cudaGraphDestroy(cudaGraph_t gr) { delete (Graph *)gr; return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t ex) { delete (Graph *)ex; return cudaSuccess; }

// ---- IPC (all "ranks" of a dry run live in one process: a handle is the pointer) ------------------------------------------------
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
    std::lock_guard<std::mutex> g(g_mu);
    if (!check_device_range(p, 1, "cudaIpcGetMemHandle")) return cudaErrorInvalidValue;
    if (!g_allocs.count((uintptr_t)p)) { violation("cudaIpcGetMemHandle(%p): not the base of an allocation", p); return cudaErrorInvalidValue; }
    memset(h, 0, sizeof(*h));
    memcpy(h, &p, sizeof(p));
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
    std::lock_guard<std::mutex> g(g_mu);
    memcpy(p, &h, sizeof(*p));
    if (!check_device_range(*p, 1, "cudaIpcOpenMemHandle")) return cudaErrorInvalidValue;
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }

// ---- TMA descriptor encoding (driver entry point) ---------------------------------------------------------------------------
static CUresult fake_encode_tiled(CUtensorMap *map, CUtensorMapDataType dt, cuuint32_t rank, void *addr, const cuuint64_t *gdim, const cuuint64_t *gstr,
                                  const cuuint32_t *box, const cuuint32_t *estr, CUtensorMapInterleave il, CUtensorMapSwizzle sw, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill) {
    std::lock_guard<std::mutex> g(g_mu);
    bool ok = true;
    const size_t es = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT64 ? 8 : dt == CU_TENSOR_MAP_DATA_TYPE_UINT8 ? 1 : 4;
    if (rank < 1 || rank > 5) { violation("cuTensorMapEncodeTiled: rank %u", rank); ok = false; }
    if (((uintptr_t)addr & 15) != 0) { violation("cuTensorMapEncodeTiled: global address %p not 16-byte aligned", addr); ok = false; }
    if (!is_device_ptr(addr)) { violation("cuTensorMapEncodeTiled: %p is not device memory", addr); ok = false; }
    size_t extent = 0;
    for (cuuint32_t i = 0; ok && i < rank; ++i) {
        if (gdim[i] < 1 || gdim[i] > ((cuuint64_t)1 << 32)) { violation("cuTensorMapEncodeTiled: globalDim[%u] = %llu", i, (unsigned long long)gdim[i]); ok = false; }
        if (box[i] < 1 || box[i] > 256) { violation("cuTensorMapEncodeTiled: boxDim[%u] = %u (must be 1..256)", i, box[i]); ok = false; }
        if (estr[i] < 1 || estr[i] > 8) { violation("cuTensorMapEncodeTiled: elementStrides[%u] = %u", i, estr[i]); ok = false; }
        if (i > 0) {
            const cuuint64_t st = gstr[i - 1];
            if (st % 16 != 0 || st >= ((cuuint64_t)1 << 40)) { violation("cuTensorMapEncodeTiled: globalStrides[%u] = %llu (multiple of 16, < 2^40)", i - 1, (unsigned long long)st); ok = false; }
            extent = std::max<size_t>(extent, (size_t)st * (gdim[i] - 1));
        }
    }
    if (ok) {
        const size_t inner = (size_t)box[0] * es;
        if (il == CU_TENSOR_MAP_INTERLEAVE_NONE && inner % 16 != 0) { violation("cuTensorMapEncodeTiled: inner box of %zu bytes is not a multiple of 16", inner); ok = false; }
        const size_t lim = sw == CU_TENSOR_MAP_SWIZZLE_32B ? 32 : sw == CU_TENSOR_MAP_SWIZZLE_64B ? 64 : sw == CU_TENSOR_MAP_SWIZZLE_128B ? 128 : (size_t)-1;
        if (inner > lim) { violation("cuTensorMapEncodeTiled: inner box of %zu bytes exceeds the swizzle span %zu", inner, lim); ok = false; }
        // the tensor the descriptor spans must lie inside the allocation (TMA never reads outside the descriptor's extents)
        if (!check_device_range(addr, extent + (size_t)gdim[0] * es, "cuTensorMapEncodeTiled (tensor extent)")) ok = false;
    }
    if (!ok) return CUDA_ERROR_INVALID_VALUE;
    memset(map, 0xab, sizeof(*map));
    if (rank == 2 && dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT64 && sw == CU_TENSOR_MAP_SWIZZLE_NONE) {
        // layout read by the functional model of the launch hook: { const double *base; long d0, d1; int box0, box1; }
        struct { const double *base; long d0, d1; int box0, box1; } m = {(const double *)addr, (long)gdim[0], (long)gdim[1], (int)box[0], (int)box[1]};
        if (gstr[0] != gdim[0] * 8) { violation("cuTensorMapEncodeTiled: the functional model needs a dense 2-D view (stride %llu)", (unsigned long long)gstr[0]); return CUDA_ERROR_INVALID_VALUE; }
        memcpy(map, &m, sizeof(m));
    }
    return CUDA_SUCCESS;
}
cudaError_t cudaGetDriverEntryPoint(const char *sym, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *q) {
    if (q) *q = cudaDriverEntryPointSuccess;
    if (strcmp(sym, "cuTensorMapEncodeTiled") == 0) { *fn = (void *)&fake_encode_tiled; return cudaSuccess; }
    *fn = nullptr;
    if (q) *q = cudaDriverEntryPointSymbolNotFound;
    return cudaErrorInvalidValue;
}

// ---- NCCL (the library binds it with dlopen("libnccl.so.2"); the test puts this object on the search path under that name) ----
struct FakeComm { int n, rank; };
int ncclGetUniqueId(void *id) { memset(id, 7, 128); return 0; }
struct FakeId { char b[128]; };
int ncclCommInitRank(void **comm, int n, FakeId, int rank) { *comm = new FakeComm{n, rank}; return 0; }
static int nccl_common(const void *src, void *dst, size_t nsrc, size_t ndst, cudaStream_t s, const char *what) {
    std::lock_guard<std::mutex> g(g_mu);
    Stream *st = as_stream(s, what);
    if (!st) return 1;
    bool ok = check_device_range(src, nsrc, what) & check_device_range(dst, ndst, what);
    if (st->cap) st->cap->nccl_calls++;
    g_launch_count[std::string("<") + what + ">"]++;
    stream_work(st);
    return ok ? 0 : 4;
}
int ncclAllReduce(const void *src, void *dst, size_t count, int dtype, int, void *, cudaStream_t s) {
    const size_t es = dtype == 8 ? 8 : 4;
    return nccl_common(src, dst, count * es, count * es, s, "ncclAllReduce");
}
int ncclAllGather(const void *src, void *dst, size_t count, int dtype, void *comm, cudaStream_t s) {
    const size_t es = dtype == 8 ? 8 : 4;
    return nccl_common(src, dst, count * es, count * es * ((FakeComm *)comm)->n, s, "ncclAllGather");
}
int ncclCommDestroy(void *c) { delete (FakeComm *)c; return 0; }
const char *ncclGetErrorString(int) { return "fake nccl error"; }
}  // extern "C"
