// Multi-GPU plumbing: one process per GPU, slab sharding along the last mode.
// NCCL is bound lazily with dlopen so that single-GPU users (and the Julia extension) carry no
// NCCL link dependency; inside a torch process the already-loaded bundled libnccl.so.2 is reused.
#include "common.cuh"
#include <dlfcn.h>

namespace itcpd {

struct NcclId { char internal[128]; };
typedef int (*fn_get_id)(NcclId *);
typedef int (*fn_init_rank)(void **, int, NcclId, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_allgather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*fn_destroy)(void *);
typedef const char *(*fn_errstr)(int);

struct NcclApi {
    void *lib = nullptr;
    fn_get_id get_id = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_allgather allgather = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    if (api.lib) return &api;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) { set_error("dlopen(libnccl.so.2) failed: %s", dlerror()); return nullptr; }
    api.get_id = (fn_get_id)dlsym(api.lib, "ncclGetUniqueId");
    api.init_rank = (fn_init_rank)dlsym(api.lib, "ncclCommInitRank");
    api.allreduce = (fn_allreduce)dlsym(api.lib, "ncclAllReduce");
    api.allgather = (fn_allgather)dlsym(api.lib, "ncclAllGather");
    api.destroy = (fn_destroy)dlsym(api.lib, "ncclCommDestroy");
    api.errstr = (fn_errstr)dlsym(api.lib, "ncclGetErrorString");
    if (!api.get_id || !api.init_rank || !api.allreduce || !api.allgather || !api.destroy) {
        set_error("libnccl is missing required symbols");
        api.lib = nullptr;
        return nullptr;
    }
    return &api;
}

struct Comm {
    void *comm = nullptr;
    int nranks = 1, rank = 0;
};

#define NCCL_TRY(expr)                                                                               \
    do {                                                                                             \
        int _r = (expr);                                                                             \
        if (_r != 0) {                                                                               \
            NcclApi *_a = nccl_api();                                                                \
            set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, (_a && _a->errstr) ? _a->errstr(_r) : "nccl error"); \
            return ITCPD_ERR_COMM;                                                                   \
        }                                                                                            \
    } while (0)

bool comm_active(const itcpd_ctx *c) { return c->comm != nullptr && c->comm->nranks > 1; }

int comm_allreduce_sum(itcpd_ctx *c, double *buf, int64_t n) {
    if (!comm_active(c)) return ITCPD_OK;
    if (peer_graph_active(c) && n <= c->peer_small_doubles) return peer_allreduce_small(c, buf, n);  // NCCL-free (capturable) sweeps
    NcclApi *a = nccl_api();
    if (!a) return ITCPD_ERR_COMM;
    NCCL_TRY(a->allreduce(buf, buf, (size_t)n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->comm->comm, c->stream));
    return ITCPD_OK;
}

int comm_allgather(itcpd_ctx *c, const double *send, double *recv, int64_t n_per_rank) {
    NcclApi *a = nccl_api();
    if (!a) return ITCPD_ERR_COMM;
    NCCL_TRY(a->allgather(send, recv, (size_t)n_per_rank, 8, c->comm->comm, c->stream));
    return ITCPD_OK;
}

int comm_rank(const itcpd_ctx *c) { return c->comm ? c->comm->rank : 0; }
int comm_size(const itcpd_ctx *c) { return c->comm ? c->comm->nranks : 1; }

}  // namespace itcpd

using namespace itcpd;

extern "C" int itcpd_comm_unique_id(void *out128) {
    NcclApi *a = nccl_api();
    if (!a) return ITCPD_ERR_COMM;
    NcclId id;
    NCCL_TRY(a->get_id(&id));
    memcpy(out128, &id, 128);
    return ITCPD_OK;
}

extern "C" int itcpd_comm_init(itcpd_ctx *c, int nranks, int rank, const void *id128) {
    ARG_CHECK(c && nranks >= 1 && rank >= 0 && rank < nranks && id128, "bad comm arguments");
    NcclApi *a = nccl_api();
    if (!a) return ITCPD_ERR_COMM;
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->comm) itcpd_comm_destroy(c);
    c->comm = new Comm();
    c->comm->nranks = nranks;
    c->comm->rank = rank;
    NcclId id;
    memcpy(&id, id128, 128);
    NCCL_TRY(a->init_rank(&c->comm->comm, nranks, id, rank));
    return ITCPD_OK;
}

extern "C" int itcpd_comm_destroy(itcpd_ctx *c) {
    if (!c || !c->comm) return ITCPD_OK;
    NcclApi *a = nccl_api();
    if (a && c->comm->comm) a->destroy(c->comm->comm);
    delete c->comm;
    c->comm = nullptr;
    return ITCPD_OK;
}
