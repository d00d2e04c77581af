// Multi-GPU exchange: replicated tree, ONE all-reduce per training iteration over the
// statistics of `current`.  libnccl.so.2 is dlopen'ed on first use so that
// single-GPU users need no NCCL at all; the handful of entry points used are declared
// here by hand (NCCL's stable C ABI).
#ifndef SDT_HOSTEMU
#include <dlfcn.h>
typedef struct { char internal[128]; } sdt_ncclUniqueId;
typedef int (*fn_ncclGetUniqueId)(sdt_ncclUniqueId*);
typedef int (*fn_ncclCommInitRank)(void**, int, sdt_ncclUniqueId, int);
typedef int (*fn_ncclCommDestroy)(void*);
typedef int (*fn_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_ncclGroup)(void);
typedef const char* (*fn_ncclGetErrorString)(int);
static struct {
    void* lib; fn_ncclGetUniqueId get_id; fn_ncclCommInitRank init; fn_ncclCommDestroy destroy; fn_ncclAllReduce allreduce;
    fn_ncclGroup group_start, group_end; fn_ncclGetErrorString errstr;
} g_nccl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

static int sdt_nccl_load(sdt_handle h) {
    if (g_nccl.lib) return SDT_OK;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return sdt_fail(h, SDT_ERR_NCCL, std::string("dlopen libnccl.so.2 failed: ") + dlerror());
    g_nccl.get_id = (fn_ncclGetUniqueId)dlsym(lib, "ncclGetUniqueId");
    g_nccl.init = (fn_ncclCommInitRank)dlsym(lib, "ncclCommInitRank");
    g_nccl.destroy = (fn_ncclCommDestroy)dlsym(lib, "ncclCommDestroy");
    g_nccl.allreduce = (fn_ncclAllReduce)dlsym(lib, "ncclAllReduce");
    g_nccl.group_start = (fn_ncclGroup)dlsym(lib, "ncclGroupStart");
    g_nccl.group_end = (fn_ncclGroup)dlsym(lib, "ncclGroupEnd");
    g_nccl.errstr = (fn_ncclGetErrorString)dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.get_id || !g_nccl.init || !g_nccl.destroy || !g_nccl.allreduce || !g_nccl.group_start || !g_nccl.group_end)
        return sdt_fail(h, SDT_ERR_NCCL, "libnccl.so.2 lacks a required symbol");
    g_nccl.lib = lib;
    return SDT_OK;
}
static int sdt_nccl_fail(sdt_handle h, const char* what, int rc) {
    return sdt_fail(h, SDT_ERR_NCCL, std::string(what) + ": " + (g_nccl.errstr ? g_nccl.errstr(rc) : "nccl error"));
}
static void sdt_nccl_destroy(sdt_handle h) {
    if (h->nccl_comm && g_nccl.destroy) g_nccl.destroy(h->nccl_comm);
    h->nccl_comm = nullptr;
}
extern "C" int sdt_comm_unique_id(void* id128) {
    if (!id128) return SDT_ERR_INVALID;
    SDT_TRY(sdt_nccl_load(nullptr));
    sdt_ncclUniqueId id;
    const int rc = g_nccl.get_id(&id);
    if (rc != 0) return sdt_nccl_fail(nullptr, "ncclGetUniqueId", rc);
    memcpy(id128, &id, 128);
    return SDT_OK;
}
extern "C" int sdt_comm_init(sdt_handle h, const void* id128, int32_t rank, int32_t nranks) {
    if (!h || !id128) return SDT_ERR_INVALID;
    SDT_ENTER(h);
    SDT_CHECK(h, nranks >= 1 && rank >= 0 && rank < nranks, SDT_ERR_INVALID, "sdt_comm_init: bad rank / nranks");
    SDT_TRY(sdt_nccl_load(h));
    SDT_CUDA(h, cudaSetDevice(h->cfg.device));
    sdt_nccl_destroy(h);
    sdt_ncclUniqueId id;
    memcpy(&id, id128, 128);
    const int rc = g_nccl.init(&h->nccl_comm, nranks, id, rank);
    if (rc != 0) return sdt_nccl_fail(h, "ncclCommInitRank", rc);
    h->rank = rank; h->nranks = nranks;
    return SDT_OK;
}
// ncclAllReduce(sum, fp32) over the LEAF statistics of current: [quadtree energies | spatial
// counts]; the interior sums are rebuilt afterwards from the reduced leaves, so every rank
// ends with bit-identical buffers and the deterministic refine needs no broadcast.
extern "C" int sdt_allreduce(sdt_handle h, sdt_stream stream) {
    SDT_ENTER(h);
    SDT_CHECK(h, h->nccl_comm, SDT_ERR_STATE, "sdt_allreduce: call sdt_comm_init first");
    cudaStream_t st = (cudaStream_t)stream;
    // element counts (identical on every rank): the sizes the last refine / upload reported -- its non-blocking header
    // read-back has landed long before an iteration's passes are over; only if it has not, wait for it
    (void)tree_view(h);
    if (h->hdr_pending) { DevHeader H; SDT_TRY(sdt_read_header(h, H)); }
    const uint32_t n_quad = h->n_quad_known, n_kd = h->kd_nodes_known;
    sdt_order_after_last(h, st);
    int rc = g_nccl.group_start();
    if (rc == 0) rc = g_nccl.allreduce(h->q_ecur, h->q_ecur, n_quad, /*ncclFloat32*/ 7, /*ncclSum*/ 0, h->nccl_comm, st);
    if (rc == 0) rc = g_nccl.allreduce(h->kd_count, h->kd_count, n_kd, 7, 0, h->nccl_comm, st);
    const int rc2 = g_nccl.group_end();
    if (rc != 0 || rc2 != 0) return sdt_nccl_fail(h, "ncclAllReduce", rc ? rc : rc2);
    h->last_stream = st;
    h->stats_complete = false; h->kd_complete = false;                // interiors are recomputed from the reduced leaves
    h->splat_bound_valid = false;                                     // counts of other ranks' records arrived
    return SDT_OK;
}
#else
static void sdt_nccl_destroy(sdt_handle) {}
extern "C" int sdt_comm_unique_id(void*) { return sdt_fail(nullptr, SDT_ERR_NCCL, "NCCL is not part of the host emulation"); }
extern "C" int sdt_comm_init(sdt_handle h, const void*, int32_t, int32_t) { return sdt_fail(h, SDT_ERR_NCCL, "NCCL is not part of the host emulation"); }
extern "C" int sdt_allreduce(sdt_handle h, sdt_stream) { return sdt_fail(h, SDT_ERR_NCCL, "NCCL is not part of the host emulation"); }
#endif
