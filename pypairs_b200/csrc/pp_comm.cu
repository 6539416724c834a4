// pp_comm.cu -- the one collective layer of the library: NCCL over NVLink / NVSwitch, one process per GPU.
//
// The reference has no distributed code at all; both hot paths shard without a data-path
// exchange (sandbag: gene-pair tiles, cyclone: cells).  What crosses GPUs is
//   sandbag  prep: all-reduce(max) of the expressed-gene flags and of the per-gene rank widths, all-gather of the
//            bit planes (every rank ranks + bit-slices only its own block of cells);
//            result: all-gather of the compacted marker keys
//   cyclone  result: all-gather of the score columns
// all enqueued on the ctx stream between the kernels of the same call, so the collectives are inside the timed step.
//
// NCCL is not linked: the entry points are resolved with dlopen/dlsym from the libnccl.so.2 the process already
// holds (the one torch bundles, nvidia/nccl/lib) or from an explicit path (pp_nccl_load).  The communicator is
// bootstrapped by the host: rank 0 calls pp_comm_unique_id, the 128 bytes travel over whatever channel the host
// has (torch.distributed in pypairs_b200/_dist.py), every rank calls pp_comm_init.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "pp_common.cuh"

namespace {

struct NcclApi {
    void *handle = nullptr;
    int version = 0;
    ppn::ncclResult_t (*GetVersion)(int *) = nullptr;
    ppn::ncclResult_t (*GetUniqueId)(ppn::ncclUniqueId *) = nullptr;
    ppn::ncclResult_t (*CommInitRank)(ppn::ncclComm_t *, int, ppn::ncclUniqueId, int) = nullptr;
    ppn::ncclResult_t (*CommDestroy)(ppn::ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ppn::ncclResult_t) = nullptr;
    ppn::ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ppn::ncclComm_t, cudaStream_t) = nullptr;
    ppn::ncclResult_t (*AllGather)(const void *, void *, size_t, int, ppn::ncclComm_t, cudaStream_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

int load_locked(const char *path) {
    if (g_nccl.handle) return PP_OK;
    const char *cands[3] = {path, "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (int i = 0; i < 3 && !h; ++i)
        if (cands[i] && cands[i][0]) h = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) return pp_fail(nullptr, PP_ENCCL, "pp_nccl_load: cannot dlopen libnccl.so.2 (%s)", dlerror());
    NcclApi a;
    a.handle = h;
#define PP_SYM(field, name)                                                                  \
    *(void **)(&a.field) = dlsym(h, name);                                                   \
    if (!a.field) return pp_fail(nullptr, PP_ENCCL, "pp_nccl_load: symbol %s not found", name);
    PP_SYM(GetVersion, "ncclGetVersion")
    PP_SYM(GetUniqueId, "ncclGetUniqueId")
    PP_SYM(CommInitRank, "ncclCommInitRank")
    PP_SYM(CommDestroy, "ncclCommDestroy")
    PP_SYM(GetErrorString, "ncclGetErrorString")
    PP_SYM(AllReduce, "ncclAllReduce")
    PP_SYM(AllGather, "ncclAllGather")
#undef PP_SYM
    if (a.GetVersion(&a.version) != 0) a.version = 0;
    g_nccl = a;
    return PP_OK;
}

int nccl_fail(pp_ctx *ctx, const char *what, ppn::ncclResult_t r) {
    return pp_fail(ctx, PP_ENCCL, "%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
}

}  // namespace

PP_API int pp_nccl_load(const char *path) {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    return load_locked(path);
}

PP_API int pp_nccl_version(void) {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    return g_nccl.handle ? g_nccl.version : 0;
}

PP_API int pp_comm_unique_id(uint8_t *out_id) {
    if (!out_id) return pp_fail(nullptr, PP_EINVAL, "pp_comm_unique_id: out_id is NULL");
    {
        std::lock_guard<std::mutex> lock(g_nccl_mutex);
        const int rc = load_locked(nullptr);
        if (rc) return rc;
    }
    ppn::ncclUniqueId id;
    const ppn::ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != 0) return nccl_fail(nullptr, "ncclGetUniqueId", r);
    memcpy(out_id, id.internal, PP_COMM_ID_BYTES);
    return PP_OK;
}

PP_API int pp_comm_init(pp_ctx *ctx, const uint8_t *id, int32_t rank, int32_t world) {
    if (!ctx || !id) return pp_fail(ctx, PP_EINVAL, "pp_comm_init: NULL argument");
    if (world < 1 || rank < 0 || rank >= world) return pp_fail(ctx, PP_EINVAL, "pp_comm_init: bad rank %d / %d", rank, world);
    if (ctx->comm) return pp_fail(ctx, PP_EINVAL, "pp_comm_init: the context already has a communicator");
    {
        std::lock_guard<std::mutex> lock(g_nccl_mutex);
        const int rc = load_locked(nullptr);
        if (rc) return rc;
    }
    PP_CUDA(cudaSetDevice(ctx->device));
    ppn::ncclUniqueId uid;
    memcpy(uid.internal, id, PP_COMM_ID_BYTES);
    ppn::ncclComm_t comm = nullptr;
    const ppn::ncclResult_t r = g_nccl.CommInitRank(&comm, world, uid, rank);
    if (r != 0) return nccl_fail(ctx, "ncclCommInitRank", r);
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    return PP_OK;
}

PP_API int pp_comm_destroy(pp_ctx *ctx) {
    if (!ctx) return pp_fail(nullptr, PP_EINVAL, "ctx is NULL");
    if (ctx->comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy((ppn::ncclComm_t)ctx->comm);
    }
    ctx->comm = nullptr;
    ctx->comm_rank = 0;
    ctx->comm_world = 0;
    return PP_OK;
}

PP_API int pp_comm_info(pp_ctx *ctx, int32_t *rank, int32_t *world) {
    if (!ctx) return pp_fail(nullptr, PP_EINVAL, "ctx is NULL");
    if (rank) *rank = ctx->comm ? ctx->comm_rank : 0;
    if (world) *world = ctx->comm ? ctx->comm_world : 0;
    return PP_OK;
}

// ---- collectives on the ctx stream (internal) -------------------------------------------------------------------
int pp_allreduce_max(pp_ctx *ctx, void *d_buf, size_t count, int nccl_dtype) {
    if (!ctx->comm) return pp_fail(ctx, PP_ENCCL, "no communicator (pp_comm_init)");
    if (count == 0) return PP_OK;
    const ppn::ncclResult_t r =
        g_nccl.AllReduce(d_buf, d_buf, count, nccl_dtype, ppn::ncclMax, (ppn::ncclComm_t)ctx->comm, ctx->stream);
    if (r != 0) return nccl_fail(ctx, "ncclAllReduce", r);
    return PP_OK;
}

int pp_allgather_bytes(pp_ctx *ctx, const void *d_send, void *d_recv, size_t bytes_per_rank) {
    if (!ctx->comm) return pp_fail(ctx, PP_ENCCL, "no communicator (pp_comm_init)");
    if (bytes_per_rank == 0) return PP_OK;
    ppn::ncclResult_t r;
    if (bytes_per_rank % 8 == 0)  // wider elements: same bytes, fewer of them
        r = g_nccl.AllGather(d_send, d_recv, bytes_per_rank / 8, ppn::ncclUint64, (ppn::ncclComm_t)ctx->comm, ctx->stream);
    else
        r = g_nccl.AllGather(d_send, d_recv, bytes_per_rank, ppn::ncclUint8, (ppn::ncclComm_t)ctx->comm, ctx->stream);
    if (r != 0) return nccl_fail(ctx, "ncclAllGather", r);
    return PP_OK;
}
