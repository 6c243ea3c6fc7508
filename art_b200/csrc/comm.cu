// The one collective on the data path: an int32 sum over the ranks that share a frame (row bands, develop.cu / shrink.cu).
//
// ncclAllReduce on the context's stream.  NCCL is loaded at run time (dlopen libnccl.so.2: inside a torch process that is the copy
// torch already mapped, otherwise the system's), so the library keeps its CUDA-runtime-only link line and a single-GPU user never
// touches NCCL.  Only the handful of entry points used here are declared; their signatures are NCCL's public C API (nccl.h).
#include <dlfcn.h>

#include <mutex>

#include "ctx.h"

namespace {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* ncclComm_t;
enum { NCCL_INT32 = 2, NCCL_SUM = 0 };      // ncclDataType_t ncclInt32, ncclRedOp_t ncclSum

struct Nccl {
    void* so = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};
Nccl g_nccl;
std::mutex g_nccl_mu;

bool nccl_load()
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.AllReduce) return true;
    void* so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!so) { g_nccl.err = dlerror(); return false; }
    auto sym = [&](const char* n) { void* p = dlsym(so, n); if (!p) g_nccl.err = std::string("missing NCCL symbol ") + n; return p; };
    g_nccl.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, NcclUniqueId, int))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))sym("ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    void* ar = sym("ncclAllReduce");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.GetErrorString || !ar) return false;
    g_nccl.so = so;
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))ar;
    return true;
}

int nccl_allreduce(void* user, int* d_buf, size_t count, void* stream)
{
    art_hp_ctx* ctx = (art_hp_ctx*)user;
    const int rc = g_nccl.AllReduce(d_buf, d_buf, count, NCCL_INT32, NCCL_SUM, (ncclComm_t)ctx->nccl_comm, (cudaStream_t)stream);
    if (rc) return ctx->fail(ART_HP_ERR_CUDA, "ncclAllReduce failed: %s", g_nccl.GetErrorString(rc));
    return ART_HP_OK;
}

}  // namespace

int art_allreduce_i32(art_hp_ctx* ctx, int* d_buf, size_t count)
{
    if (!ctx->allreduce) return ctx->fail(ART_HP_ERR_INVALID, "a row band needs the collective: art_hp_comm_init or art_hp_set_allreduce first");
    const int rc = ctx->allreduce(ctx->allreduce_user, d_buf, count, (void*)ctx->stream);
    if (rc && ctx->err.empty()) return ctx->fail(ART_HP_ERR_CUDA, "the all-reduce hook returned %d", rc);
    return rc;
}

extern "C" {

int art_hp_set_allreduce(art_hp_ctx* ctx, art_hp_allreduce_fn fn, void* user)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (ctx->nccl_comm) return ctx->fail(ART_HP_ERR_INVALID, "the context owns an NCCL communicator: art_hp_comm_destroy first");
    ctx->allreduce = fn;
    ctx->allreduce_user = user;
    return ART_HP_OK;
}

int art_hp_comm_unique_id(unsigned char id[128])
{
    if (!id || !nccl_load()) return ART_HP_ERR_UNSUPPORTED;
    NcclUniqueId u;
    if (g_nccl.GetUniqueId(&u)) return ART_HP_ERR_CUDA;
    std::memcpy(id, u.internal, 128);
    return ART_HP_OK;
}

int art_hp_comm_init(art_hp_ctx* ctx, const unsigned char id[128], int rank, int nranks)
{
    if (!ctx || !id) return ART_HP_ERR_INVALID;
    if (nranks < 1 || rank < 0 || rank >= nranks) return ctx->fail(ART_HP_ERR_INVALID, "rank %d of %d", rank, nranks);
    if (ctx->nccl_comm) return ctx->fail(ART_HP_ERR_INVALID, "the context already has a communicator");
    if (!nccl_load()) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "NCCL is not available: %s", g_nccl.err.c_str());
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    NcclUniqueId u;
    std::memcpy(u.internal, id, 128);
    ncclComm_t comm = nullptr;
    const int rc = g_nccl.CommInitRank(&comm, nranks, u, rank);
    if (rc) return ctx->fail(ART_HP_ERR_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(rc));
    ctx->nccl_comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_size = nranks;
    ctx->allreduce = nccl_allreduce;
    ctx->allreduce_user = ctx;
    return ART_HP_OK;
}

int art_hp_comm_destroy(art_hp_ctx* ctx)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (ctx->nccl_comm) {
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
        ctx->allreduce = nullptr;
        ctx->allreduce_user = nullptr;
        ctx->comm_rank = 0;
        ctx->comm_size = 1;
    }
    return ART_HP_OK;
}

}  // extern "C"
