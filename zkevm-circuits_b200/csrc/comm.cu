// comm.cu -- NCCL communicator of a context (one process per GPU) for the multi-GPU create_proof.
//
// The proving session stays replicated (every rank runs the same host code on the same inputs and therefore produces
// the same transcript and the same proof bytes); only the two heavy, embarrassingly parallel stages are dealt across
// ranks: commitment batches (column i is committed by rank i mod P) and the quotient's coset parts (part j
// by rank j mod P).  Results are exchanged with ONE all-reduce each over NVLink: every rank writes its share into a
// zero-initialised buffer, so a u64 sum of the disjoint supports is an exact gather (no 256-bit NCCL type exists).
// NCCL is resolved at run time (dlopen of libnccl.so.2: inside a torch process that is torch's bundled copy).
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

namespace zkb {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    if (api.handle) return &api;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("cannot load libnccl.so.2: %s", dlerror()); return nullptr; }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) { set_error("libnccl.so.2 lacks required symbols"); return nullptr; }
    api.handle = h;
    return &api;
}

#define ZKB_NCCL(api, expr)                                                                                           \
    do {                                                                                                              \
        ncclResult_t _r = (expr);                                                                                     \
        if (_r != ncclSuccess) {                                                                                      \
            zkb::set_error("%s:%d NCCL: %s", __FILE__, __LINE__, (api)->GetErrorString ? (api)->GetErrorString(_r) : "error"); \
            return ZKB_ERR_CUDA;                                                                                      \
        }                                                                                                             \
    } while (0)

// in-place sum of `count` u64 words on the device (disjoint supports -> exact gather)
int32_t comm_allreduce_u64(zkb_ctx *ctx, void *dev_buf, size_t count, cudaStream_t st) {
    if (ctx->nranks <= 1) return ZKB_OK;
    NcclApi *api = nccl_api();
    if (!api || !ctx->nccl_comm) { set_error("communicator not initialised"); return ZKB_ERR_STATE; }
    ZKB_NCCL(api, api->AllReduce(dev_buf, dev_buf, count, ncclUint64, ncclSum, (ncclComm_t)ctx->nccl_comm, st));
    return ZKB_OK;
}

}  // namespace zkb
using namespace zkb;

extern "C" int32_t zkb_comm_unique_id(uint8_t out[128]) {
    ZKB_ARG(out != nullptr);
    NcclApi *api = nccl_api();
    if (!api) return ZKB_ERR_CUDA;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ZKB_NCCL(api, api->GetUniqueId(&id));
    memcpy(out, &id, 128);
    return ZKB_OK;
}

extern "C" int32_t zkb_comm_init(zkb_ctx *ctx, const uint8_t unique_id[128], int32_t rank, int32_t nranks) {
    ZKB_ARG(ctx && unique_id && nranks >= 1 && rank >= 0 && rank < nranks);
    if (ctx->nccl_comm) { set_error("communicator already initialised"); return ZKB_ERR_STATE; }
    ZKB_CUDA(cudaSetDevice(ctx->device));
    NcclApi *api = nccl_api();
    if (!api) return ZKB_ERR_CUDA;
    ncclUniqueId id;
    memcpy(&id, unique_id, 128);
    ncclComm_t comm = nullptr;
    ZKB_NCCL(api, api->CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return ZKB_OK;
}

extern "C" int32_t zkb_comm_destroy(zkb_ctx *ctx) {
    ZKB_ARG(ctx);
    if (ctx->nccl_comm) {
        NcclApi *api = nccl_api();
        cudaStreamSynchronize(ctx->stream);
        if (api) api->CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->rank = 0;
    ctx->nranks = 1;
    return ZKB_OK;
}
