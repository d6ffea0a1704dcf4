// comm.cu — multi-GPU plumbing for row-sharded tensor parallelism: one process per
// GPU, NCCL over NVLink 5 / NVSwitch, allgather at the layer boundaries where the row shards meet.
// The reference has no communication backend; nothing is ported.
// NCCL is resolved with dlopen so the library loads (and single-GPU paths run) without it.
#include <dlfcn.h>

#include <nccl.h>

#include "common.cuh"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_world = 1;

int load_nccl() {
    if (g_nccl.handle) return GGB_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
        ggb_set_error("cannot dlopen libnccl.so.2: %s", dlerror());
        return GGB_ERR_NCCL;
    }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.handle, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.handle, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(g_nccl.handle, "ncclAllGather");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.handle, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather) {
        ggb_set_error("libnccl is missing required symbols");
        return GGB_ERR_NCCL;
    }
    return GGB_OK;
}

#define GGB_NCCL(call)                                                                                    \
    do {                                                                                                  \
        ncclResult_t r__ = (call);                                                                        \
        if (r__ != ncclSuccess) {                                                                         \
            ggb_set_error("%s -> %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error"); \
            return GGB_ERR_NCCL;                                                                          \
        }                                                                                                 \
    } while (0)

} // namespace

extern "C" {

int ggb_comm_unique_id(void *id128) {
    int rc = load_nccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    GGB_NCCL(g_nccl.GetUniqueId((ncclUniqueId *)id128));
    return GGB_OK;
}

int ggb_comm_init(const void *id128, int rank, int world) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
    g_rank = rank;
    g_world = world;
    if (world == 1) return GGB_OK;
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    GGB_NCCL(g_nccl.CommInitRank(&g_comm, world, id, rank));
    return GGB_OK;
}

int ggb_comm_destroy(void) {
    if (g_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
    g_rank = 0;
    g_world = 1;
    return GGB_OK;
}

int ggb_comm_rank(void) { return g_rank; }
int ggb_comm_size(void) { return g_world; }

int ggb_op_allgather(const void *send, void *recv, size_t bytes) {
    GGB_REQUIRE_INIT();
    if (g_world == 1) {
        if (send != recv) GGB_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, ggb_cuda_stream()));
        return GGB_OK;
    }
    GGB_CHECK_ARG(g_comm != nullptr, "ggb_comm_init has not been called");
    GGB_NCCL(g_nccl.AllGather(send, recv, bytes, ncclInt8, g_comm, ggb_cuda_stream()));
    ggb_count_launch();
    return GGB_OK;
}

} // extern "C"
