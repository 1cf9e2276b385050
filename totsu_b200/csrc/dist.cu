// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch (north_star).
// The reference has no collectives at all (single device, totsu_f32cuda/src/cuda_mgr.rs:37-39).  NCCL is bound
// at run time with dlopen so the library has no link-time dependency on a particular libnccl: inside a
// torchrun-launched process this resolves to the libnccl.so.2 torch already loaded.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>

namespace tb {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
};
static NcclApi g_nccl;

static void load_nccl() {
    if (g_nccl.lib) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) fail(TB_ERR_NCCL, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* s) {
        void* p = dlsym(g_nccl.lib, s);
        if (!p) fail(TB_ERR_NCCL, std::string("libnccl is missing symbol ") + s);
        return p;
    };
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
}

#define TB_NCCL(expr)                                                                                          \
    do {                                                                                                       \
        ncclResult_t _r = (expr);                                                                              \
        if (_r != ncclSuccess) fail(TB_ERR_NCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(_r));       \
    } while (0)

void dist_allreduce_sum(void* buf, size_t count, int dtype) {
    Context& c = ctx();
    if (c.world <= 1 || count == 0) return;
    TB_NCCL(g_nccl.AllReduce(buf, buf, count, dtype == TB_F32 ? ncclFloat32 : ncclFloat64, ncclSum, (ncclComm_t)c.nccl_comm, c.stream));
    count_launch();
}

void dist_allgather_inplace(void* base, size_t count_per_rank, int dtype) {
    Context& c = ctx();
    if (c.world <= 1 || count_per_rank == 0) return;
    const size_t es = dtype == TB_F32 ? 4 : 8;
    const char* send = reinterpret_cast<const char*>(base) + (size_t)c.rank * count_per_rank * es;
    TB_NCCL(g_nccl.AllGather(send, base, count_per_rank, dtype == TB_F32 ? ncclFloat32 : ncclFloat64, (ncclComm_t)c.nccl_comm, c.stream));
    count_launch();
}

}  // namespace tb

using namespace tb;
extern "C" {

int tb_dist_unique_id(void* id_out) {
    return api([&] {
        static_assert(TB_NCCL_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
        load_nccl();
        ncclUniqueId id;
        TB_NCCL(g_nccl.GetUniqueId(&id));
        std::memcpy(id_out, &id, sizeof(id));
    });
}

int tb_dist_init(int rank, int world, const void* id_bytes) {
    return api([&] {
        require_init();
        Context& c = ctx();
        TB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
        if (c.nccl_comm) fail(TB_ERR_STATE, "tb_dist_init called twice");
        if (world == 1) { c.rank = 0; c.world = 1; return; }
        load_nccl();
        ncclUniqueId id;
        std::memcpy(&id, id_bytes, sizeof(id));
        ncclComm_t comm;
        TB_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
        c.nccl_comm = comm;
        c.rank = rank;
        c.world = world;
    });
}

int tb_dist_finalize(void) {
    return api([&] {
        Context& c = ctx();
        if (c.nccl_comm) {
            cudaStreamSynchronize(c.stream);
            g_nccl.CommDestroy((ncclComm_t)c.nccl_comm);
            c.nccl_comm = nullptr;
        }
        c.rank = 0;
        c.world = 1;
    });
}

int tb_dist_info(int* rank, int* world) {
    return api([&] {
        *rank = ctx().rank;
        *world = ctx().world;
    });
}
}
