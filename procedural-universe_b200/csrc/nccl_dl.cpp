// Per-step position exchange: ncclAllGather of float4 {x, y, z, G m} over NVLink / NVSwitch.
//
// The reference has no communication layer (single process, shared memory: every worker thread
// reads all particles, BruteForceCPU.cpp:29-35).  Here each GPU integrates N/world bodies and needs
// all N positions for the next force pass, so the one exchange step per time step is an in-place
// all-gather of 16 B per body, issued on the handle's stream right behind the kick-drift kernel.
//
// NCCL is resolved at run time (dlopen) so that the library has no link-time dependency on a
// particular libnccl: inside a torch process this picks up the libnccl.so.2 torch already loaded.
#include <dlfcn.h>
#include <cstring>

#include "nb_internal.h"

namespace
{

typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
enum { kNcclSuccess = 0, kNcclChar = 0 };

struct NcclApi
{
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

NcclApi& api()
{
    static NcclApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names)
    {
        a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return a;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.lib, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.lib, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.lib, "ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.lib, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString;
    return a;
}

int fail(const char* what, int rc)
{
    nb::set_error("%s: %s", what, api().GetErrorString ? api().GetErrorString(rc) : "NCCL error");
    return NB_ERR_NCCL;
}

}  // namespace

namespace nb
{

int comm_unique_id(uint8_t id[128])
{
    if (id == nullptr) { set_error("nb_comm_unique_id: null argument"); return NB_ERR_ARG; }
    NcclApi& a = api();
    if (!a.ok) { set_error("libnccl.so.2 not found or incomplete (%s)", dlerror() ? dlerror() : "symbols missing"); return NB_ERR_NCCL; }
    NcclUniqueId u;
    const int rc = a.GetUniqueId(&u);
    if (rc != kNcclSuccess) return fail("ncclGetUniqueId", rc);
    std::memcpy(id, u.internal, 128);
    return NB_OK;
}

int comm_init(nb_sim* h, const uint8_t id[128])
{
    NcclApi& a = api();
    if (!a.ok) { set_error("libnccl.so.2 not found or incomplete"); return NB_ERR_NCCL; }
    if (h->n > 0 && h->n % (size_t)h->cfg.world != 0)
    {
        set_error("nb_comm_init: the NCCL exchange needs the body count (%zu) to be divisible by the number of ranks (%d)", h->n, h->cfg.world);
        return NB_ERR_ARG;
    }
    if (h->nccl_comm != nullptr) comm_destroy(h);
    NcclUniqueId u;
    std::memcpy(u.internal, id, 128);
    NcclComm comm = nullptr;
    const int rc = a.CommInitRank(&comm, h->cfg.world, u, h->cfg.rank);
    if (rc != kNcclSuccess) return fail("ncclCommInitRank", rc);
    h->nccl_comm = comm;
    return NB_OK;
}

// In-place all-gather: rank r contributes posw[first_r, first_r + count_r).  Requires equal shard
// sizes (n divisible by world): nb_comm_init and nb_step refuse anything else before a kernel is
// launched.  Ragged shards use the peer-memory exchange (p2p.cu), which has no such constraint.
int comm_allgather_posw(nb_sim* h)
{
    NcclApi& a = api();
    if (h->n % (size_t)h->cfg.world != 0)
    {
        set_error("NCCL exchange needs the body count to be divisible by the number of ranks");
        return NB_ERR_ARG;
    }
    const size_t bytes = h->count * sizeof(float4);
    const int rc = a.AllGather(h->posw + h->first, h->posw, bytes, kNcclChar, static_cast<NcclComm>(h->nccl_comm), h->stream);
    if (rc != kNcclSuccess) return fail("ncclAllGather", rc);
    h->exchanged = true;
    return NB_OK;
}

void comm_destroy(nb_sim* h)
{
    if (h->nccl_comm != nullptr && api().ok) api().CommDestroy(static_cast<NcclComm>(h->nccl_comm));
    h->nccl_comm = nullptr;
}

}  // namespace nb
