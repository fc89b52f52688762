// NCCL plumbing of bin-sharded sessions.  The reference has no multi-device form of this path (it is one
// process with host threads, GC.cpp:1676); what is exchanged here is described at BatchCtx::exchange_tuples (session.cpp).
// NCCL is looked up at run time: the copy already mapped into the process (e.g. the one PyTorch loaded) is preferred, then
// $GANON_B200_NCCL, then the system's libnccl.so.2.  Only types come from <nccl.h>.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "comm.h"

namespace gnb
{
namespace
{
struct NcclApi
{
    void *handle = nullptr;
    ncclResult_t (*GetVersion)(int *)                                                                              = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *)                                                                    = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int)                                             = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t)                                                                        = nullptr;
    const char *(*GetErrorString)(ncclResult_t)                                                                    = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)              = nullptr;
    std::string error;
};

NcclApi &api()
{
    static NcclApi    a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("GANON_B200_NCCL");
        a.handle        = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!a.handle && env && env[0])
            a.handle = dlopen(env, RTLD_NOW | RTLD_LOCAL);
        if (!a.handle)
            a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!a.handle)
        {
            a.error = std::string("NCCL not found (libnccl.so.2; set GANON_B200_NCCL to its path): ") + dlerror();
            return;
        }
        bool ok = true;
        auto sym = [&](const char *name) {
            void *p = dlsym(a.handle, name);
            if (!p)
            {
                ok      = false;
                a.error = std::string("NCCL symbol missing: ") + name;
            }
            return p;
        };
        a.GetVersion     = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
        a.GetUniqueId    = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank   = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommDestroy    = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        a.AllGather      = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        if (!ok)
        {
            dlclose(a.handle);
            a.handle = nullptr;
        }
    });
    return a;
}

int need_api()
{
    if (!api().handle)
        return fail(GNB_ERR_CUDA, api().error);
    return GNB_OK;
}

#define GNB_NCCL(call)                                                                                   \
    do                                                                                                   \
    {                                                                                                    \
        ncclResult_t _r = (call);                                                                        \
        if (_r != ncclSuccess)                                                                           \
            return gnb::fail(GNB_ERR_CUDA, std::string(#call) + ": " + gnb::api().GetErrorString(_r));  \
    } while (0)
} // namespace

int comm_all_gather(void *nccl_comm, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t st)
{
    GNB_TRY(need_api());
    GNB_NCCL(api().AllGather(send, recv, bytes_per_rank, ncclUint8, static_cast<ncclComm_t>(nccl_comm), st));
    return GNB_OK;
}

} // namespace gnb

using namespace gnb;

gnb_comm::~gnb_comm()
{
    if (api().handle)
    {
        if (nccl)
            api().CommDestroy(static_cast<ncclComm_t>(nccl));
        if (nccl_in)
            api().CommDestroy(static_cast<ncclComm_t>(nccl_in));
    }
}

extern "C" int gnb_comm_unique_id(void *id, uint64_t cap)
{
    if (!id || cap < GNB_COMM_ID_BYTES)
        return fail(GNB_ERR_ARG, "gnb_comm_unique_id: the buffer must hold GNB_COMM_ID_BYTES bytes");
    GNB_TRY(need_api());
    static_assert(GNB_COMM_ID_BYTES == 2 * sizeof(ncclUniqueId), "two NCCL ids: compute-stream and ingest-stream communicators");
    ncclUniqueId a, b;
    GNB_NCCL(api().GetUniqueId(&a));
    GNB_NCCL(api().GetUniqueId(&b));
    memcpy(id, &a, sizeof a);
    memcpy(static_cast<char *>(id) + sizeof a, &b, sizeof b);
    return GNB_OK;
}

extern "C" int gnb_comm_create(const void *id, int rank, int n_ranks, int device, gnb_comm **out)
{
    if (!id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return fail(GNB_ERR_ARG, "gnb_comm_create: bad arguments");
    GNB_TRY(need_api());
    GNB_CUDA(cudaSetDevice(device));
    ncclUniqueId a, b;
    memcpy(&a, id, sizeof a);
    memcpy(&b, static_cast<const char *>(id) + sizeof a, sizeof b);
    auto c     = new gnb_comm();
    c->rank    = rank;
    c->n_ranks = n_ranks;
    c->device  = device;
    api().GetVersion(&c->nccl_version);
    ncclComm_t x = nullptr, y = nullptr;
    ncclResult_t r = api().CommInitRank(&x, n_ranks, a, rank);
    if (r == ncclSuccess)
    {
        c->nccl = x;
        r       = api().CommInitRank(&y, n_ranks, b, rank);
    }
    if (r != ncclSuccess)
    {
        delete c;
        return fail(GNB_ERR_CUDA, std::string("ncclCommInitRank: ") + api().GetErrorString(r));
    }
    c->nccl_in = y;
    *out       = c;
    return GNB_OK;
}

extern "C" int gnb_comm_info(const gnb_comm *c, int *rank, int *n_ranks, int *device, int *nccl_version)
{
    if (!c)
        return fail(GNB_ERR_ARG, "gnb_comm_info: null communicator");
    if (rank)
        *rank = c->rank;
    if (n_ranks)
        *n_ranks = c->n_ranks;
    if (device)
        *device = c->device;
    if (nccl_version)
        *nccl_version = c->nccl_version;
    return GNB_OK;
}

extern "C" void gnb_comm_free(gnb_comm *c) { delete c; }
