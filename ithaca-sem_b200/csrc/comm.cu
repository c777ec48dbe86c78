// comm.cu -- one rank per GPU: NCCL communicator plus the interface-DOF exchange that replaces
// the reference's Gs::Gather(gs_add) (AssemblyMapCG::v_UniversalAssemble, AssemblyMapCG.cpp:
// 2925-2939; LibUtilities/Communication/GsLib.hpp:145-151).  gslib's semantic is "every copy of
// a universal id ends up holding the sum over all copies"; here every rank packs the current
// values of the DOFs it shares with each neighbour, exchanges them with grouped
// ncclSend/ncclRecv over NVLink, and adds what it received.
//
// NCCL is bound lazily with dlopen so that the library loads on machines without it and picks
// up the copy the host process (torch) has already loaded.
#include "map_internal.h"
#include <dlfcn.h>
#include <string.h>

namespace nekmf
{
typedef struct { char internal[128]; } nccl_uid;
struct NcclApi
{
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_uid *)                                              = nullptr;
    int (*CommInitRank)(void **, int, nccl_uid, int)                            = nullptr;
    int (*CommDestroy)(void *)                                                  = nullptr;
    int (*GroupStart)()                                                         = nullptr;
    int (*GroupEnd)()                                                           = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t)           = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t)                 = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int)                                          = nullptr;
};
static NcclApi g_nccl;
static const int kNcclFloat64 = 8, kNcclSum = 0; // ncclDataType_t / ncclRedOp_t values (nccl.h)

static bool nccl_load()
{
    if (g_nccl.lib) return true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h)
    {
        set_error("cannot load libnccl.so.2: %s", dlerror());
        return false;
    }
#define SYM(field, name)                                                       \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                \
    if (!g_nccl.field)                                                         \
    {                                                                          \
        set_error("libnccl: missing symbol %s", name);                         \
        return false;                                                          \
    }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = h;
    return true;
}

#define NEKMF_NCCL(call)                                                                          \
    do                                                                                            \
    {                                                                                             \
        int _r = (call);                                                                          \
        if (_r != 0)                                                                              \
        {                                                                                         \
            set_error("%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
            return NEKMF_ERR_COMM;                                                                \
        }                                                                                         \
    } while (0)

__global__ void pack_kernel(const int *__restrict__ idx, const double *__restrict__ glob, double *__restrict__ buf, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = glob[idx[i]];
}
// a DOF can appear in several neighbour lists (edges / corners of the partition): atomics keep
// the adds race-free; each (DOF, neighbour) pair appears once so the result is the full sum
__global__ void unpack_add_kernel(const int *__restrict__ idx, const double *__restrict__ buf, double *__restrict__ glob, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(glob + idx[i], buf[i]);
}

int exchange_add_device(nekmf_exchange_s *ex, double *glob, cudaStream_t st)
{
    if (!ex || ex->total == 0) return NEKMF_OK;
    const int threads = 256, blocks = (ex->total + threads - 1) / threads;
    pack_kernel<<<blocks, threads, 0, st>>>(ex->d_idx, glob, ex->d_send, ex->total);
    ++g_launches;
    NEKMF_NCCL(g_nccl.GroupStart());
    for (int n = 0; n < ex->nNeighbours; ++n)
    {
        const int o = ex->offsets[n], cnt = ex->offsets[n + 1] - o;
        if (cnt == 0) continue;
        NEKMF_NCCL(g_nccl.Send(ex->d_send + o, cnt, kNcclFloat64, ex->peers[n], ex->comm->nccl, st));
        NEKMF_NCCL(g_nccl.Recv(ex->d_recv + o, cnt, kNcclFloat64, ex->peers[n], ex->comm->nccl, st));
    }
    NEKMF_NCCL(g_nccl.GroupEnd());
    unpack_add_kernel<<<blocks, threads, 0, st>>>(ex->d_idx, ex->d_recv, glob, ex->total);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

int comm_allreduce_sum(nekmf_comm_s *c, double *d_buf, int n, cudaStream_t st)
{
    if (!c || c->nranks == 1) return NEKMF_OK;
    NEKMF_NCCL(g_nccl.AllReduce(d_buf, d_buf, n, kNcclFloat64, kNcclSum, c->nccl, st));
    return NEKMF_OK;
}
} // namespace nekmf

using namespace nekmf;

extern "C" {

int nekmf_comm_unique_id(unsigned char id[128])
{
    if (!id) { set_error("nekmf_comm_unique_id: null id"); return NEKMF_ERR_ARG; }
    if (!nccl_load()) return NEKMF_ERR_COMM;
    nccl_uid u;
    NEKMF_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return NEKMF_OK;
}

int nekmf_comm_create(const unsigned char id[128], int rank, int nranks, nekmf_comm_t *out)
{
    if (!id || !out || rank < 0 || rank >= nranks) { set_error("nekmf_comm_create: bad argument"); return NEKMF_ERR_ARG; }
    *out = nullptr;
    if (!nccl_load()) return NEKMF_ERR_COMM;
    nccl_uid u;
    memcpy(u.internal, id, 128);
    void *comm = nullptr;
    NEKMF_NCCL(g_nccl.CommInitRank(&comm, nranks, u, rank));
    nekmf_comm_s *c = new nekmf_comm_s;
    c->nccl   = comm;
    c->rank   = rank;
    c->nranks = nranks;
    *out      = c;
    return NEKMF_OK;
}

int nekmf_comm_destroy(nekmf_comm_t c)
{
    if (!c) return NEKMF_OK;
    if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
    delete c;
    return NEKMF_OK;
}

int nekmf_exchange_create(nekmf_comm_t comm, int nNeighbours, const int *peerRanks, const int *offsets, const int *idx,
                          nekmf_exchange_t *out)
{
    if (!out || nNeighbours < 0) { set_error("nekmf_exchange_create: bad argument"); return NEKMF_ERR_ARG; }
    *out = nullptr;
    if (nNeighbours > 0 && (!comm || !peerRanks || !offsets || !idx))
    {
        set_error("nekmf_exchange_create: neighbours given without communicator / lists");
        return NEKMF_ERR_ARG;
    }
    nekmf_exchange_s *ex = new nekmf_exchange_s;
    ex->comm        = comm;
    ex->nNeighbours = nNeighbours;
    ex->offsets.assign(1, 0);
    if (nNeighbours > 0)
    {
        ex->peers.assign(peerRanks, peerRanks + nNeighbours);
        ex->offsets.assign(offsets, offsets + nNeighbours + 1);
        ex->total = offsets[nNeighbours];
    }
    if (ex->total > 0)
    {
        cudaError_t e = cudaMalloc(&ex->d_idx, (size_t)ex->total * 4);
        if (e == cudaSuccess) e = cudaMalloc(&ex->d_send, (size_t)ex->total * 8);
        if (e == cudaSuccess) e = cudaMalloc(&ex->d_recv, (size_t)ex->total * 8);
        if (e == cudaSuccess) e = cudaMemcpy(ex->d_idx, idx, (size_t)ex->total * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess)
        {
            set_error("nekmf_exchange_create: %s", cudaGetErrorString(e));
            nekmf_exchange_destroy(ex);
            return NEKMF_ERR_CUDA;
        }
    }
    *out = ex;
    return NEKMF_OK;
}

int nekmf_exchange_add(nekmf_exchange_t ex, double *glob, void *stream)
{
    if (!ex || !glob) { set_error("nekmf_exchange_add: null argument"); return NEKMF_ERR_ARG; }
    return exchange_add_device(ex, glob, static_cast<cudaStream_t>(stream));
}

int nekmf_exchange_destroy(nekmf_exchange_t ex)
{
    if (!ex) return NEKMF_OK;
    cudaFree(ex->d_idx);
    cudaFree(ex->d_send);
    cudaFree(ex->d_recv);
    delete ex;
    return NEKMF_OK;
}

} // extern "C"
