// comm.cu -- one rank per GPU: communicator plus the interface-DOF exchange that replaces the
// reference's Gs::Gather(gs_add) (AssemblyMapCG::v_UniversalAssemble, AssemblyMapCG.cpp:2925-2939;
// LibUtilities/Communication/GsLib.hpp:145-151).  gslib's semantic is "every copy of a universal id
// ends up holding the sum over all copies".
//
// Two transports behind the same entry points:
//   peer memory (default on an NVLink / NVSwitch box): every rank exports a receive window with CUDA IPC
//     and maps its neighbours' windows.  The kernel that PRODUCES the interface values (the assemble kernel
//     of the CG mat-vec, or a pack kernel for a stand-alone exchange) stores them straight into the
//     neighbours' windows over NVLink and raises a flag there; the consumer kernel spins on its own flags.
//     No NCCL call, no host involvement, capturable in a CUDA graph.  The 3-double all-reduce of the CG
//     works the same way (nekmf_redwin).
//   NCCL (NEKMF_TRANSPORT=nccl, or when the windows cannot be mapped): pack -> grouped ncclSend/ncclRecv
//     -> the same unpack kernel; ncclAllReduce for the dot products.
// Either way the unpack adds own + received values of a shared DOF in ascending RANK order (the own value
// takes its rank's place), one thread per DOF, no atomics: deterministic, and every holder of a DOF computes
// the bit-identical sum.
//
// NCCL is bound lazily with dlopen so that the library loads on machines without it and picks up the copy
// the host process (torch) has already loaded; it is always used to bootstrap (handle exchange).
#include "map_internal.h"
#include <algorithm>
#include <dlfcn.h>
#include <stdlib.h>
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
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t)   = nullptr;
    const char *(*GetErrorString)(int)                                          = nullptr;
};
static NcclApi g_nccl;
static const int kNcclInt8 = 0, kNcclInt32 = 2, kNcclFloat64 = 8, kNcclSum = 0, kNcclMin = 3; // nccl.h values

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
    SYM(AllGather, "ncclAllGather")
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

static size_t round_2mb(size_t n) { return ((n + (2u << 20) - 1) >> 21) << 21; }

// every rank contributes n bytes, everybody gets all of them (setup only; synchronous)
static int comm_allgather_bytes(nekmf_comm_s *c, const void *mine, size_t n, std::vector<unsigned char> &all)
{
    all.assign(n * c->nranks, 0);
    unsigned char *d = nullptr;
    NEKMF_CUDA(cudaMalloc(&d, n * c->nranks));
    cudaError_t e = cudaMemcpy(d + n * c->rank, mine, n, cudaMemcpyHostToDevice);
    int r = 0;
    if (e == cudaSuccess) r = g_nccl.AllGather(d + n * c->rank, d, n, kNcclInt8, c->nccl, nullptr);
    if (e == cudaSuccess && r == 0) e = cudaStreamSynchronize(nullptr);
    if (e == cudaSuccess && r == 0) e = cudaMemcpy(all.data(), d, n * c->nranks, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (r != 0) { set_error("ncclAllGather failed: %s", g_nccl.GetErrorString(r)); return NEKMF_ERR_COMM; }
    if (e != cudaSuccess) { set_error("allgather staging failed: %s", cudaGetErrorString(e)); return NEKMF_ERR_CUDA; }
    return NEKMF_OK;
}
// do all ranks say yes?
static int comm_all_agree(nekmf_comm_s *c, bool mine, bool *all)
{
    int *d = nullptr;
    int v  = mine ? 1 : 0;
    NEKMF_CUDA(cudaMalloc(&d, 4));
    cudaError_t e = cudaMemcpy(d, &v, 4, cudaMemcpyHostToDevice);
    int r = 0;
    if (e == cudaSuccess) r = g_nccl.AllReduce(d, d, 1, kNcclInt32, kNcclMin, c->nccl, nullptr);
    if (e == cudaSuccess && r == 0) e = cudaStreamSynchronize(nullptr);
    if (e == cudaSuccess && r == 0) e = cudaMemcpy(&v, d, 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (r != 0) { set_error("ncclAllReduce failed: %s", g_nccl.GetErrorString(r)); return NEKMF_ERR_COMM; }
    if (e != cudaSuccess) { set_error("agreement staging failed: %s", cudaGetErrorString(e)); return NEKMF_ERR_CUDA; }
    *all = v != 0;
    return NEKMF_OK;
}

// Export `mine` (a cudaMalloc allocation) and map the allocations of the ranks in `want` (mapped[r] for r in want;
// mapped[rank] = mine).  ok = false (on every rank) when any mapping failed anywhere.
static int ipc_share(nekmf_comm_s *c, void *mine, const std::vector<int> &want, std::vector<void *> &mapped,
                     std::vector<void *> &opened, bool *ok)
{
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    bool good = cudaIpcGetMemHandle(&h, mine) == cudaSuccess;
    if (!good) cudaGetLastError();
    std::vector<unsigned char> all;
    int rc = comm_allgather_bytes(c, &h, sizeof(h), all);
    if (rc) return rc;
    mapped.assign(c->nranks, nullptr);
    mapped[c->rank] = mine;
    if (good)
        for (int r : want)
        {
            if (r == c->rank) continue;
            cudaIpcMemHandle_t hr;
            memcpy(&hr, all.data() + sizeof(h) * r, sizeof(hr));
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
            {
                cudaGetLastError();
                good = false;
                break;
            }
            mapped[r] = p;
            opened.push_back(p);
        }
    rc = comm_all_agree(c, good, ok);
    if (rc) return rc;
    if (!*ok)
    {
        for (void *p : opened) cudaIpcCloseMemHandle(p);
        opened.clear();
    }
    return NEKMF_OK;
}

// ------------------------------------------------------------------------------------------ kernels
// stand-alone exchange, deposit half: value of every (neighbour, position) entry -> the neighbour's window
__global__ void __launch_bounds__(256) exchange_put_kernel(const __grid_constant__ nekmf_exdev ex, const double *__restrict__ glob)
{
    const unsigned long long epoch = *ex.epoch + 1ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ex.total; i += gridDim.x * blockDim.x)
        exchange_put(ex, i, epoch, glob[ex.idx[i]]);
    exchange_signal(ex, epoch, gridDim.x);
}

// second half: wait for the neighbours' flags, then every shared DOF = sum of its copies in ascending rank order
__global__ void __launch_bounds__(RED_T)
    exchange_finish_kernel(const __grid_constant__ nekmf_exdev ex, double *__restrict__ glob, const double *__restrict__ w,
                           const unsigned char *__restrict__ flags, int nDir, double *__restrict__ part)
{
    __shared__ double sh[RED_T / 32];
    const unsigned long long epoch = *ex.epoch;
    if (ex.wait)
    {
        // one thread per neighbour spins on that neighbour's flag in my window (a timeout is recorded in *ex.err
        // and reported by the host; the values added below are then meaningless)
        if (threadIdx.x < ex.nNbr)
            wait_flag(ex.my_flag + (epoch & 1ull) * ex.nranks + ex.nbr_rank[threadIdx.x], epoch, ex.err);
        __syncthreads();
    }
    const double *recv = ex.recv + (long long)(epoch & 1ull) * ex.rstride;
    double mu          = 0.0;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < ex.nU; u += gridDim.x * blockDim.x)
    {
        const int g = ex.uidx[u], b = ex.uptr[u], e = ex.uptr[u + 1], nlow = ex.ulow[u];
        const double own = glob[g];
        double s;
        int k = b;
        if (nlow > 0)
        {
            s = ld_relaxed_sys(recv + ex.uslot[k++]);
            for (; k < b + nlow; ++k) s += ld_relaxed_sys(recv + ex.uslot[k]);
            s += own;
        }
        else
            s = own;
        for (; k < e; ++k) s += ld_relaxed_sys(recv + ex.uslot[k]);
        glob[g] = s;
        if (w && g >= nDir && (!flags || (flags[g] & 1))) mu = fma(s, w[g], mu);
    }
    if (part)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mu += __shfl_xor_sync(0xffffffffu, mu, o);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mu;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            double r = threadIdx.x < RED_T / 32 ? sh[threadIdx.x] : 0.0;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            if (threadIdx.x == 0) part[blockIdx.x] = r;
        }
    }
}

// all-reduce of n <= 4 doubles through the reduction windows (one block)
__global__ void __launch_bounds__(32)
    allreduce_window_kernel(nekmf_redwin *const *__restrict__ peer_win, int me, int nranks, unsigned long long *epoch_ctr,
                            int *err, double *__restrict__ buf, int n)
{
    const unsigned long long epoch = *epoch_ctr + 1ull;
    const int par                  = (int)(epoch & 1ull);
    const int t                    = threadIdx.x;
    if (t < nranks)
    {
        nekmf_redwin *pw = peer_win[t];
        for (int k = 0; k < n; ++k) pw->val[par][me][k] = buf[k];
        __threadfence_system();
        st_release_sys(&pw->flag[par][me], epoch);
    }
    nekmf_redwin *mw = peer_win[me];
    if (t < nranks) wait_flag(&mw->flag[par][t], epoch, err);
    __syncwarp();
    if (t < n)
    {
        double s = 0.0;
        for (int r = 0; r < nranks; ++r) s += ld_relaxed_sys(&mw->val[par][r][t]);
        buf[t] = s;
    }
    if (t == 0) *epoch_ctr = epoch;
}

int exchange_transport_device(nekmf_exchange_s *ex, cudaStream_t st)
{
    if (!ex || ex->total == 0 || ex->p2p) return NEKMF_OK;
    int first_err = 0;
    int r = g_nccl.GroupStart();
    if (r != 0) { set_error("ncclGroupStart failed: %s", g_nccl.GetErrorString(r)); return NEKMF_ERR_COMM; }
    for (int n = 0; n < ex->nNeighbours && !first_err; ++n)
    {
        const int o = ex->offsets[n], cnt = ex->offsets[n + 1] - o;
        if (cnt == 0) continue;
        first_err = g_nccl.Send(ex->d_send + o, cnt, kNcclFloat64, ex->peers[n], ex->comm->nccl, st);
        if (!first_err) first_err = g_nccl.Recv(ex->d_recv + o, cnt, kNcclFloat64, ex->peers[n], ex->comm->nccl, st);
    }
    r = g_nccl.GroupEnd(); // always close the group, also on the error path
    if (first_err || r)
    {
        set_error("ncclSend/ncclRecv failed: %s", g_nccl.GetErrorString(first_err ? first_err : r));
        return NEKMF_ERR_COMM;
    }
    return NEKMF_OK;
}

int exchange_finish_device(nekmf_exchange_s *ex, double *glob, const double *w, const unsigned char *flags, int nDir,
                           double *part, cudaStream_t st)
{
    if (!ex) return NEKMF_OK;
    if (ex->total == 0)
    {
        if (part) NEKMF_CUDA(cudaMemsetAsync(part, 0, IF_BLOCKS * 8, st));
        return NEKMF_OK;
    }
    exchange_finish_kernel<<<IF_BLOCKS, RED_T, 0, st>>>(ex->dev, glob, w, flags, nDir, part);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

int exchange_add_device(nekmf_exchange_s *ex, double *glob, cudaStream_t st)
{
    if (!ex || ex->total == 0) return NEKMF_OK;
    int blocks = (ex->total + 255) / 256;
    if (blocks > 4 * NUM_SMS) blocks = 4 * NUM_SMS;
    exchange_put_kernel<<<blocks, 256, 0, st>>>(ex->dev, glob);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    int rc = exchange_transport_device(ex, st);
    if (rc) return rc;
    return exchange_finish_device(ex, glob, nullptr, nullptr, 0, nullptr, st);
}

int comm_allreduce_sum(nekmf_comm_s *c, double *d_buf, int n, cudaStream_t st)
{
    if (!c || c->nranks == 1) return NEKMF_OK;
    if (c->p2p && n <= 4)
    {
        allreduce_window_kernel<<<1, 32, 0, st>>>(c->d_peer_win, c->rank, c->nranks, c->d_red_epoch, c->d_err, d_buf, n);
        ++g_launches;
        NEKMF_CUDA(cudaGetLastError());
        return NEKMF_OK;
    }
    NEKMF_NCCL(g_nccl.AllReduce(d_buf, d_buf, n, kNcclFloat64, kNcclSum, c->nccl, st));
    return NEKMF_OK;
}

int comm_check_error(nekmf_comm_s *c)
{
    if (!c || !c->d_err) return NEKMF_OK;
    int e = 0;
    NEKMF_CUDA(cudaMemcpy(&e, c->d_err, 4, cudaMemcpyDeviceToHost));
    if (e)
    {
        set_error("peer-memory exchange timed out waiting for another rank");
        return NEKMF_ERR_COMM;
    }
    return NEKMF_OK;
}

static bool want_p2p()
{
    const char *v = getenv("NEKMF_TRANSPORT");
    return !(v && (strcmp(v, "nccl") == 0 || strcmp(v, "NCCL") == 0));
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
    cudaError_t e = cudaMalloc(&c->d_err, 4);
    if (e == cudaSuccess) e = cudaMemset(c->d_err, 0, 4);
    if (e != cudaSuccess)
    {
        set_error("nekmf_comm_create: %s", cudaGetErrorString(e));
        nekmf_comm_destroy(c);
        return NEKMF_ERR_CUDA;
    }
    // peer-memory transport: one reduction window per rank, mapped everywhere
    if (nranks > 1 && nranks <= NEKMF_MAX_RANKS && want_p2p())
    {
        void *w = nullptr;
        e       = cudaMalloc(&w, round_2mb(sizeof(nekmf_redwin)));
        if (e == cudaSuccess) e = cudaMemset(w, 0, sizeof(nekmf_redwin));
        if (e == cudaSuccess) e = cudaMalloc(&c->d_red_epoch, 8);
        if (e == cudaSuccess) e = cudaMemset(c->d_red_epoch, 0, 8);
        if (e == cudaSuccess) e = cudaMalloc(&c->d_peer_win, sizeof(void *) * NEKMF_MAX_RANKS);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess)
        {
            set_error("nekmf_comm_create: %s", cudaGetErrorString(e));
            cudaFree(w);
            nekmf_comm_destroy(c);
            return NEKMF_ERR_CUDA;
        }
        c->win = static_cast<nekmf_redwin *>(w);
        std::vector<int> all_ranks(nranks);
        for (int r = 0; r < nranks; ++r) all_ranks[r] = r;
        std::vector<void *> mapped;
        bool ok = false;
        int rc  = ipc_share(c, w, all_ranks, mapped, c->ipc_opened, &ok);
        if (rc)
        {
            nekmf_comm_destroy(c);
            return rc;
        }
        if (ok)
        {
            for (int r = 0; r < nranks; ++r) c->peer_win[r] = static_cast<nekmf_redwin *>(mapped[r]);
            e = cudaMemcpy(c->d_peer_win, c->peer_win, sizeof(void *) * NEKMF_MAX_RANKS, cudaMemcpyHostToDevice);
            if (e != cudaSuccess)
            {
                set_error("nekmf_comm_create: %s", cudaGetErrorString(e));
                nekmf_comm_destroy(c);
                return NEKMF_ERR_CUDA;
            }
            c->p2p = true;
        }
    }
    *out = c;
    return NEKMF_OK;
}

int nekmf_comm_transport(nekmf_comm_t c)
{
    return c && c->p2p ? 1 : 0;
}

int nekmf_comm_destroy(nekmf_comm_t c)
{
    if (!c) return NEKMF_OK;
    cudaDeviceSynchronize();
    for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    cudaFree(c->win);
    cudaFree(c->d_peer_win);
    cudaFree(c->d_red_epoch);
    cudaFree(c->d_err);
    if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
    delete c;
    return NEKMF_OK;
}

int nekmf_exchange_create(nekmf_comm_t comm, int nGlobal, int nNeighbours, const int *peerRanks, const int *offsets,
                          const int *idx, nekmf_exchange_t *out)
{
    if (!out || nNeighbours < 0 || nNeighbours > 255) { set_error("nekmf_exchange_create: bad argument"); return NEKMF_ERR_ARG; }
    *out = nullptr;
    if (nNeighbours > 0 && (!comm || !peerRanks || !offsets || !idx))
    {
        set_error("nekmf_exchange_create: neighbours given without communicator / lists");
        return NEKMF_ERR_ARG;
    }
    const int total = nNeighbours > 0 ? offsets[nNeighbours] : 0;
    for (int n = 0; n < nNeighbours; ++n)
    {
        if (peerRanks[n] < 0 || peerRanks[n] >= comm->nranks || peerRanks[n] == comm->rank || offsets[n + 1] < offsets[n])
        {
            set_error("nekmf_exchange_create: bad peer rank or offsets at neighbour %d", n);
            return NEKMF_ERR_ARG;
        }
        for (int m = 0; m < n; ++m)
            if (peerRanks[m] == peerRanks[n]) { set_error("nekmf_exchange_create: peer %d listed twice", peerRanks[n]); return NEKMF_ERR_ARG; }
    }
    for (int i = 0; i < total; ++i)
        if (idx[i] < 0 || idx[i] >= nGlobal)
        {
            set_error("nekmf_exchange_create: idx[%d] = %d out of range [0,%d)", i, idx[i], nGlobal);
            return NEKMF_ERR_ARG;
        }
    nekmf_exchange_s *ex = new nekmf_exchange_s;
    ex->comm        = comm;
    ex->nNeighbours = nNeighbours;
    ex->nGlobal     = nGlobal;
    ex->offsets.assign(1, 0);
    ex->total = total;
    if (nNeighbours > 0)
    {
        ex->peers.assign(peerRanks, peerRanks + nNeighbours);
        ex->offsets.assign(offsets, offsets + nNeighbours + 1);
    }
    const int nranks = comm ? comm->nranks : 1, me = comm ? comm->rank : 0;
    ex->p2p = comm && comm->p2p && nranks > 1;

    // ---- host tables: neighbour of every entry; every shared DOF once with its receive slots by ascending rank
    std::vector<unsigned char> nbr(total ? total : 1);
    for (int n = 0; n < nNeighbours; ++n)
        for (int i = offsets[n]; i < offsets[n + 1]; ++i) nbr[i] = (unsigned char)n;
    std::vector<int> order(total);
    for (int i = 0; i < total; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        if (idx[a] != idx[b]) return idx[a] < idx[b];
        return peerRanks[nbr[a]] < peerRanks[nbr[b]];
    });
    std::vector<int> uidx, uptr(1, 0), uslot(total ? total : 1), ulow;
    for (int k = 0; k < total; ++k)
    {
        const int i = order[k];
        if (k == 0 || idx[i] != idx[order[k - 1]])
        {
            if (k) uptr.push_back(k);
            uidx.push_back(idx[i]);
            ulow.push_back(0);
        }
        else if (peerRanks[nbr[i]] == peerRanks[nbr[order[k - 1]]])
        {
            set_error("nekmf_exchange_create: DOF %d listed twice for peer %d", idx[i], peerRanks[nbr[i]]);
            delete ex;
            return NEKMF_ERR_ARG;
        }
        uslot[k] = i;
        if (peerRanks[nbr[i]] < me) ++ulow.back();
    }
    uptr.push_back(total);
    if (total == 0) uptr.assign(1, 0);
    ex->nU     = (int)uidx.size();
    ex->h_uidx = uidx;
    if (total == 0 && !ex->p2p)
    {
        *out = ex;
        return NEKMF_OK;
    }

    // ---- device memory.  Window = [flags: 2 x nranks u64, padded to 256 B][receive data: 2 x total doubles]
    const size_t hdr = ((size_t)2 * nranks * 8 + 255) & ~(size_t)255;
    const size_t wbytes = hdr + (size_t)2 * (total ? total : 1) * 8;
    cudaError_t e = cudaSuccess;
    auto dalloc   = [&](void **p, size_t bytes) {
        if (e != cudaSuccess) return;
        e = cudaMalloc(p, bytes ? bytes : 8);
        if (e == cudaSuccess) ex->dev_allocs.push_back(*p);
    };
    auto upload = [&](const void *h, size_t bytes) -> void * {
        void *p = nullptr;
        dalloc(&p, bytes);
        if (e == cudaSuccess && bytes) e = cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice);
        return p;
    };
    e = cudaMalloc(&ex->d_window, round_2mb(wbytes));
    if (e == cudaSuccess) e = cudaMemset(ex->d_window, 0, wbytes);
    ex->d_recv = reinterpret_cast<double *>(static_cast<unsigned char *>(ex->d_window) + hdr);
    dalloc(reinterpret_cast<void **>(&ex->d_send), (size_t)(total ? total : 1) * 8);
    nekmf_exdev &d = ex->dev;
    d.total = total; d.nU = ex->nU; d.nNbr = nNeighbours; d.me = me; d.nranks = nranks;
    d.idx      = static_cast<const int *>(upload(idx, (size_t)total * 4));
    d.nbr      = static_cast<const unsigned char *>(upload(nbr.data(), (size_t)total));
    d.off      = static_cast<const int *>(upload(ex->offsets.data(), ex->offsets.size() * 4));
    d.nbr_rank = static_cast<const int *>(upload(ex->peers.data(), (size_t)nNeighbours * 4));
    d.uidx     = static_cast<const int *>(upload(uidx.data(), uidx.size() * 4));
    d.uptr     = static_cast<const int *>(upload(uptr.data(), uptr.size() * 4));
    d.uslot    = static_cast<const int *>(upload(uslot.data(), (size_t)total * 4));
    d.ulow     = static_cast<const int *>(upload(ulow.data(), ulow.size() * 4));
    d.recv     = ex->d_recv;
    d.my_flag  = static_cast<const unsigned long long *>(ex->d_window);
    d.err      = comm ? comm->d_err : nullptr;
    void *ctr  = nullptr;
    dalloc(&ctr, 16);
    if (e == cudaSuccess) e = cudaMemset(ctr, 0, 16);
    d.epoch  = static_cast<unsigned long long *>(ctr);
    d.ticket = reinterpret_cast<unsigned int *>(static_cast<unsigned char *>(ctr) + 8);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess)
    {
        set_error("nekmf_exchange_create: %s", cudaGetErrorString(e));
        nekmf_exchange_destroy(ex);
        return NEKMF_ERR_CUDA;
    }

    std::vector<double *> put_base(nNeighbours ? nNeighbours : 1, nullptr);
    std::vector<long long> put_stride(nNeighbours ? nNeighbours : 1, 0);
    std::vector<unsigned long long *> put_flag(nNeighbours ? nNeighbours : 1, nullptr);
    if (ex->p2p)
    {
        // where does my segment start in each neighbour's window?  every rank publishes, per rank r, the offset
        // of r's segment in its own window (-1: not a neighbour) and its total
        std::vector<long long> mine(nranks + 1, -1), table;
        for (int n = 0; n < nNeighbours; ++n) mine[peerRanks[n]] = offsets[n];
        mine[nranks] = total;
        std::vector<unsigned char> all;
        int rc = comm_allgather_bytes(comm, mine.data(), mine.size() * 8, all);
        std::vector<void *> mapped;
        bool ok = false;
        if (!rc) rc = ipc_share(comm, ex->d_window, ex->peers, mapped, ex->ipc_opened, &ok);
        if (rc)
        {
            nekmf_exchange_destroy(ex);
            return rc;
        }
        table.resize((size_t)nranks * (nranks + 1));
        memcpy(table.data(), all.data(), table.size() * 8);
        bool consistent = true;
        for (int n = 0; n < nNeighbours && ok; ++n)
        {
            const int r          = peerRanks[n];
            const long long *row = table.data() + (size_t)r * (nranks + 1);
            if (row[me] < 0) { consistent = false; break; }
            const int rtotal   = (int)row[nranks];
            const size_t rhdr  = hdr; // same nranks everywhere
            unsigned char *base = static_cast<unsigned char *>(mapped[r]);
            put_base[n]   = reinterpret_cast<double *>(base + rhdr) + row[me];
            put_stride[n] = rtotal ? rtotal : 1;
            put_flag[n]   = reinterpret_cast<unsigned long long *>(base) + me;
        }
        bool all_consistent = false;
        rc = comm_all_agree(comm, consistent, &all_consistent);
        if (rc)
        {
            nekmf_exchange_destroy(ex);
            return rc;
        }
        if (!all_consistent)
        {
            set_error("nekmf_exchange_create: neighbour lists are not symmetric across ranks");
            nekmf_exchange_destroy(ex);
            return NEKMF_ERR_ARG;
        }
        if (!ok) ex->p2p = false; // windows could not be mapped somewhere: every rank falls back to NCCL transport
    }
    if (ex->p2p)
    {
        d.wait    = 1;
        d.rstride = total ? total : 1;
    }
    else
    {
        for (int n = 0; n < nNeighbours; ++n) put_base[n] = ex->d_send + offsets[n];
        d.wait    = 0;
        d.rstride = 0;
    }
    d.put_base   = static_cast<double *const *>(upload(put_base.data(), put_base.size() * sizeof(void *)));
    d.put_stride = static_cast<const long long *>(upload(put_stride.data(), put_stride.size() * 8));
    d.put_flag   = static_cast<unsigned long long *const *>(upload(put_flag.data(), put_flag.size() * sizeof(void *)));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess)
    {
        set_error("nekmf_exchange_create: %s", cudaGetErrorString(e));
        nekmf_exchange_destroy(ex);
        return NEKMF_ERR_CUDA;
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
    cudaDeviceSynchronize();
    for (void *p : ex->ipc_opened) cudaIpcCloseMemHandle(p);
    for (void *p : ex->dev_allocs) cudaFree(p);
    cudaFree(ex->d_window);
    delete ex;
    return NEKMF_OK;
}

} // extern "C"
