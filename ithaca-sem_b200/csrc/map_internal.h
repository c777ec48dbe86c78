// map_internal.h -- objects behind nekmf_map_t / nekmf_comm_t / nekmf_exchange_t.
#pragma once
#include "op_internal.h"

struct nekmf_map_s
{
    int nLocal = 0, nGlobal = 0;
    int *d_map = nullptr;      // localToGlobal [nLocal]
    double *d_sign = nullptr;  // [nLocal] or null
    int *d_rowptr = nullptr;   // transpose CSR [nGlobal+1]
    int *d_col = nullptr;      // [nLocal] local indices grouped by global DOF, ascending
    double *d_stage_loc = nullptr, *d_stage_glob = nullptr;
};

// Reduction window of a communicator in peer-memory mode: every rank owns one, every rank maps all of
// them (CUDA IPC).  Rank r deposits its partial sums in slot r of EVERY window and raises flag r there;
// each rank then adds the slots of its own window in ascending rank order, so all ranks hold bit-identical
// sums.  Two parities: a rank can run at most one reduction ahead of its slowest peer.
struct nekmf_redwin
{
    unsigned long long flag[2][16]; // [parity][source rank] = epoch of the deposit
    double val[2][16][4];           // [parity][source rank][value]
};
constexpr int NEKMF_MAX_RANKS = 16;

struct nekmf_comm_s
{
    void *nccl = nullptr; // ncclComm_t
    int rank = 0, nranks = 1;
    // peer-memory transport (NVLink loads/stores into CUDA-IPC mapped windows); false -> NCCL calls
    bool p2p = false;
    nekmf_redwin *win = nullptr;                 // my window (device memory, exported)
    nekmf_redwin *peer_win[NEKMF_MAX_RANKS] = {}; // every rank's window as mapped here ([rank] = win)
    nekmf_redwin **d_peer_win = nullptr;          // the same table on the device
    unsigned long long *d_red_epoch = nullptr;    // reductions done so far (device counter)
    int *d_err = nullptr;                         // set by a kernel whose peer wait timed out
    std::vector<void *> ipc_opened;               // mappings to close at destroy
};

// Device-side description of one interface exchange (passed to the kernels by value).
struct nekmf_exdev
{
    int total = 0, nU = 0, nNbr = 0, me = 0, nranks = 1;
    int wait = 0;                  // 1: peers deposit straight into my window and raise flags (peer-memory mode)
    long long rstride = 0;         // doubles between the two parities of my receive window (0: single buffer)
    const int *idx = nullptr;      // [total] my global index of every (neighbour, position) entry
    const unsigned char *nbr = nullptr; // [total] neighbour number of the entry
    const int *off = nullptr;      // [nNbr+1]
    double *const *put_base = nullptr;      // [nNbr] where entry (n, 0) goes: my segment in peer n's window
                                            // (peer-memory mode) or in my own send buffer (NCCL mode)
    const long long *put_stride = nullptr;  // [nNbr] parity stride at the destination
    unsigned long long *const *put_flag = nullptr; // [nNbr] &flag[0][me] in peer n's window header
    const unsigned long long *my_flag = nullptr;   // [2][nranks] in my window header
    const int *nbr_rank = nullptr; // [nNbr]
    const double *recv = nullptr;  // my receive window [2][total] (NCCL mode: [total])
    // every shared DOF once (ascending global index) with its receive slots in ascending peer rank
    const int *uidx = nullptr, *uptr = nullptr, *uslot = nullptr, *ulow = nullptr; // ulow: slots from ranks < me
    unsigned long long *epoch = nullptr; // exchanges completed (device counter)
    unsigned int *ticket = nullptr;      // last-block detection of the deposit kernel
    int *err = nullptr;
};

struct nekmf_exchange_s
{
    nekmf_comm_s *comm = nullptr;
    int nNeighbours = 0;
    std::vector<int> peers, offsets; // offsets[nNeighbours+1]
    int total = 0, nU = 0, nGlobal = 0;
    bool p2p = false;
    nekmf_exdev dev;          // what the kernels see
    void *d_window = nullptr; // header (flags) + receive data; exported in peer-memory mode
    double *d_send = nullptr; // NCCL mode staging [total]
    double *d_recv = nullptr; // = window data
    std::vector<void *> dev_allocs;
    std::vector<void *> ipc_opened;
    std::vector<int> h_uidx; // shared DOFs (unique, ascending) -- used by the CG to flag interface DOFs
};

namespace nekmf
{
constexpr int RED_BLOCKS = 1184; // 8 x 148: every SM holds its 2048 threads
constexpr int RED_T      = 256;
constexpr int IF_BLOCKS  = 148;  // blocks of the interface unpack kernel (one partial sum each)

int map_g2l_device(nekmf_map_s *m, const double *glob, double *loc, cudaStream_t st);
int map_assemble_device(nekmf_map_s *m, const double *loc, double *glob, cudaStream_t st);
// Assemble + partial sums (one per block, RED_BLOCKS blocks of 256 threads) of sum_{g>=nDir} owned(g)*glob*w.
// flags[g]: bit 0 = this rank owns g (dot products), bit 1 = g is shared with another rank (its s.w term is
// added after the exchange).  ex != null: the first blocks also deposit the interface values with the peers.
int map_assemble_dot_device(nekmf_map_s *m, const double *loc, double *glob, const double *w,
                            const unsigned char *flags, int nDir, double *part, const nekmf_exdev *ex,
                            cudaStream_t st, const int *skip = nullptr); // skip: single-rank only (no peer waits on it)
// glob[idx] += peers' glob[idx]; the building blocks are also used by the CG
int exchange_add_device(nekmf_exchange_s *ex, double *glob, cudaStream_t st);
int exchange_transport_device(nekmf_exchange_s *ex, cudaStream_t st); // NCCL send/recv of the staged values (no-op in peer-memory mode)
// wait for the peers' deposits, add own + received in ascending rank order; w != null: also the partial sums of
// owned(g) * glob[g] * w[g] over the shared DOFs g >= nDir into part[0..IF_BLOCKS)
int exchange_finish_device(nekmf_exchange_s *ex, double *glob, const double *w, const unsigned char *flags, int nDir,
                           double *part, cudaStream_t st);
int comm_allreduce_sum(nekmf_comm_s *c, double *d_buf, int n, cudaStream_t st);
int comm_check_error(nekmf_comm_s *c); // after a stream sync: did a peer wait time out?

// ---- device helpers shared by comm.cu / assembly.cu / cg.cu
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= epoch; gives up after ~4 s (a peer died) and records the failure instead of hanging the GPU
__device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long epoch, int *err)
{
    if (ld_acquire_sys(flag) >= epoch) return true;
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while (ld_acquire_sys(flag) < epoch)
    {
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 4000000000ull)
        {
            if (err) atomicExch(err, 1);
            return false;
        }
    }
    return true;
}
// the deposit half of an exchange for entry i: value v goes to my segment at neighbour nbr[i]
__device__ __forceinline__ void exchange_put(const nekmf_exdev &ex, int i, unsigned long long epoch, double v)
{
    const int n = ex.nbr[i];
    double *dst = ex.put_base[n] + (long long)(epoch & 1ull) * ex.put_stride[n] + (i - ex.off[n]);
    *dst        = v;
}
// called by every thread of a deposit block after its stores; the last of nBlocks blocks raises the flags
__device__ __forceinline__ void exchange_signal(const nekmf_exdev &ex, unsigned long long epoch, unsigned int nBlocks)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned int t = atomicAdd(ex.ticket, 1u);
        if (t == nBlocks - 1)
        {
            __threadfence_system();
            if (ex.wait)
                for (int n = 0; n < ex.nNbr; ++n) st_release_sys(ex.put_flag[n] + (epoch & 1ull) * ex.nranks, epoch);
            *ex.ticket = 0u;
            *ex.epoch  = epoch;
            __threadfence();
        }
    }
}
} // namespace nekmf
