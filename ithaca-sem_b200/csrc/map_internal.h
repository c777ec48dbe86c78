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

struct nekmf_comm_s
{
    void *nccl = nullptr; // ncclComm_t
    int rank = 0, nranks = 1;
};

struct nekmf_exchange_s
{
    nekmf_comm_s *comm = nullptr;
    int nNeighbours = 0;
    std::vector<int> peers, offsets; // offsets[nNeighbours+1]
    int total = 0;
    int *d_idx = nullptr;     // [total]
    double *d_send = nullptr; // [total]
    double *d_recv = nullptr; // [total]
};

namespace nekmf
{
int map_g2l_device(nekmf_map_s *m, const double *glob, double *loc, cudaStream_t st);
int map_assemble_device(nekmf_map_s *m, const double *loc, double *glob, cudaStream_t st);
// Assemble + partial sums (one per block, nBlocks blocks of 256 threads) of sum_{g>=nDir} mask*glob*w
int map_assemble_dot_device(nekmf_map_s *m, const double *loc, double *glob, const double *w, const double *mask,
                            int nDir, double *part, int nBlocks, cudaStream_t st);
int exchange_add_device(nekmf_exchange_s *ex, double *glob, cudaStream_t st);
int comm_allreduce_sum(nekmf_comm_s *c, double *d_buf, int n, cudaStream_t st);
} // namespace nekmf
