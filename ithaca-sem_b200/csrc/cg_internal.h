// cg_internal.h -- the solver object behind nekmf_cg_t (cg.cu), shared with the HelmSolve chain (contfield.cu).
#pragma once
#include "map_internal.h"
#include <string>

namespace nekmf
{
// recurrence state of one solve, resident on the device
struct CgScal
{
    double alpha, beta, rho, mu, eps, rhs_mag, tol2;
    int its;     // m_totalIterations
    int k;       // the reference's loop counter
    int maxiter;
    int done;    // 0 running, 1 converged, 2 iteration cap reached (reference: efatal), 3 peer wait timed out
};
} // namespace nekmf

struct nekmf_cg_s
{
    nekmf_op_s *op           = nullptr;
    nekmf_map_s *map         = nullptr;
    nekmf_exchange_s *ex     = nullptr;
    nekmf_comm_s *comm       = nullptr;
    int nDir = 0, nGlobal = 0, nLocal = 0, nNonDir = 0;
    double *d_invdiag = nullptr;
    unsigned char *d_flags = nullptr; // [nGlobal] bit 0: owned by this rank, bit 1: shared with another rank; null = all owned
    double *d_w = nullptr, *d_s = nullptr, *d_p = nullptr, *d_r = nullptr, *d_q = nullptr; // w,s: nGlobal
    double *d_lin = nullptr, *d_lout = nullptr;                                            // nLocal
    double *d_x = nullptr, *d_rhs = nullptr;                                               // staging for host calls
    double *d_part = nullptr; // [3][RED_BLOCKS] + [IF_BLOCKS] partial sums
    double *d_red  = nullptr; // [4] reduced values
    double *h_red  = nullptr; // pinned [4]
    nekmf::CgScal *d_scal = nullptr;
    nekmf::CgScal *h_scal = nullptr; // pinned
    int *h_done = nullptr;           // pinned [2]
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t ev_loop[2] = {nullptr, nullptr}; // around the iteration loop of the last solve
    int loop_iterations = 0;                     // iterations launched between them
    cudaStream_t stream = nullptr;
    // captured iterations: graph[0] = GRAPH_ITERS iterations, graph[1] = one; valid for (x, lambda, kernel)
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    const double *graph_x = nullptr;
    double graph_lambda   = 0.0;
    std::string graph_kernel;
};

namespace nekmf
{
void cg_invalidate_graphs(nekmf_cg_s *cg); // cg.cu: drop the captured iterations (a pointer they bake in changed)
int op_diagonal_device(nekmf_op_s *op, double *d_diag, double *x, double *y, cudaStream_t st); // jacobi.cu
} // namespace nekmf
