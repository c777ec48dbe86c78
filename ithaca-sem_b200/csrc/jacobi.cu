// jacobi.cu -- matrix-free diagonal (Jacobi) preconditioner for any shape and geometry.
//
// The reference builds it from the assembled ELEMENTAL MATRICES: PreconditionerDiagonal::DiagonalPreconditionerSum
// (MultiRegions/PreconditionerDiagonal.cpp:98-162) walks every local matrix block, adds loc_mat(i,i) into the global
// DOF of i, UniversalAssembles across ranks and inverts.  Here no elemental matrix exists.  The diagonal of the
// operator the solver actually applies is obtained from that operator itself: nmTot applies to the unit vectors
//     x^(k)[e][j] = delta_jk  for every element e at once   ->   A_e(k,k) = y^(k)[e][k],
// i.e. the same kernels, tables and geometric factors as the mat-vec (regular, deformed, collapsed, every kernel
// variant), so the preconditioner is consistent with the operator to the last bit.  Cost: nmTot operator applies
// once per (operator, lambda) -- 125 applies at hex P=4 -- against one dense nmTot x nmTot matrix per element in the
// reference.  The elemental diagonals are then assembled with the transposed map (assembly.cu), exchanged across
// the partition interfaces (comm.cu) and inverted.  (Two local DOFs of ONE element mapped to the same global DOF
// -- which the reference's double loop would also pick up -- do not occur for the meshes AssemblyMapCG produces
// without periodic single-element directions; such maps are rejected by the caller's test, not silently wrong:
// the result is then only an approximation of the diagonal, still a valid preconditioner.)
#include "cg_internal.h"

namespace nekmf
{
// x[e*nm + k] = 1, x[e*nm + kprev] = 0 (kprev < 0: nothing to clear)
__global__ void __launch_bounds__(256) probe_set_kernel(double *__restrict__ x, int nElmt, int nm, int k, int kprev)
{
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= nElmt) return;
    if (kprev >= 0) x[(size_t)e * nm + kprev] = 0.0;
    x[(size_t)e * nm + k] = 1.0;
}
__global__ void __launch_bounds__(256)
    probe_take_kernel(const double *__restrict__ y, double *__restrict__ diag, int nElmt, int nm, int k)
{
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= nElmt) return;
    diag[(size_t)e * nm + k] = y[(size_t)e * nm + k];
}
__global__ void __launch_bounds__(256) invert_kernel(const double *__restrict__ d, double *__restrict__ inv, int n)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) inv[i] = 1.0 / d[i];
}

// elemental diagonals of a Helmholtz operator into d_diag [nElmt*nmTot] (device); x, y: device work arrays of the
// same size (+2 doubles of slack for the TMA-fed kernels)
int op_diagonal_device(nekmf_op_s *op, double *d_diag, double *x, double *y, cudaStream_t st)
{
    const int nE = op->nElmt, nm = op->nmTot;
    if (nE == 0) return NEKMF_OK;
    NEKMF_CUDA(cudaMemsetAsync(x, 0, (size_t)nE * nm * 8, st));
    const int B = (nE + 255) / 256;
    const double *ins[3] = {x, x, x};
    double *outs[3]      = {y, y, y};
    for (int k = 0; k < nm; ++k)
    {
        probe_set_kernel<<<B, 256, 0, st>>>(x, nE, nm, k, k - 1);
        op->run_e0     = 0;
        op->run_ne     = nE;
        op->run_stream = st;
        const int rc   = op->launch(op, ins, outs);
        op->run_stream = op->stream;
        if (rc) return rc;
        probe_take_kernel<<<B, 256, 0, st>>>(y, d_diag, nE, nm, k);
        g_launches += 2;
    }
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}
} // namespace nekmf

using namespace nekmf;

extern "C" {

int nekmf_op_diagonal(nekmf_op_t op, double *diag, int memkind)
{
    if (!op || !diag) { set_error("nekmf_op_diagonal: null argument"); return NEKMF_ERR_ARG; }
    if (op->optype != NEKMF_HELMHOLTZ) { set_error("nekmf_op_diagonal: operator is not Helmholtz"); return NEKMF_ERR_ARG; }
    if (memkind != NEKMF_HOST && memkind != NEKMF_DEVICE) { set_error("nekmf_op_diagonal: bad memkind"); return NEKMF_ERR_ARG; }
    int rc = op_check_ready(op);
    if (rc) return rc;
    const size_t n = (size_t)op->nElmt * op->nmTot;
    if (n == 0) return NEKMF_OK;
    double *x = nullptr, *y = nullptr, *d = nullptr;
    cudaError_t e = cudaMalloc(&x, (n + 2) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&y, (n + 2) * 8);
    if (e == cudaSuccess && memkind == NEKMF_HOST) e = cudaMalloc(&d, n * 8);
    if (e == cudaSuccess)
    {
        rc = op_diagonal_device(op, memkind == NEKMF_HOST ? d : diag, x, y, op->stream);
        if (!rc && memkind == NEKMF_HOST) e = cudaMemcpyAsync(diag, d, n * 8, cudaMemcpyDeviceToHost, op->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(op->stream);
    }
    cudaFree(x); cudaFree(y); cudaFree(d);
    if (e != cudaSuccess) { set_error("nekmf_op_diagonal: %s", cudaGetErrorString(e)); return NEKMF_ERR_CUDA; }
    return rc;
}

int nekmf_cg_set_jacobi(nekmf_cg_t cg)
{
    if (!cg) { set_error("nekmf_cg_set_jacobi: null argument"); return NEKMF_ERR_ARG; }
    int rc = op_check_ready(cg->op);
    if (rc) return rc;
    cudaStream_t st = cg->stream;
    const size_t nL = (size_t)cg->nLocal;
    double *x = nullptr, *y = nullptr;
    cudaError_t e = cudaMalloc(&x, (nL + 2) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&y, (nL + 2) * 8);
    if (e == cudaSuccess && !cg->d_invdiag) e = cudaMalloc(&cg->d_invdiag, ((size_t)cg->nNonDir + 2) * 8);
    if (e != cudaSuccess)
    {
        cudaFree(x); cudaFree(y);
        set_error("nekmf_cg_set_jacobi: %s", cudaGetErrorString(e));
        return NEKMF_ERR_CUDA;
    }
    // elemental diagonals -> d_lout, Assemble -> d_s (global work vector of the solver), exchange, invert [nDir, nGlobal)
    rc = op_diagonal_device(cg->op, cg->d_lout, x, y, st);
    if (!rc) rc = map_assemble_device(cg->map, cg->d_lout, cg->d_s, st);
    if (!rc && cg->ex) rc = exchange_add_device(cg->ex, cg->d_s, st);
    if (!rc && cg->nNonDir > 0)
    {
        invert_kernel<<<(cg->nNonDir + 255) / 256, 256, 0, st>>>(cg->d_s + cg->nDir, cg->d_invdiag, cg->nNonDir);
        ++g_launches;
    }
    e = cudaStreamSynchronize(st);
    cudaFree(x); cudaFree(y);
    if (rc) return rc;
    if (e != cudaSuccess) { set_error("nekmf_cg_set_jacobi: %s", cudaGetErrorString(e)); return NEKMF_ERR_CUDA; }
    cg_invalidate_graphs(cg); // captured iterations may have been built without a preconditioner pointer
    return comm_check_error(cg->comm);
}

} // extern "C"
