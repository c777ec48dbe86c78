// cg.cu -- matrix-free preconditioned conjugate gradient on A = Assemble o Helmholtz o GlobalToLocal.
//
// Mirrors NekLinSysIterCG::DoConjugateGradient (LibUtilities/LinearAlgebra/NekLinSysIterCG.cpp:
// 104-265): the single-reduction Demmel/Heath/van der Vorst variant with ONE 3-value reduction
// per iteration, the mat-vec of GlobalLinSysIterativeFull::v_DoMatrixMultiply
// (MultiRegions/GlobalLinSysIterativeFull.cpp:215-257) and a diagonal (or identity)
// preconditioner (MultiRegions/PreconditionerDiagonal.cpp).
//
// Device design: all vectors resident.  One iteration is four kernels:
//   cg_update_dots       the four axpy updates + preconditioner + partial sums of r.w and r.r
//   Helmholtz (gather)   GlobalToLocal fused into the operator's loads (hex_kron.cu GATHER variant:
//                        cp.async indirect loads straight into shared memory); other operator
//                        kernels are preceded by the separate gather kernel
//   assemble_dot         transpose-CSR Assemble + partial sums of s.w (a separate dot pass after
//                        the NCCL interface exchange when the solve is sharded)
//   dot3_final           fixed-order reduction of the partials -> one ncclAllReduce of 3 doubles
// Reductions use a fixed grid and order (deterministic, ownership-masked for multi-rank runs);
// work buffers are allocated once (the reference allocates 2 x nLocal every mat-vec,
// GlobalLinSysIterativeFull.cpp:225-226).
#include "map_internal.h"

struct nekmf_cg_s
{
    nekmf_op_s *op           = nullptr;
    nekmf_map_s *map         = nullptr;
    nekmf_exchange_s *ex     = nullptr;
    nekmf_comm_s *comm       = nullptr;
    int nDir = 0, nGlobal = 0, nLocal = 0, nNonDir = 0;
    double *d_invdiag = nullptr, *d_mask = nullptr;
    double *d_w = nullptr, *d_s = nullptr, *d_p = nullptr, *d_r = nullptr, *d_q = nullptr; // w,s: nGlobal
    double *d_lin = nullptr, *d_lout = nullptr;                                            // nLocal
    double *d_x = nullptr, *d_rhs = nullptr;                                               // staging for host calls
    double *d_part = nullptr; // [3][RED_BLOCKS] partial sums
    double *d_red  = nullptr; // [4] reduced values
    double *h_red  = nullptr; // pinned [4]
    cudaStream_t stream = nullptr;
};

namespace nekmf
{
constexpr int RED_BLOCKS = 1184; // 8 x 148: every SM holds its 2048 threads
constexpr int RED_T      = 256;

__device__ __forceinline__ double block_sum(double v, double *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0)
    {
        r = l < RED_T / 32 ? sh[l] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;
}

// partial sums of (a.b, c.d, e.f) with optional 0/1 ownership mask; null pointers skip a product
__global__ void __launch_bounds__(RED_T)
    dot3_partial(const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ c,
                 const double *__restrict__ d, const double *__restrict__ e, const double *__restrict__ f,
                 const double *__restrict__ mask, int n, double *__restrict__ part)
{
    __shared__ double sh[RED_T / 32];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = blockIdx.x * RED_T + threadIdx.x; i < n; i += RED_BLOCKS * RED_T)
    {
        const double m = mask ? mask[i] : 1.0;
        if (a) s0 = fma(a[i] * m, b[i], s0);
        if (c) s1 = fma(c[i] * m, d[i], s1);
        if (e) s2 = fma(e[i] * m, f[i], s2);
    }
    s0 = block_sum(s0, sh);
    s1 = block_sum(s1, sh);
    s2 = block_sum(s2, sh);
    if (threadIdx.x == 0)
    {
        part[blockIdx.x]                  = s0;
        part[RED_BLOCKS + blockIdx.x]     = s1;
        part[2 * RED_BLOCKS + blockIdx.x] = s2;
    }
}
__global__ void __launch_bounds__(RED_T) dot3_final(const double *__restrict__ part, double *__restrict__ red)
{
    __shared__ double sh[RED_T / 32];
    for (int k = 0; k < 3; ++k)
    {
        double s = 0.0;
        for (int i = threadIdx.x; i < RED_BLOCKS; i += RED_T) s += part[k * RED_BLOCKS + i];
        s = block_sum(s, sh);
        if (threadIdx.x == 0) red[k] = s;
    }
}

// p = beta p + w ; q = beta q + s ; x += alpha p ; r -= alpha q ; w = M^-1 r
// (NekLinSysIterCG.cpp:209-220); w,s,x are offset to the first non-Dirichlet DOF.
// Fused with the two dot products that do not depend on the next mat-vec:
// rho = r.w and eps = r.r (ownership-masked).  Fixed grid, grid-stride -> deterministic partial sums in
// part[0..RED_BLOCKS) (rho) and part[2*RED_BLOCKS..) (eps); the s.w partials come from the assemble kernel.
__global__ void __launch_bounds__(RED_T)
    cg_update_dots(double *__restrict__ p, double *__restrict__ q, double *__restrict__ x, double *__restrict__ r,
                   double *__restrict__ w, const double *__restrict__ s, const double *__restrict__ invdiag,
                   const double *__restrict__ mask, double alpha, double beta, int n, double *__restrict__ part)
{
    __shared__ double sh[RED_T / 32];
    double rho = 0.0, eps = 0.0;
    for (int i = blockIdx.x * RED_T + threadIdx.x; i < n; i += RED_BLOCKS * RED_T)
    {
        const double pi = fma(beta, p[i], w[i]);
        const double qi = fma(beta, q[i], s[i]);
        const double ri = fma(-alpha, qi, r[i]);
        const double wi = invdiag ? ri * invdiag[i] : ri;
        p[i] = pi;
        q[i] = qi;
        x[i] = fma(alpha, pi, x[i]);
        r[i] = ri;
        w[i] = wi;
        const double rm = mask ? ri * mask[i] : ri;
        rho = fma(rm, wi, rho);
        eps = fma(rm, ri, eps);
    }
    rho = block_sum(rho, sh);
    eps = block_sum(eps, sh);
    if (threadIdx.x == 0)
    {
        part[blockIdx.x]                  = rho;
        part[2 * RED_BLOCKS + blockIdx.x] = eps;
    }
}
// partial sums of one masked product into part[0..RED_BLOCKS)
__global__ void __launch_bounds__(RED_T)
    dot1_partial(const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ mask, int n,
                 double *__restrict__ part)
{
    __shared__ double sh[RED_T / 32];
    double s0 = 0.0;
    for (int i = blockIdx.x * RED_T + threadIdx.x; i < n; i += RED_BLOCKS * RED_T)
        s0 = fma(mask ? a[i] * mask[i] : a[i], b[i], s0);
    s0 = block_sum(s0, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = s0;
}
__global__ void cg_precon(double *__restrict__ w, const double *__restrict__ r, const double *__restrict__ invdiag, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = invdiag ? r[i] * invdiag[i] : r[i];
}

// s = Assemble(Helmholtz(GlobalToLocal(w))) (+ interface exchange).  When the operator kernel can gather
// (op->gather_ok) the GlobalToLocal pass is fused into its loads and no local input vector is written.
// mu_part != null: the assemble kernel also leaves the partial sums of s.w over [nDir, nGlobal) there
// (only valid without an exchange step: the interface contributions arrive after the assemble).
static int cg_matvec_device(nekmf_cg_s *cg, const double *w, double *s, double *mu_part = nullptr)
{
    int rc;
    nekmf_op_s *op = cg->op;
    double *out[3] = {cg->d_lout, cg->d_lout, cg->d_lout};
    op->run_e0     = 0;
    op->run_ne     = op->nElmt;
    op->run_stream = cg->stream;
    if (op->gather_ok)
    {
        const double *in[3] = {w, w, w};
        op->gather_map      = cg->map->d_map;
        op->gather_sign     = cg->map->d_sign;
        rc                  = op->launch(op, in, out);
        op->gather_map      = nullptr;
        op->gather_sign     = nullptr;
    }
    else
    {
        rc = map_g2l_device(cg->map, w, cg->d_lin, cg->stream);
        if (rc) return rc;
        const double *in[3] = {cg->d_lin, cg->d_lin, cg->d_lin};
        rc                  = op->launch(op, in, out);
    }
    if (rc) return rc;
    if (mu_part && cg->nGlobal > 0)
        rc = map_assemble_dot_device(cg->map, cg->d_lout, s, w, cg->d_mask, cg->nDir, mu_part, RED_BLOCKS, cg->stream);
    else
        rc = map_assemble_device(cg->map, cg->d_lout, s, cg->stream);
    if (rc) return rc;
    if (cg->ex) rc = exchange_add_device(cg->ex, s, cg->stream);
    return rc;
}

// reduce the three partial-sum rows in d_part, all-reduce across ranks, bring the 3 values to the host
static int cg_finish_dots(nekmf_cg_s *cg, double out[3])
{
    dot3_final<<<1, RED_T, 0, cg->stream>>>(cg->d_part, cg->d_red);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    int rc = comm_allreduce_sum(cg->comm, cg->d_red, 3, cg->stream);
    if (rc) return rc;
    NEKMF_CUDA(cudaMemcpyAsync(cg->h_red, cg->d_red, 3 * 8, cudaMemcpyDeviceToHost, cg->stream));
    NEKMF_CUDA(cudaStreamSynchronize(cg->stream));
    out[0] = cg->h_red[0];
    out[1] = cg->h_red[1];
    out[2] = cg->h_red[2];
    return NEKMF_OK;
}

static int cg_dots(nekmf_cg_s *cg, const double *a, const double *b, const double *c, const double *d, const double *e,
                   const double *f, const double *mask, int n, double out[3])
{
    dot3_partial<<<RED_BLOCKS, RED_T, 0, cg->stream>>>(a, b, c, d, e, f, mask, n, cg->d_part);
    dot3_final<<<1, RED_T, 0, cg->stream>>>(cg->d_part, cg->d_red);
    g_launches += 2;
    NEKMF_CUDA(cudaGetLastError());
    int rc = comm_allreduce_sum(cg->comm, cg->d_red, 3, cg->stream);
    if (rc) return rc;
    NEKMF_CUDA(cudaMemcpyAsync(cg->h_red, cg->d_red, 3 * 8, cudaMemcpyDeviceToHost, cg->stream));
    NEKMF_CUDA(cudaStreamSynchronize(cg->stream));
    out[0] = cg->h_red[0];
    out[1] = cg->h_red[1];
    out[2] = cg->h_red[2];
    return NEKMF_OK;
}
} // namespace nekmf

using namespace nekmf;

extern "C" {

int nekmf_cg_create(nekmf_op_t op, nekmf_map_t map, nekmf_exchange_t ex, nekmf_comm_t comm, int nDir,
                    const double *invdiag, const double *ownerMask, nekmf_cg_t *out)
{
    if (!out || !op || !map) { set_error("nekmf_cg_create: null argument"); return NEKMF_ERR_ARG; }
    *out = nullptr;
    if (op->optype != NEKMF_HELMHOLTZ) { set_error("nekmf_cg_create: operator is not Helmholtz"); return NEKMF_ERR_ARG; }
    if ((size_t)op->nElmt * op->nmTot != (size_t)map->nLocal)
    {
        set_error("nekmf_cg_create: map nLocal %d != nElmt*ncoeff %zu", map->nLocal, (size_t)op->nElmt * op->nmTot);
        return NEKMF_ERR_ARG;
    }
    if (nDir < 0 || nDir > map->nGlobal) { set_error("nekmf_cg_create: bad nDir"); return NEKMF_ERR_ARG; }
    nekmf_cg_s *cg = new nekmf_cg_s;
    cg->op = op; cg->map = map; cg->ex = ex; cg->comm = comm;
    cg->nDir = nDir; cg->nGlobal = map->nGlobal; cg->nLocal = map->nLocal; cg->nNonDir = map->nGlobal - nDir;
    const size_t ng = cg->nGlobal + 2, nn = cg->nNonDir + 2, nl = cg->nLocal + 2;
    cudaError_t e = cudaSuccess;
#define ALLOC(p, n) if (e == cudaSuccess) e = cudaMalloc(&cg->p, (n) * 8)
    ALLOC(d_w, ng); ALLOC(d_s, ng); ALLOC(d_p, nn); ALLOC(d_r, nn); ALLOC(d_q, nn);
    ALLOC(d_lin, nl); ALLOC(d_lout, nl); ALLOC(d_x, ng); ALLOC(d_rhs, ng);
    ALLOC(d_part, (size_t)3 * RED_BLOCKS); ALLOC(d_red, 4);
    if (invdiag) { ALLOC(d_invdiag, nn); if (e == cudaSuccess) e = cudaMemcpy(cg->d_invdiag, invdiag, (size_t)cg->nNonDir * 8, cudaMemcpyHostToDevice); }
    if (ownerMask) { ALLOC(d_mask, ng); if (e == cudaSuccess) e = cudaMemcpy(cg->d_mask, ownerMask, (size_t)cg->nGlobal * 8, cudaMemcpyHostToDevice); }
#undef ALLOC
    if (e == cudaSuccess) e = cudaMallocHost(&cg->h_red, 4 * 8);
    if (e != cudaSuccess)
    {
        set_error("nekmf_cg_create: %s", cudaGetErrorString(e));
        nekmf_cg_destroy(cg);
        return NEKMF_ERR_CUDA;
    }
    *out = cg;
    return NEKMF_OK;
}

int nekmf_cg_matvec(nekmf_cg_t cg, const double *w, double *s)
{
    if (!cg || !w || !s) { set_error("nekmf_cg_matvec: null argument"); return NEKMF_ERR_ARG; }
    return cg_matvec_device(cg, w, s);
}

int nekmf_cg_solve(nekmf_cg_t cg, const double *rhs_in, double *x_out, int memkind, double tol, int maxiter,
                   int *iterations, double *final_eps)
{
    if (!cg || !rhs_in || !x_out) { set_error("nekmf_cg_solve: null argument"); return NEKMF_ERR_ARG; }
    const int nDir = cg->nDir, nG = cg->nGlobal, nN = cg->nNonDir;
    cudaStream_t st = cg->stream;
    const double *rhs = rhs_in;
    double *x         = x_out;
    if (memkind == NEKMF_HOST)
    {
        NEKMF_CUDA(cudaMemcpyAsync(cg->d_rhs, rhs_in, (size_t)nG * 8, cudaMemcpyHostToDevice, st));
        NEKMF_CUDA(cudaMemcpyAsync(cg->d_x, x_out, (size_t)nG * 8, cudaMemcpyHostToDevice, st));
        rhs = cg->d_rhs;
        x   = cg->d_x;
    }
    const int T = 256, B = (nN + T - 1) / T;
    const double *mask_nd = cg->d_mask ? cg->d_mask + nDir : nullptr;
    double red[3];
    int rc, its = 0, k = 0;
    double alpha = 0.0, beta = 0.0, rho = 0.0, rho_new, mu, eps, rhs_mag;

    // r = rhs[nDir:], x[nDir:] = 0, w = s = 0
    NEKMF_CUDA(cudaMemcpyAsync(cg->d_r, rhs + nDir, (size_t)nN * 8, cudaMemcpyDeviceToDevice, st));
    NEKMF_CUDA(cudaMemsetAsync(x + nDir, 0, (size_t)nN * 8, st));
    NEKMF_CUDA(cudaMemsetAsync(cg->d_w, 0, (size_t)nG * 8, st));
    NEKMF_CUDA(cudaMemsetAsync(cg->d_s, 0, (size_t)nG * 8, st));
    NEKMF_CUDA(cudaMemsetAsync(cg->d_p, 0, (size_t)nN * 8, st));
    NEKMF_CUDA(cudaMemsetAsync(cg->d_q, 0, (size_t)nN * 8, st));
    // eps = r.r over non-Dirichlet DOFs; rhs magnitude over the whole vector (NekLinSysIter.cpp:128-153)
    rc = cg_dots(cg, cg->d_r, cg->d_r, nullptr, nullptr, nullptr, nullptr, mask_nd, nN, red);
    if (rc) return rc;
    eps = red[0];
    rc  = cg_dots(cg, rhs, rhs, nullptr, nullptr, nullptr, nullptr, cg->d_mask, nG, red);
    if (rc) return rc;
    rhs_mag = red[0] > 1e-6 ? red[0] : 1.0;
    if (!(eps < tol * tol * rhs_mag))
    {
        if (B > 0) cg_precon<<<B, T, 0, st>>>(cg->d_w + nDir, cg->d_r, cg->d_invdiag, nN);
        ++g_launches;
        rc = cg_matvec_device(cg, cg->d_w, cg->d_s);
        if (rc) return rc;
        rc = cg_dots(cg, cg->d_r, cg->d_w + nDir, cg->d_s + nDir, cg->d_w + nDir, nullptr, nullptr, mask_nd, nN, red);
        if (rc) return rc;
        rho = red[0]; mu = red[1];
        beta = 0.0; alpha = rho / mu; its = 1;
        for (;;)
        {
            if (k >= maxiter) break;
            // update + (rho, eps) partials | mat-vec with the gather fused into the operator and the mu
            // partials fused into the assemble | one final reduction + all-reduce + 24-byte D2H
            cg_update_dots<<<RED_BLOCKS, RED_T, 0, st>>>(cg->d_p, cg->d_q, x + nDir, cg->d_r, cg->d_w + nDir,
                                                         cg->d_s + nDir, cg->d_invdiag, mask_nd, alpha, beta, nN,
                                                         cg->d_part);
            ++g_launches;
            rc = cg_matvec_device(cg, cg->d_w, cg->d_s, cg->ex ? nullptr : cg->d_part + RED_BLOCKS);
            if (rc) return rc;
            if (cg->ex)
            {
                // interface contributions arrive after the assemble: s.w needs its own pass
                dot1_partial<<<RED_BLOCKS, RED_T, 0, st>>>(cg->d_s + nDir, cg->d_w + nDir, mask_nd, nN,
                                                           cg->d_part + RED_BLOCKS);
                ++g_launches;
            }
            rc = cg_finish_dots(cg, red);
            if (rc) return rc;
            rho_new = red[0]; mu = red[1]; eps = red[2];
            ++its;
            if (eps < tol * tol * rhs_mag) break;
            beta  = rho_new / rho;
            alpha = rho_new / (mu - rho_new * beta / alpha);
            rho   = rho_new;
            ++k;
        }
    }
    if (memkind == NEKMF_HOST)
    {
        NEKMF_CUDA(cudaMemcpyAsync(x_out, cg->d_x, (size_t)nG * 8, cudaMemcpyDeviceToHost, st));
    }
    NEKMF_CUDA(cudaStreamSynchronize(st));
    if (iterations) *iterations = its;
    if (final_eps) *final_eps = eps;
    return NEKMF_OK;
}

int nekmf_cg_destroy(nekmf_cg_t cg)
{
    if (!cg) return NEKMF_OK;
    cudaFree(cg->d_invdiag); cudaFree(cg->d_mask); cudaFree(cg->d_w); cudaFree(cg->d_s); cudaFree(cg->d_p);
    cudaFree(cg->d_r); cudaFree(cg->d_q); cudaFree(cg->d_lin); cudaFree(cg->d_lout); cudaFree(cg->d_x);
    cudaFree(cg->d_rhs); cudaFree(cg->d_part); cudaFree(cg->d_red);
    if (cg->h_red) cudaFreeHost(cg->h_red);
    delete cg;
    return NEKMF_OK;
}

} // extern "C"
