// cg.cu -- matrix-free preconditioned conjugate gradient on A = Assemble o Helmholtz o GlobalToLocal.
//
// Mirrors NekLinSysIterCG::DoConjugateGradient (LibUtilities/LinearAlgebra/NekLinSysIterCG.cpp:
// 104-265): the single-reduction Demmel/Heath/van der Vorst variant with ONE 3-value reduction
// per iteration, the mat-vec of GlobalLinSysIterativeFull::v_DoMatrixMultiply
// (MultiRegions/GlobalLinSysIterativeFull.cpp:215-257) and a diagonal (or identity)
// preconditioner (MultiRegions/PreconditionerDiagonal.cpp).
//
// Device design: all vectors AND the recurrence scalars (alpha, beta, rho, iteration counters, the
// convergence flag) are resident; the host never waits inside the iteration.  One iteration is
//   cg_update_dots       the four axpy updates + preconditioner + partial sums of r.w and r.r (alpha, beta read
//                        from device memory; a no-op once the convergence flag is up)
//   Helmholtz (gather)   GlobalToLocal fused into the operator's loads (hex_kron.cu GATHER variant:
//                        cp.async indirect loads straight into shared memory); other operator
//                        kernels are preceded by the separate gather kernel
//   assemble_dot         transpose-CSR Assemble + partial sums of s.w; sharded: its first blocks deposit the
//                        partition-interface values with the neighbours (NVLink stores + flag, comm.cu)
//   exchange_finish      sharded only: wait for the neighbours' deposits, rank-ordered add, s.w of the shared DOFs
//   cg_reduce_step       fixed-order reduction of the partials, the cross-rank sum through the peer reduction
//                        windows (or ncclAllReduce + cg_step), then the scalar recurrences and the convergence /
//                        iteration-cap test of the reference, all on the device
// Iterations are captured in CUDA graphs (4 per graph) and launched back to back; the host reads the 4-byte
// state flag one graph behind the launches, so the result is exactly the reference's (the iterations launched
// after convergence do not touch x, r, p, q).  Reductions use a fixed grid and order (deterministic,
// ownership-masked and rank-ordered for multi-rank runs); work buffers are allocated once (the reference
// allocates 2 x nLocal every mat-vec, GlobalLinSysIterativeFull.cpp:225-226).
#include "cg_internal.h"
#include <cmath>
#include <stdlib.h>
#include <string.h>

namespace nekmf
{
constexpr int GRAPH_ITERS = 4;
constexpr int PART_LEN    = 3 * RED_BLOCKS + IF_BLOCKS;

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

// partial sums of (a.b, c.d, e.f) over owned DOFs (flags bit 0; null = all); null pointers skip a product
__global__ void __launch_bounds__(RED_T)
    dot3_partial(const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ c,
                 const double *__restrict__ d, const double *__restrict__ e, const double *__restrict__ f,
                 const unsigned char *__restrict__ flags, int n, double *__restrict__ part)
{
    __shared__ double sh[RED_T / 32];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = blockIdx.x * RED_T + threadIdx.x; i < n; i += RED_BLOCKS * RED_T)
    {
        if (flags && !(flags[i] & 1)) continue;
        if (a) s0 = fma(a[i], b[i], s0);
        if (c) s1 = fma(c[i], d[i], s1);
        if (e) s2 = fma(e[i], f[i], s2);
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
// fixed-order sums of the three partial rows (row 1 also takes the IF_BLOCKS interface partials when with_if).  This
// runs in ONE block at the end of every iteration, i.e. it is serial time: all loads are issued before the first add and
// the three block reductions share one shuffle tree / barrier pair (11 -> 6 us per iteration; per value the order of the
// adds is the one of three separate block_sum calls, so the sums are bit-identical to the plain version).
__device__ __forceinline__ void reduce_rows(const double *__restrict__ part, int with_if, double *sh, double out[3])
{
    constexpr int TRIPS = (RED_BLOCKS + RED_T - 1) / RED_T, IFT = (IF_BLOCKS + RED_T - 1) / RED_T;
    double v[3][TRIPS], vi[IFT];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < TRIPS; ++j)
        {
            const int i = threadIdx.x + j * RED_T;
            v[k][j]     = i < RED_BLOCKS ? part[k * RED_BLOCKS + i] : 0.0;
        }
#pragma unroll
    for (int j = 0; j < IFT; ++j)
    {
        const int i = threadIdx.x + j * RED_T;
        vi[j]       = (with_if && i < IF_BLOCKS) ? part[3 * RED_BLOCKS + i] : 0.0;
    }
    double s[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        s[k] = 0.0;
#pragma unroll
        for (int j = 0; j < TRIPS; ++j)
            if (threadIdx.x + j * RED_T < RED_BLOCKS) s[k] += v[k][j];
    }
    if (with_if)
    {
#pragma unroll
        for (int j = 0; j < IFT; ++j)
            if (threadIdx.x + j * RED_T < IF_BLOCKS) s[1] += vi[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
#pragma unroll
        for (int k = 0; k < 3; ++k) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __shared__ double sh3[3][RED_T / 32];
    if (l == 0)
    {
#pragma unroll
        for (int k = 0; k < 3; ++k) sh3[k][w] = s[k];
    }
    __syncthreads();
    double r[3] = {0.0, 0.0, 0.0};
    if (w == 0)
    {
#pragma unroll
        for (int k = 0; k < 3; ++k) r[k] = l < RED_T / 32 ? sh3[k][l] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1)
        {
#pragma unroll
            for (int k = 0; k < 3; ++k) r[k] += __shfl_xor_sync(0xffffffffu, r[k], o);
        }
    }
    __syncthreads();
    (void)sh;
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
__global__ void __launch_bounds__(RED_T) dot3_final(const double *__restrict__ part, int with_if, double *__restrict__ red)
{
    __shared__ double sh[RED_T / 32];
    double v[3];
    reduce_rows(part, with_if, sh, v);
    if (threadIdx.x == 0) red[0] = v[0], red[1] = v[1], red[2] = v[2];
}

// the scalar part of one iteration (NekLinSysIterCG.cpp:237-262), one thread
__device__ __forceinline__ void cg_scalar_step(CgScal *sc, double rho_new, double mu, double eps)
{
    if (sc->done) return;
    sc->mu  = mu;
    sc->eps = eps;
    sc->its += 1;
    if (eps < sc->tol2 * sc->rhs_mag)
    {
        sc->done = 1;
        return;
    }
    const double beta = rho_new / sc->rho;
    sc->alpha         = rho_new / (mu - rho_new * beta / sc->alpha);
    sc->beta          = beta;
    sc->rho           = rho_new;
    sc->k += 1;
    if (sc->k >= sc->maxiter) sc->done = 2;
}

// Last kernel of an iteration.  WIN: the cross-rank sum goes through the peer reduction windows inside this
// kernel (every rank deposits its three partial sums in every window over NVLink, then adds the slots of its own
// window in rank order); otherwise the sums are already global (one rank) or only reduced locally here and
// all-reduced by NCCL before cg_step_kernel.
template <bool WIN, bool STEP>
__global__ void __launch_bounds__(RED_T)
    cg_reduce_step(const double *__restrict__ part, int with_if, double *__restrict__ red, CgScal *sc,
                   nekmf_redwin *const *__restrict__ peer_win, int me, int nranks, unsigned long long *epoch_ctr, int *err)
{
    __shared__ double sh[RED_T / 32];
    __shared__ double tot[3];
    double v[3];
    reduce_rows(part, with_if, sh, v);
    if (threadIdx.x == 0) tot[0] = v[0], tot[1] = v[1], tot[2] = v[2];
    __syncthreads();
    if (WIN)
    {
        const unsigned long long epoch = *epoch_ctr + 1ull;
        const int par = (int)(epoch & 1ull), t = threadIdx.x;
        nekmf_redwin *mw = peer_win[me];
        if (t < nranks)
        {
            nekmf_redwin *pw = peer_win[t];
            pw->val[par][me][0] = tot[0];
            pw->val[par][me][1] = tot[1];
            pw->val[par][me][2] = tot[2];
            __threadfence_system();
            st_release_sys(&pw->flag[par][me], epoch);
            if (!wait_flag(&mw->flag[par][t], epoch, err)) sc->done = 3;
        }
        __syncthreads();
        if (t < 3)
        {
            double s = 0.0;
            for (int r = 0; r < nranks; ++r) s += ld_relaxed_sys(&mw->val[par][r][t]);
            tot[t] = s;
        }
        __syncthreads();
        if (t == 0) *epoch_ctr = epoch;
    }
    if (threadIdx.x == 0)
    {
        red[0] = tot[0], red[1] = tot[1], red[2] = tot[2];
        if (STEP) cg_scalar_step(sc, tot[0], tot[1], tot[2]);
    }
}
__global__ void cg_step_kernel(const double *__restrict__ red, CgScal *sc) { cg_scalar_step(sc, red[0], red[1], red[2]); }

// p = beta p + w ; q = beta q + s ; x += alpha p ; r -= alpha q ; w = M^-1 r
// (NekLinSysIterCG.cpp:209-220); w,s,x are offset to the first non-Dirichlet DOF.  alpha and beta come from the
// device-resident recurrence state; once it says "done" the kernel changes nothing.
// Fused with the two dot products that do not depend on the next mat-vec:
// rho = r.w and eps = r.r (owned DOFs only).  Fixed grid, grid-stride -> deterministic partial sums in
// part[0..RED_BLOCKS) (rho) and part[2*RED_BLOCKS..) (eps); the s.w partials come from the assemble kernel.
__global__ void __launch_bounds__(RED_T)
    cg_update_dots(double *__restrict__ p, double *__restrict__ q, double *__restrict__ x, double *__restrict__ r,
                   double *__restrict__ w, const double *__restrict__ s, const double *__restrict__ invdiag,
                   const unsigned char *__restrict__ flags, const CgScal *__restrict__ sc, int n,
                   double *__restrict__ part)
{
    __shared__ double sh[RED_T / 32];
    if (sc->done) return;
    const double alpha = sc->alpha, beta = sc->beta;
    double rho = 0.0, eps = 0.0;
    // w is not read: the previous iteration stored w = r * invdiag (or r), which is recomputed from the r and invdiag
    // values this pass loads anyway -- the same product, bit for bit, one vector stream less
    for (int i = blockIdx.x * RED_T + threadIdx.x; i < n; i += RED_BLOCKS * RED_T)
    {
        const double ro = r[i], dinv = invdiag ? invdiag[i] : 1.0;
        const double wo = invdiag ? ro * dinv : ro;
        const double pi = fma(beta, p[i], wo);
        const double qi = fma(beta, q[i], s[i]);
        const double ri = fma(-alpha, qi, ro);
        const double wi = invdiag ? ri * dinv : ri;
        p[i] = pi;
        q[i] = qi;
        x[i] = fma(alpha, pi, x[i]);
        r[i] = ri;
        w[i] = wi;
        if (!flags || (flags[i] & 1))
        {
            rho = fma(ri, wi, rho);
            eps = fma(ri, ri, eps);
        }
    }
    rho = block_sum(rho, sh);
    eps = block_sum(eps, sh);
    if (threadIdx.x == 0)
    {
        part[blockIdx.x]                  = rho;
        part[2 * RED_BLOCKS + blockIdx.x] = eps;
    }
}
// The same pass with U chunks of 256 DOFs per trip: 6 U independent loads in flight per thread instead of six.  Measured on
// 2^20 hex elements (whole iteration): U = 1 2.24 ms, U = 2 2.12 ms.  The partial sums are taken in a different order
// than with U = 1 (still fixed: deterministic), so residual histories differ in the last digits between the variants.
// NEKMF_CG_UPDATE_CHUNKS = 1..4 selects (default 2).
template <int U>
__global__ void __launch_bounds__(RED_T)
    cg_update_dots_u(double *__restrict__ p, double *__restrict__ q, double *__restrict__ x, double *__restrict__ r,
                     double *__restrict__ w, const double *__restrict__ s, const double *__restrict__ invdiag,
                     const unsigned char *__restrict__ flags, const CgScal *__restrict__ sc, int n,
                     double *__restrict__ part)
{
    __shared__ double sh[RED_T / 32];
    if (sc->done) return;
    const double alpha = sc->alpha, beta = sc->beta;
    double rho = 0.0, eps = 0.0;
    for (int i0 = blockIdx.x * (U * RED_T) + threadIdx.x; i0 < n; i0 += RED_BLOCKS * U * RED_T)
    {
        double ro[U], dinv[U], pv[U], qv[U], sv[U], xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const int i = i0 + u * RED_T;
            const bool ok = i < n;
            ro[u]   = ok ? r[i] : 0.0;
            dinv[u] = (ok && invdiag) ? invdiag[i] : 1.0;
            pv[u]   = ok ? p[i] : 0.0;
            qv[u]   = ok ? q[i] : 0.0;
            sv[u]   = ok ? s[i] : 0.0;
            xv[u]   = ok ? x[i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const int i = i0 + u * RED_T;
            if (i < n)
            {
                const double wo = invdiag ? ro[u] * dinv[u] : ro[u];
                const double pi = fma(beta, pv[u], wo), qi = fma(beta, qv[u], sv[u]);
                const double ri = fma(-alpha, qi, ro[u]);
                const double wi = invdiag ? ri * dinv[u] : ri;
                p[i] = pi;
                q[i] = qi;
                x[i] = fma(alpha, pi, xv[u]);
                r[i] = ri;
                w[i] = wi;
                if (!flags || (flags[i] & 1))
                {
                    rho = fma(ri, wi, rho);
                    eps = fma(ri, ri, eps);
                }
            }
        }
    }
    rho = block_sum(rho, sh);
    eps = block_sum(eps, sh);
    if (threadIdx.x == 0)
    {
        part[blockIdx.x]                  = rho;
        part[2 * RED_BLOCKS + blockIdx.x] = eps;
    }
}
__global__ void cg_precon(double *__restrict__ w, const double *__restrict__ r, const double *__restrict__ invdiag, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = invdiag ? r[i] * invdiag[i] : r[i];
}

// s = Assemble(Helmholtz(GlobalToLocal(w))) + interface exchange.  When the operator kernel can gather
// (op->gather_ok) the GlobalToLocal pass is fused into its loads and no local input vector is written.
// mu_part != null: the partial sums of s.w over the owned DOFs of [nDir, nGlobal) are left in
// mu_part[0..RED_BLOCKS) (assemble kernel) and mu_part[2*RED_BLOCKS .. +IF_BLOCKS) (shared DOFs, unpack kernel).
static int cg_matvec_device(nekmf_cg_s *cg, const double *w, double *s, double *mu_part = nullptr)
{
    int rc;
    nekmf_op_s *op = cg->op;
    double *out[3] = {cg->d_lout, cg->d_lout, cg->d_lout};
    op->run_e0     = 0;
    op->run_ne     = op->nElmt;
    op->run_stream = cg->stream;
    // single-rank iteration (mu_part given): launches enqueued after the convergence flag went up return at once (the
    // host reads the flag one graph behind the launches).  Sharded solves keep running them: peers wait on the deposits.
    const bool single = !cg->ex && !(cg->comm && cg->comm->nranks > 1);
    const int *skip   = (mu_part && single) ? &cg->d_scal->done : nullptr;
    if (op->gather_ok)
    {
        const double *in[3] = {w, w, w};
        op->gather_map      = cg->map->d_map;
        op->gather_sign     = cg->map->d_sign;
        op->gather_skip     = skip;
        rc                  = op->launch(op, in, out);
        op->gather_map      = nullptr;
        op->gather_sign     = nullptr;
        op->gather_skip     = nullptr;
    }
    else
    {
        if (!cg->d_lin) NEKMF_CUDA(cudaMalloc(&cg->d_lin, ((size_t)cg->nLocal + 2) * 8)); // first mat-vec, never under capture
        rc = map_g2l_device(cg->map, w, cg->d_lin, cg->stream);
        if (rc) return rc;
        const double *in[3] = {cg->d_lin, cg->d_lin, cg->d_lin};
        rc                  = op->launch(op, in, out);
    }
    if (rc) return rc;
    nekmf_exchange_s *ex = cg->ex && cg->ex->total > 0 ? cg->ex : nullptr;
    if (!mu_part)
    {
        rc = map_assemble_device(cg->map, cg->d_lout, s, cg->stream);
        if (rc) return rc;
        return ex ? exchange_add_device(ex, s, cg->stream) : NEKMF_OK;
    }
    if (cg->nGlobal == 0) return NEKMF_OK;
    rc = map_assemble_dot_device(cg->map, cg->d_lout, s, w, cg->d_flags, cg->nDir, mu_part, ex ? &ex->dev : nullptr,
                                 cg->stream, skip);
    if (rc || !cg->ex) return rc;
    rc = exchange_transport_device(cg->ex, cg->stream);
    if (rc) return rc;
    return exchange_finish_device(cg->ex, s, w, cg->d_flags, cg->nDir, mu_part + 2 * RED_BLOCKS, cg->stream);
}

// three owned-DOF dot products, summed over the ranks, on the host (setup phase only: synchronises)
static int cg_dots(nekmf_cg_s *cg, const double *a, const double *b, const double *c, const double *d, const double *e,
                   const double *f, const unsigned char *flags, int n, double out[3])
{
    dot3_partial<<<RED_BLOCKS, RED_T, 0, cg->stream>>>(a, b, c, d, e, f, flags, n, cg->d_part);
    dot3_final<<<1, RED_T, 0, cg->stream>>>(cg->d_part, 0, cg->d_red);
    g_launches += 2;
    NEKMF_CUDA(cudaGetLastError());
    int rc = comm_allreduce_sum(cg->comm, cg->d_red, 3, cg->stream);
    if (rc) return rc;
    NEKMF_CUDA(cudaMemcpyAsync(cg->h_red, cg->d_red, 3 * 8, cudaMemcpyDeviceToHost, cg->stream));
    NEKMF_CUDA(cudaStreamSynchronize(cg->stream));
    out[0] = cg->h_red[0];
    out[1] = cg->h_red[1];
    out[2] = cg->h_red[2];
    return comm_check_error(cg->comm);
}

// one iteration, enqueued on cg->stream (directly or under stream capture)
static int cg_enqueue_iteration(nekmf_cg_s *cg, double *x)
{
    const int nDir = cg->nDir, nN = cg->nNonDir;
    cudaStream_t st            = cg->stream;
    const unsigned char *fl_nd = cg->d_flags ? cg->d_flags + nDir : nullptr;
    static const int chunks = [] { const char *v = getenv("NEKMF_CG_UPDATE_CHUNKS"); return (v && v[0] >= '1' && v[0] <= '4') ? v[0] - '0' : 2; }();
    auto upd = chunks == 1 ? cg_update_dots : (chunks == 2 ? cg_update_dots_u<2> : (chunks == 3 ? cg_update_dots_u<3> : cg_update_dots_u<4>));
    upd<<<RED_BLOCKS, RED_T, 0, st>>>(cg->d_p, cg->d_q, x + nDir, cg->d_r, cg->d_w + nDir, cg->d_s + nDir, cg->d_invdiag, fl_nd,
                                      cg->d_scal, nN, cg->d_part);
    ++g_launches;
    int rc = cg_matvec_device(cg, cg->d_w, cg->d_s, cg->d_part + RED_BLOCKS);
    if (rc) return rc;
    const int with_if = cg->ex ? 1 : 0;
    nekmf_comm_s *c   = cg->comm;
    if (c && c->nranks > 1 && c->p2p)
        cg_reduce_step<true, true><<<1, RED_T, 0, st>>>(cg->d_part, with_if, cg->d_red, cg->d_scal, c->d_peer_win, c->rank,
                                                       c->nranks, c->d_red_epoch, c->d_err);
    else if (c && c->nranks > 1)
    {
        cg_reduce_step<false, false><<<1, RED_T, 0, st>>>(cg->d_part, with_if, cg->d_red, cg->d_scal, nullptr, 0, 1, nullptr,
                                                         nullptr);
        rc = comm_allreduce_sum(c, cg->d_red, 3, st);
        if (rc) return rc;
        cg_step_kernel<<<1, 1, 0, st>>>(cg->d_red, cg->d_scal);
        ++g_launches;
    }
    else
        cg_reduce_step<false, true><<<1, RED_T, 0, st>>>(cg->d_part, with_if, cg->d_red, cg->d_scal, nullptr, 0, 1, nullptr,
                                                        nullptr);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

static void cg_drop_graphs(nekmf_cg_s *cg)
{
    for (int g = 0; g < 2; ++g)
        if (cg->graph[g])
        {
            cudaGraphExecDestroy(cg->graph[g]);
            cg->graph[g] = nullptr;
        }
}

void cg_invalidate_graphs(nekmf_cg_s *cg)
{
    cg_drop_graphs(cg);
    cg->graph_x = nullptr;
}

// capture nIter iterations into an executable graph
static int cg_capture(nekmf_cg_s *cg, double *x, int nIter, cudaGraphExec_t *out)
{
    cudaGraph_t graph = nullptr;
    NEKMF_CUDA(cudaStreamBeginCapture(cg->stream, cudaStreamCaptureModeThreadLocal));
    const long long l0 = g_launches;
    int rc = NEKMF_OK;
    for (int i = 0; i < nIter && rc == NEKMF_OK; ++i) rc = cg_enqueue_iteration(cg, x);
    g_launches = l0; // capturing launches nothing; replays are counted when the graph is launched
    cudaError_t e = cudaStreamEndCapture(cg->stream, &graph);
    if (rc != NEKMF_OK)
    {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (e != cudaSuccess) { set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(e)); return NEKMF_ERR_CUDA; }
    e = cudaGraphInstantiate(out, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); return NEKMF_ERR_CUDA; }
    return NEKMF_OK;
}

// kernels per iteration (for the launch counter when graphs replay)
static int cg_launches_per_iteration(nekmf_cg_s *cg)
{
    int n = 4;                                                      // update, operator, assemble, reduce(+step)
    if (!cg->op->gather_ok) ++n;                                    // separate GlobalToLocal
    if (cg->ex && cg->ex->total > 0) ++n;                           // interface unpack
    if (cg->comm && cg->comm->nranks > 1 && !cg->comm->p2p) ++n;    // scalar step after ncclAllReduce
    return n;
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
    int rc = op_check_ready(op); // geometry and lambda must be set: the solver launches the operator directly
    if (rc) return rc;
    if ((size_t)op->nElmt * op->nmTot != (size_t)map->nLocal)
    {
        set_error("nekmf_cg_create: map nLocal %d != nElmt*ncoeff %zu", map->nLocal, (size_t)op->nElmt * op->nmTot);
        return NEKMF_ERR_ARG;
    }
    if (nDir < 0 || nDir > map->nGlobal) { set_error("nekmf_cg_create: bad nDir"); return NEKMF_ERR_ARG; }
    if (ex && ex->total > 0 && ex->nGlobal != map->nGlobal)
    {
        set_error("nekmf_cg_create: exchange built for %d global DOFs, map has %d", ex->nGlobal, map->nGlobal);
        return NEKMF_ERR_ARG;
    }
    if (ex && ex->total > 0 && !comm) { set_error("nekmf_cg_create: exchange without communicator"); return NEKMF_ERR_ARG; }
    nekmf_cg_s *cg = new nekmf_cg_s;
    cg->op = op; cg->map = map; cg->ex = ex; cg->comm = comm;
    cg->nDir = nDir; cg->nGlobal = map->nGlobal; cg->nLocal = map->nLocal; cg->nNonDir = map->nGlobal - nDir;
    const size_t ng = cg->nGlobal + 2, nn = cg->nNonDir + 2, nl = cg->nLocal + 2;
    cudaError_t e = cudaSuccess;
#define ALLOC(p, n) if (e == cudaSuccess) e = cudaMalloc(&cg->p, (n) * 8)
    ALLOC(d_w, ng); ALLOC(d_s, ng); ALLOC(d_p, nn); ALLOC(d_r, nn); ALLOC(d_q, nn);
    ALLOC(d_lout, nl); ALLOC(d_x, ng); ALLOC(d_rhs, ng);
    ALLOC(d_part, (size_t)PART_LEN); ALLOC(d_red, 4);
    if (e == cudaSuccess) e = cudaMemset(cg->d_part, 0, (size_t)PART_LEN * 8);
    if (invdiag) { ALLOC(d_invdiag, nn); if (e == cudaSuccess) e = cudaMemcpy(cg->d_invdiag, invdiag, (size_t)cg->nNonDir * 8, cudaMemcpyHostToDevice); }
#undef ALLOC
    if (ownerMask || (ex && ex->total > 0))
    {
        std::vector<unsigned char> fl((size_t)cg->nGlobal + 1, 1);
        if (ownerMask)
            for (int g = 0; g < cg->nGlobal; ++g) fl[g] = ownerMask[g] != 0.0 ? 1 : 0;
        if (ex)
            for (int g : ex->h_uidx) fl[g] |= 2;
        if (e == cudaSuccess) e = cudaMalloc(&cg->d_flags, fl.size());
        if (e == cudaSuccess) e = cudaMemcpy(cg->d_flags, fl.data(), fl.size(), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMalloc(&cg->d_scal, sizeof(CgScal));
    if (e == cudaSuccess) e = cudaMallocHost(&cg->h_red, 4 * 8);
    if (e == cudaSuccess) e = cudaMallocHost(&cg->h_scal, sizeof(CgScal));
    if (e == cudaSuccess) e = cudaMallocHost(&cg->h_done, 2 * sizeof(int));
    // a blocking stream: ordered with the legacy default stream the caller's arrays are produced on, and capturable
    if (e == cudaSuccess) e = cudaStreamCreate(&cg->stream);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&cg->ev[i], cudaEventDisableTiming);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreate(&cg->ev_loop[i]);
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
    int rc = op_check_ready(cg->op);
    if (rc) return rc;
    return cg_matvec_device(cg, w, s);
}

int nekmf_cg_solve(nekmf_cg_t cg, const double *rhs_in, double *x_out, int memkind, double tol, int maxiter,
                   int *iterations, double *final_eps)
{
    if (!cg || !rhs_in || !x_out) { set_error("nekmf_cg_solve: null argument"); return NEKMF_ERR_ARG; }
    int rc = op_check_ready(cg->op); // geometry or lambda may have changed since create
    if (rc) return rc;
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
    const unsigned char *fl_nd = cg->d_flags ? cg->d_flags + nDir : nullptr;
    double red[3];
    CgScal sc;
    memset(&sc, 0, sizeof(sc));
    sc.tol2    = tol * tol;
    sc.maxiter = maxiter;
    cg->loop_iterations = 0;

    // r = rhs[nDir:], x[nDir:] = 0, w = s = 0
    NEKMF_CUDA(cudaMemcpyAsync(cg->d_r, rhs + nDir, (size_t)nN * 8, cudaMemcpyDeviceToDevice, st));
    NEKMF_CUDA(cudaMemsetAsync(x + nDir, 0, (size_t)nN * 8, st));
    NEKMF_CUDA(cudaMemsetAsync(cg->d_w, 0, (size_t)nG * 8, st));
    NEKMF_CUDA(cudaMemsetAsync(cg->d_s, 0, (size_t)nG * 8, st));
    NEKMF_CUDA(cudaMemsetAsync(cg->d_p, 0, (size_t)nN * 8, st));
    NEKMF_CUDA(cudaMemsetAsync(cg->d_q, 0, (size_t)nN * 8, st));
    // eps = r.r over non-Dirichlet DOFs; rhs magnitude over the whole vector (NekLinSysIter.cpp:128-153)
    rc = cg_dots(cg, cg->d_r, cg->d_r, nullptr, nullptr, nullptr, nullptr, fl_nd, nN, red);
    if (rc) return rc;
    sc.eps = red[0];
    rc     = cg_dots(cg, rhs, rhs, nullptr, nullptr, nullptr, nullptr, cg->d_flags, nG, red);
    if (rc) return rc;
    sc.rhs_mag = red[0] > 1e-6 ? red[0] : 1.0;
    bool comm_failed = false;
    if (!(sc.eps < sc.tol2 * sc.rhs_mag))
    {
        if (B > 0) cg_precon<<<B, T, 0, st>>>(cg->d_w + nDir, cg->d_r, cg->d_invdiag, nN);
        ++g_launches;
        rc = cg_matvec_device(cg, cg->d_w, cg->d_s);
        if (rc) return rc;
        rc = cg_dots(cg, cg->d_r, cg->d_w + nDir, cg->d_s + nDir, cg->d_w + nDir, nullptr, nullptr, fl_nd, nN, red);
        if (rc) return rc;
        sc.rho = red[0]; sc.mu = red[1];
        sc.beta = 0.0; sc.alpha = sc.rho / sc.mu; sc.its = 1; sc.k = 0;
        sc.done = maxiter <= 0 ? 2 : 0;
        *cg->h_scal = sc;
        NEKMF_CUDA(cudaMemcpyAsync(cg->d_scal, cg->h_scal, sizeof(CgScal), cudaMemcpyHostToDevice, st));

        static const bool use_graph = [] { const char *v = getenv("NEKMF_CG_GRAPH"); return !(v && v[0] == '0'); }();
        if (use_graph && (cg->graph_x != x || cg->graph_lambda != cg->op->lambda || cg->graph_kernel != cg->op->kname))
        {
            cg_drop_graphs(cg);
            rc = cg_capture(cg, x, GRAPH_ITERS, &cg->graph[0]);
            if (!rc) rc = cg_capture(cg, x, 1, &cg->graph[1]);
            if (rc) { cg_drop_graphs(cg); return rc; }
            cg->graph_x = x; cg->graph_lambda = cg->op->lambda; cg->graph_kernel = cg->op->kname;
        }
        const int per_it = cg_launches_per_iteration(cg);
        int launched = 0, slot = 0;
        bool have_prev = false;
        cg->h_done[0] = cg->h_done[1] = 0;
        NEKMF_CUDA(cudaEventRecord(cg->ev_loop[0], st));
        while (launched < maxiter)
        {
            const int n = maxiter - launched < GRAPH_ITERS ? maxiter - launched : GRAPH_ITERS;
            if (use_graph)
            {
                if (n == GRAPH_ITERS) NEKMF_CUDA(cudaGraphLaunch(cg->graph[0], st));
                else
                    for (int i = 0; i < n; ++i) NEKMF_CUDA(cudaGraphLaunch(cg->graph[1], st));
                g_launches += (long long)n * per_it;
            }
            else
                for (int i = 0; i < n; ++i)
                {
                    rc = cg_enqueue_iteration(cg, x);
                    if (rc) return rc;
                }
            launched += n;
            NEKMF_CUDA(cudaMemcpyAsync(&cg->h_done[slot], &cg->d_scal->done, sizeof(int), cudaMemcpyDeviceToHost, st));
            NEKMF_CUDA(cudaEventRecord(cg->ev[slot], st));
            if (have_prev)
            {
                // the state one block of iterations behind the launches: never stalls the device
                NEKMF_CUDA(cudaEventSynchronize(cg->ev[slot ^ 1]));
                if (cg->h_done[slot ^ 1]) break;
            }
            have_prev = true;
            slot ^= 1;
        }
        NEKMF_CUDA(cudaEventRecord(cg->ev_loop[1], st));
        cg->loop_iterations = launched;
        NEKMF_CUDA(cudaMemcpyAsync(cg->h_scal, cg->d_scal, sizeof(CgScal), cudaMemcpyDeviceToHost, st));
        NEKMF_CUDA(cudaStreamSynchronize(st));
        sc = *cg->h_scal;
        comm_failed = sc.done == 3 || comm_check_error(cg->comm) != NEKMF_OK;
    }
    if (memkind == NEKMF_HOST)
    {
        NEKMF_CUDA(cudaMemcpyAsync(x_out, cg->d_x, (size_t)nG * 8, cudaMemcpyDeviceToHost, st));
    }
    NEKMF_CUDA(cudaStreamSynchronize(st));
    if (iterations) *iterations = sc.its;
    if (final_eps) *final_eps = sc.eps;
    if (comm_failed)
    {
        set_error("nekmf_cg_solve: peer-memory exchange timed out waiting for another rank");
        return NEKMF_ERR_COMM;
    }
    if (sc.done == 2)
    {
        // NekLinSysIterCG.cpp:190-203: ROOTONLY_NEKERROR(efatal, "Exceeded maximum number of iterations")
        set_error("Exceeded maximum number of iterations (CG iterations made = %d, error = %.6e, rhs_mag = %.6e)", sc.its,
                  sqrt(sc.eps / sc.rhs_mag), sqrt(sc.rhs_mag));
        return NEKMF_ERR_NOCONVERGE;
    }
    return NEKMF_OK;
}

int nekmf_cg_last_loop(nekmf_cg_t cg, float *ms, int *iterations)
{
    if (!cg || !ms || !iterations) { set_error("nekmf_cg_last_loop: null argument"); return NEKMF_ERR_ARG; }
    *ms         = -1.0f;
    *iterations = cg->loop_iterations;
    if (cg->loop_iterations <= 0) return NEKMF_OK;
    NEKMF_CUDA(cudaEventSynchronize(cg->ev_loop[1]));
    NEKMF_CUDA(cudaEventElapsedTime(ms, cg->ev_loop[0], cg->ev_loop[1]));
    return NEKMF_OK;
}

int nekmf_cg_destroy(nekmf_cg_t cg)
{
    if (!cg) return NEKMF_OK;
    if (cg->stream) cudaStreamSynchronize(cg->stream);
    cg_drop_graphs(cg);
    cudaFree(cg->d_invdiag); cudaFree(cg->d_flags); cudaFree(cg->d_w); cudaFree(cg->d_s); cudaFree(cg->d_p);
    cudaFree(cg->d_r); cudaFree(cg->d_q); cudaFree(cg->d_lin); cudaFree(cg->d_lout); cudaFree(cg->d_x);
    cudaFree(cg->d_rhs); cudaFree(cg->d_part); cudaFree(cg->d_red); cudaFree(cg->d_scal);
    if (cg->h_red) cudaFreeHost(cg->h_red);
    if (cg->h_scal) cudaFreeHost(cg->h_scal);
    if (cg->h_done) cudaFreeHost(cg->h_done);
    for (int i = 0; i < 2; ++i)
    {
        if (cg->ev[i]) cudaEventDestroy(cg->ev[i]);
        if (cg->ev_loop[i]) cudaEventDestroy(cg->ev_loop[i]);
    }
    if (cg->stream) cudaStreamDestroy(cg->stream);
    delete cg;
    return NEKMF_OK;
}

} // extern "C"
