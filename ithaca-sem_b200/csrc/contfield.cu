// contfield.cu -- ContField::v_HelmSolve as ONE device-resident chain.
//
// The reference's solve of  (L - lambda M) u = f  goes through a string of host-array calls, each an ExpList /
// Collections / AssemblyMap operation (MultiRegions/ContField.cpp:878-945 -> GlobalSolve :516-535 ->
// GlobalLinSysIterativeFull::v_Solve, GlobalLinSysIterativeFull.cpp:110-211):
//
//     wsp    = -IProductWRTBase(f)                          ContField.cpp:894-900
//     tmp    = Helmholtz(inout)                             GeneralMatrixOp on the Dirichlet values + initial guess
//     tmp1   = wsp - tmp                                    GlobalLinSysIterativeFull.cpp:161-181
//     rhs    = Assemble(tmp1)                               :184  (AssemblyMapCG::v_Assemble ends with UniversalAssemble)
//     global = CG(rhs), zero start, Dirichlet entries 0     :187-188
//     inout += GlobalToLocal(global)                        :190-193
//     (no Dirichlet DOF on any rank: rhs = Assemble(wsp), inout = GlobalToLocal(global)   :200-204)
//
// and the solver then evaluates the field with BwdTrans(inout).  With every operator a drop-in on host arrays each
// of those steps would cross PCIe twice; here the forcing goes to the device once, the whole chain runs on the
// solver's stream out of device memory, and the coefficients (and, optionally, the physical values) come back once.
#include "cg_internal.h"

struct nekmf_helmsolve_s
{
    nekmf_cg_s *cg    = nullptr;
    nekmf_op_s *iprod = nullptr, *bwd = nullptr;
    int nLocal = 0, nPhys = 0;
    bool anyDir = true; // nDir summed over the ranks > 0
    double *d_phys = nullptr, *d_coef = nullptr, *d_wsp = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t ph[4] = {nullptr, nullptr, nullptr, nullptr}; // phase boundaries between ev[0] and ev[1]
    bool timed        = false;
};

namespace nekmf
{
// tmp1 = (-wsp) - tmp  (Vmath::Neg then Vmath::Vsub, the reference's rounding); tmp == null: tmp1 = -wsp
__global__ void __launch_bounds__(256)
    rhs_local_kernel(const double *__restrict__ wsp, const double *tmp, double *out, size_t n) // out may alias tmp
{
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        out[i] = tmp ? (-wsp[i]) - tmp[i] : -wsp[i];
}
// inout = GlobalToLocal(global) (+ inout when add)
__global__ void __launch_bounds__(256)
    g2l_add_kernel(const int *__restrict__ map, const double *__restrict__ sign, const double *__restrict__ glob,
                   double *__restrict__ inout, size_t n, int add)
{
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
    {
        double v = __ldg(glob + map[i]);
        if (sign) v *= sign[i];
        inout[i] = add ? v + inout[i] : v;
    }
}
static int op_launch_on(nekmf_op_s *op, const double *in, double *out, cudaStream_t st)
{
    const double *ins[3] = {in, in, in};
    double *outs[3]      = {out, out, out};
    op->run_e0           = 0;
    op->run_ne           = op->nElmt;
    op->run_stream       = st;
    const int rc         = op->launch(op, ins, outs);
    op->run_stream       = op->stream;
    return rc;
}
static int grid_for(size_t n)
{
    size_t b = (n + 255) / 256;
    return (int)(b > (size_t)8 * NUM_SMS ? (size_t)8 * NUM_SMS : (b ? b : 1));
}
} // namespace nekmf

using namespace nekmf;

extern "C" {

int nekmf_helmsolve_create(nekmf_cg_t cg, nekmf_op_t iprod, nekmf_op_t bwd, nekmf_helmsolve_t *out)
{
    if (!out || !cg || !iprod) { set_error("nekmf_helmsolve_create: null argument"); return NEKMF_ERR_ARG; }
    *out = nullptr;
    nekmf_op_s *helm = cg->op;
    if (iprod->optype != NEKMF_IPRODUCTWRTBASE || (bwd && bwd->optype != NEKMF_BWDTRANS))
    {
        set_error("nekmf_helmsolve_create: operators must be IProductWRTBase and BwdTrans");
        return NEKMF_ERR_ARG;
    }
    for (nekmf_op_s *o : {iprod, bwd})
        if (o && (o->nElmt != helm->nElmt || o->nmTot != helm->nmTot || o->nqTot != helm->nqTot || o->shape != helm->shape))
        {
            set_error("nekmf_helmsolve_create: operators describe different collections");
            return NEKMF_ERR_ARG;
        }
    int rc = op_check_ready(iprod);
    if (rc) return rc;
    nekmf_helmsolve_s *hs = new nekmf_helmsolve_s;
    hs->cg = cg; hs->iprod = iprod; hs->bwd = bwd;
    hs->nLocal = cg->nLocal;
    hs->nPhys  = helm->nElmt * helm->nqTot;
    cudaError_t e = cudaMalloc(&hs->d_phys, ((size_t)hs->nPhys + 2) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&hs->d_coef, ((size_t)hs->nLocal + 2) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&hs->d_wsp, ((size_t)hs->nLocal + 2) * 8);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreate(&hs->ev[i]);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&hs->ph[i]);
    if (e != cudaSuccess)
    {
        set_error("nekmf_helmsolve_create: %s", cudaGetErrorString(e));
        nekmf_helmsolve_destroy(hs);
        return NEKMF_ERR_CUDA;
    }
    // nDirTotal of GlobalLinSysIterativeFull.cpp:137-139 (AllReduce over the row communicator): COLLECTIVE
    double nd = (double)cg->nDir;
    e = cudaMemcpyAsync(cg->d_red, &nd, 8, cudaMemcpyHostToDevice, cg->stream);
    if (e == cudaSuccess) rc = comm_allreduce_sum(cg->comm, cg->d_red, 1, cg->stream);
    if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(&nd, cg->d_red, 8, cudaMemcpyDeviceToHost, cg->stream);
    if (e == cudaSuccess && !rc) e = cudaStreamSynchronize(cg->stream);
    if (e != cudaSuccess || rc)
    {
        if (!rc) set_error("nekmf_helmsolve_create: %s", cudaGetErrorString(e));
        nekmf_helmsolve_destroy(hs);
        return rc ? rc : NEKMF_ERR_CUDA;
    }
    hs->anyDir = nd > 0.0;
    *out = hs;
    return NEKMF_OK;
}

int nekmf_helmsolve(nekmf_helmsolve_t hs, const double *forcing, double *inout, double *phys_out, int memkind, double tol,
                    int maxiter, int *iterations, double *final_eps)
{
    if (!hs || !forcing || !inout) { set_error("nekmf_helmsolve: null argument"); return NEKMF_ERR_ARG; }
    if (phys_out && !hs->bwd) { set_error("nekmf_helmsolve: phys_out given but no BwdTrans operator"); return NEKMF_ERR_ARG; }
    if (memkind != NEKMF_HOST && memkind != NEKMF_DEVICE) { set_error("nekmf_helmsolve: bad memkind"); return NEKMF_ERR_ARG; }
    nekmf_cg_s *cg   = hs->cg;
    nekmf_op_s *helm = cg->op;
    int rc = op_check_ready(helm);
    if (!rc) rc = op_check_ready(hs->iprod);
    if (!rc && hs->bwd && phys_out) rc = op_check_ready(hs->bwd);
    if (rc) return rc;
    cudaStream_t st  = cg->stream;
    const size_t nL = (size_t)hs->nLocal, nP = (size_t)hs->nPhys;
    const bool host = memkind == NEKMF_HOST;
    const double *f = forcing;
    double *coef    = inout;
    NEKMF_CUDA(cudaEventRecord(hs->ev[0], st));
    if (host)
    {
        NEKMF_CUDA(cudaMemcpyAsync(hs->d_phys, forcing, nP * 8, cudaMemcpyHostToDevice, st));
        NEKMF_CUDA(cudaMemcpyAsync(hs->d_coef, inout, nL * 8, cudaMemcpyHostToDevice, st));
        f    = hs->d_phys;
        coef = hs->d_coef;
    }
    NEKMF_CUDA(cudaEventRecord(hs->ph[0], st));
    if (helm->nElmt > 0)
    {
        rc = op_launch_on(hs->iprod, f, hs->d_wsp, st);
        if (rc) return rc;
        if (hs->anyDir)
        {
            rc = op_launch_on(helm, coef, cg->d_lout, st);
            if (rc) return rc;
        }
        rhs_local_kernel<<<grid_for(nL), 256, 0, st>>>(hs->d_wsp, hs->anyDir ? cg->d_lout : nullptr, cg->d_lout, nL);
        ++g_launches;
        NEKMF_CUDA(cudaGetLastError());
    }
    rc = map_assemble_device(cg->map, cg->d_lout, cg->d_rhs, st);
    if (!rc && cg->ex) rc = exchange_add_device(cg->ex, cg->d_rhs, st);
    if (rc) return rc;
    NEKMF_CUDA(cudaMemsetAsync(cg->d_x, 0, (size_t)cg->nGlobal * 8, st));
    NEKMF_CUDA(cudaEventRecord(hs->ph[1], st));
    int its = 0;
    double eps = 0.0;
    const int src = nekmf_cg_solve(cg, cg->d_rhs, cg->d_x, NEKMF_DEVICE, tol, maxiter, &its, &eps);
    if (iterations) *iterations = its;
    if (final_eps) *final_eps = eps;
    if (src != NEKMF_OK && src != NEKMF_ERR_NOCONVERGE) return src;
    NEKMF_CUDA(cudaEventRecord(hs->ph[2], st));
    if (nL > 0)
    {
        g2l_add_kernel<<<grid_for(nL), 256, 0, st>>>(cg->map->d_map, cg->map->d_sign, cg->d_x, coef, nL, hs->anyDir ? 1 : 0);
        ++g_launches;
        NEKMF_CUDA(cudaGetLastError());
    }
    double *phys = host ? hs->d_phys : phys_out;
    if (phys_out && helm->nElmt > 0)
    {
        rc = op_launch_on(hs->bwd, coef, phys, st);
        if (rc) return rc;
    }
    NEKMF_CUDA(cudaEventRecord(hs->ph[3], st));
    if (host)
    {
        NEKMF_CUDA(cudaMemcpyAsync(inout, hs->d_coef, nL * 8, cudaMemcpyDeviceToHost, st));
        if (phys_out) NEKMF_CUDA(cudaMemcpyAsync(phys_out, hs->d_phys, nP * 8, cudaMemcpyDeviceToHost, st));
    }
    NEKMF_CUDA(cudaEventRecord(hs->ev[1], st));
    hs->timed = true;
    NEKMF_CUDA(cudaStreamSynchronize(st));
    return src; // NEKMF_OK, or NEKMF_ERR_NOCONVERGE with the outputs of the capped solve in place
}

int nekmf_helmsolve_last_ms(nekmf_helmsolve_t hs, float *ms)
{
    if (!hs || !ms) { set_error("nekmf_helmsolve_last_ms: null argument"); return NEKMF_ERR_ARG; }
    *ms = -1.0f;
    if (!hs->timed) return NEKMF_OK;
    NEKMF_CUDA(cudaEventSynchronize(hs->ev[1]));
    NEKMF_CUDA(cudaEventElapsedTime(ms, hs->ev[0], hs->ev[1]));
    return NEKMF_OK;
}

int nekmf_helmsolve_last_phases(nekmf_helmsolve_t hs, float ms[5])
{
    if (!hs || !ms) { set_error("nekmf_helmsolve_last_phases: null argument"); return NEKMF_ERR_ARG; }
    for (int i = 0; i < 5; ++i) ms[i] = -1.0f;
    if (!hs->timed) return NEKMF_OK;
    NEKMF_CUDA(cudaEventSynchronize(hs->ev[1]));
    cudaEvent_t seq[6] = {hs->ev[0], hs->ph[0], hs->ph[1], hs->ph[2], hs->ph[3], hs->ev[1]};
    for (int i = 0; i < 5; ++i) NEKMF_CUDA(cudaEventElapsedTime(&ms[i], seq[i], seq[i + 1]));
    return NEKMF_OK;
}

int nekmf_helmsolve_destroy(nekmf_helmsolve_t hs)
{
    if (!hs) return NEKMF_OK;
    if (hs->cg && hs->cg->stream) cudaStreamSynchronize(hs->cg->stream);
    cudaFree(hs->d_phys);
    cudaFree(hs->d_coef);
    cudaFree(hs->d_wsp);
    for (int i = 0; i < 2; ++i)
        if (hs->ev[i]) cudaEventDestroy(hs->ev[i]);
    for (int i = 0; i < 4; ++i)
        if (hs->ph[i]) cudaEventDestroy(hs->ph[i]);
    delete hs;
    return NEKMF_OK;
}

} // extern "C"
