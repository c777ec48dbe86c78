// abi.cu -- the C ABI of include/nekmf_b200.h: library, memory helpers and the elemental
// operator object (create / set_geom / set_lambda / apply / destroy).
#include <stdlib.h>
#include "op_internal.h"
#include <stdarg.h>
#include <string.h>

namespace nekmf
{
static thread_local char g_err[512] = "";
long long g_launches                = 0;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int basis_rows(int btype, int nm)
{
    switch (btype)
    {
        case NEKMF_MODIFIED_A: return nm;
        case NEKMF_MODIFIED_B: return nm * (nm + 1) / 2;
        case NEKMF_MODIFIED_C: return nm * (nm + 1) * (nm + 2) / 6;
        case NEKMF_MODIFIEDPYR_C: return nm * (nm + 1) * (2 * nm + 1) / 6;
    }
    return -1;
}

// number of modes per element, LibUtilities/BasicUtils/ShapeType.hpp:111-337
static int count_modes(int shape, int nm)
{
    switch (shape)
    {
        case NEKMF_QUAD: return nm * nm;
        case NEKMF_TRI: return nm * (nm + 1) / 2;
        case NEKMF_HEX: return nm * nm * nm;
        case NEKMF_PRISM: return nm * nm * (nm + 1) / 2;
        case NEKMF_TET: return nm * (nm + 1) * (nm + 2) / 6;
        case NEKMF_PYR: return nm * (nm + 1) * (2 * nm + 1) / 6;
        case NEKMF_SEG: return nm;
    }
    return -1;
}
} // namespace nekmf

using namespace nekmf;

extern "C" {

int nekmf_abi_version(void) { return NEKMF_ABI_VERSION; }
const char *nekmf_last_error(void) { return g_err; }

int nekmf_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int nekmf_set_device(int device)
{
    NEKMF_CUDA(cudaSetDevice(device));
    return NEKMF_OK;
}

long long nekmf_launch_count(void) { return g_launches; }

int nekmf_malloc_device(void **ptr, size_t bytes)
{
    if (!ptr) { set_error("nekmf_malloc_device: null ptr"); return NEKMF_ERR_ARG; }
    NEKMF_CUDA(cudaMalloc(ptr, bytes ? bytes : 8));
    return NEKMF_OK;
}
int nekmf_free_device(void *ptr)
{
    NEKMF_CUDA(cudaFree(ptr));
    return NEKMF_OK;
}
int nekmf_malloc_pinned(void **ptr, size_t bytes)
{
    if (!ptr) { set_error("nekmf_malloc_pinned: null ptr"); return NEKMF_ERR_ARG; }
    NEKMF_CUDA(cudaMallocHost(ptr, bytes ? bytes : 8));
    return NEKMF_OK;
}
int nekmf_free_pinned(void *ptr)
{
    NEKMF_CUDA(cudaFreeHost(ptr));
    return NEKMF_OK;
}
int nekmf_host_register(void *ptr, size_t bytes)
{
    if (!ptr) { set_error("nekmf_host_register: null ptr"); return NEKMF_ERR_ARG; }
    NEKMF_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return NEKMF_OK;
}
int nekmf_host_unregister(void *ptr)
{
    NEKMF_CUDA(cudaHostUnregister(ptr));
    return NEKMF_OK;
}
int nekmf_memcpy_h2d(void *dst, const void *src, size_t bytes)
{
    NEKMF_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return NEKMF_OK;
}
int nekmf_memcpy_d2h(void *dst, const void *src, size_t bytes)
{
    NEKMF_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return NEKMF_OK;
}
int nekmf_memset_device(void *dst, int value, size_t bytes)
{
    NEKMF_CUDA(cudaMemset(dst, value, bytes));
    return NEKMF_OK;
}
int nekmf_sync(void)
{
    NEKMF_CUDA(cudaDeviceSynchronize());
    return NEKMF_OK;
}

// ------------------------------------------------------------------------------ operators

int nekmf_op_create(int shape, int optype, const int nm[3], const int nq[3], const int basistype[3],
                    const int pointstype[3], const double *const bdata[3], const double *const dbdata[3],
                    const double *const D[3], const double *const Z[3], const double *const W[3], int nElmt,
                    int deformed, int coordim, nekmf_op_t *out)
{
    if (!out) { set_error("nekmf_op_create: null output handle"); return NEKMF_ERR_ARG; }
    *out = nullptr;
    if (!nm || !nq || !basistype || !pointstype || !bdata || !dbdata || !D || !Z || !W)
    {
        set_error("nekmf_op_create: null table argument");
        return NEKMF_ERR_ARG;
    }
    if (shape != NEKMF_QUAD && shape != NEKMF_TRI && shape != NEKMF_HEX && shape != NEKMF_PRISM && shape != NEKMF_PYR &&
        shape != NEKMF_TET && shape != NEKMF_SEG)
    {
        set_error("nekmf_op_create: unknown shape %d", shape);
        return NEKMF_ERR_ARG;
    }
    if (optype < NEKMF_BWDTRANS || optype > NEKMF_PHYSDERIV)
    {
        set_error("nekmf_op_create: unknown operator type %d", optype);
        return NEKMF_ERR_ARG;
    }
    if (nElmt < 0) { set_error("nekmf_op_create: negative element count"); return NEKMF_ERR_ARG; }
    const int dim = shape == NEKMF_SEG ? 1 : ((shape == NEKMF_QUAD || shape == NEKMF_TRI) ? 2 : 3);
    if (shape == NEKMF_SEG)
    {
        if (coordim < 1 || coordim > 3)
        {
            set_error("nekmf_op_create: segment coordim %d out of range 1..3", coordim);
            return NEKMF_ERR_ARG;
        }
        if (optype == NEKMF_HELMHOLTZ)
        {
            // the reference registers no (eSegment, eHelmholtz, eMatrixFree) operator (Collections/Helmholtz.cpp:484-504)
            set_error("nekmf_op_create: Helmholtz has no segment variant");
            return NEKMF_ERR_UNSUPPORTED;
        }
        if (optype == NEKMF_IPRODUCTWRTDERIVBASE && coordim == 3 && !deformed)
        {
            // the reference's regular coordim-3 branch multiplies the third input by df[1] (IProductWRTDerivBase.h:321-323);
            // neither that nor a silently different result is offered
            set_error("nekmf_op_create: IProductWRTDerivBase on regular segments in 3 space dimensions is not supported");
            return NEKMF_ERR_UNSUPPORTED;
        }
    }
    else if (coordim != dim)
    {
        // the reference's 2-D kernels reject a third output (PhysDeriv.h:285-286,412-413)
        set_error("nekmf_op_create: coordim %d != element dimension %d is not supported", coordim, dim);
        return NEKMF_ERR_UNSUPPORTED;
    }
    // expected basis / points per direction (SpatialDomains/MeshGraph.cpp:1609-1762) and the
    // isotropy preconditions asserted by the reference (Helmholtz.h:42-44,329-331,1042-1046,1519-1525,2017-2021)
    int ebt[3] = {NEKMF_MODIFIED_A, NEKMF_MODIFIED_A, NEKMF_MODIFIED_A};
    int ept[3] = {NEKMF_GLL, NEKMF_GLL, NEKMF_GLL};
    int enq[3] = {nq[0], nq[0], nq[0]};
    switch (shape)
    {
        case NEKMF_TRI: ebt[1] = NEKMF_MODIFIED_B; ept[1] = NEKMF_GRJM_A1B0; enq[1] = nq[0] - 1; break;
        case NEKMF_PRISM: ebt[2] = NEKMF_MODIFIED_B; ept[2] = NEKMF_GRJM_A1B0; enq[2] = nq[0] - 1; break;
        case NEKMF_PYR: ebt[2] = NEKMF_MODIFIEDPYR_C; ept[2] = NEKMF_GRJM_A2B0; enq[2] = nq[0] - 1; break;
        case NEKMF_TET:
            ebt[1] = NEKMF_MODIFIED_B; ept[1] = NEKMF_GRJM_A1B0; enq[1] = nq[0] - 1;
            ebt[2] = NEKMF_MODIFIED_C; ept[2] = NEKMF_GRJM_A2B0; enq[2] = nq[0] - 1;
            break;
        default: break;
    }
    for (int d = 0; d < dim; ++d)
    {
        if (nm[d] != nm[0] || nm[d] < 2)
        {
            set_error("nekmf_op_create: modes must be isotropic and >= 2 (got nm[%d]=%d, nm[0]=%d)", d, nm[d], nm[0]);
            return NEKMF_ERR_UNSUPPORTED;
        }
        if (nq[d] != enq[d] || nq[d] < 2)
        {
            set_error("nekmf_op_create: quadrature order nq[%d]=%d not supported for this shape (expected %d)", d,
                      nq[d], enq[d]);
            return NEKMF_ERR_UNSUPPORTED;
        }
        if (basistype[d] != ebt[d] || pointstype[d] != ept[d])
        {
            set_error("nekmf_op_create: basis/points type in direction %d not supported (got %d/%d, expected %d/%d)", d,
                      basistype[d], pointstype[d], ebt[d], ept[d]);
            return NEKMF_ERR_UNSUPPORTED;
        }
        if (!bdata[d] || !dbdata[d] || !D[d] || !Z[d] || !W[d])
        {
            set_error("nekmf_op_create: null table in direction %d", d);
            return NEKMF_ERR_ARG;
        }
    }
    if (nm[0] > 16 || nq[0] > 24)
    {
        set_error("nekmf_op_create: order too large (nm=%d nq=%d)", nm[0], nq[0]);
        return NEKMF_ERR_UNSUPPORTED;
    }

    nekmf_op_s *op = new nekmf_op_s;
    op->shape    = shape;
    op->optype   = optype;
    op->dim      = dim;
    op->coordim  = coordim;
    op->nElmt    = nElmt;
    op->deformed = deformed ? 1 : 0;
    op->ndf      = dim * coordim;
    op->nmTot    = count_modes(shape, nm[0]);
    op->nqTot    = 1;
    int off      = 0;
    for (int d = 0; d < dim; ++d)
    {
        op->nm[d]    = nm[d];
        op->nq[d]    = nq[d];
        op->btype[d] = basistype[d];
        op->ptype[d] = pointstype[d];
        op->rows[d]  = basis_rows(basistype[d], nm[d]);
        op->nqTot *= nq[d];
        const int nb = op->rows[d] * nq[d];
        op->b[d].assign(bdata[d], bdata[d] + nb);
        op->db[d].assign(dbdata[d], dbdata[d] + nb);
        op->D[d].assign(D[d], D[d] + nq[d] * nq[d]);
        op->Z[d].assign(Z[d], Z[d] + nq[d]);
        op->W[d].assign(W[d], W[d] + nq[d]);
        // Helper<DIM>: collapsed-coordinate Jacobian folded into the weights, Operator.hpp:244-258
        const double fac = pointstype[d] == NEKMF_GRJM_A1B0 ? 0.5 : (pointstype[d] == NEKMF_GRJM_A2B0 ? 0.25 : 1.0);
        op->ws[d].resize(nq[d]);
        for (int i = 0; i < nq[d]; ++i) op->ws[d][i] = fac * W[d][i];
        const int lens[5] = {nb, nb, nq[d] * nq[d], nq[d], nq[d]};
        for (int t = 0; t < 5; ++t)
        {
            op->tab_off[d][t] = off;
            off += (lens[t] + 1) & ~1; // keep every table 16-byte aligned
        }
    }
    op->tab_len = off;

    if (nekmf_device_count() < 1)
    {
        delete op;
        set_error("nekmf_op_create: no CUDA device (this library has no CPU fallback)");
        return NEKMF_ERR_CUDA;
    }
    // packed device tables for the runtime-sized kernels
    {
        std::vector<double> h(off > 0 ? off : 1, 0.0);
        for (int d = 0; d < dim; ++d)
        {
            const std::vector<double> *src[5] = {&op->b[d], &op->db[d], &op->D[d], &op->Z[d], &op->ws[d]};
            for (int t = 0; t < 5; ++t) memcpy(h.data() + op->tab_off[d][t], src[t]->data(), src[t]->size() * 8);
        }
        cudaError_t e = cudaMalloc(&op->d_tab, h.size() * 8);
        if (e == cudaSuccess) e = cudaMemcpy(op->d_tab, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess)
        {
            set_error("nekmf_op_create: table upload failed: %s", cudaGetErrorString(e));
            delete op;
            return NEKMF_ERR_CUDA;
        }
    }
    bool ok       = false;
    op->geo_pitch = op->nqTot;
    if (shape == NEKMF_SEG) ok = select_seg(op);
    if (!ok && shape == NEKMF_HEX) ok = select_hex_fast(op);
    if (!ok && shape != NEKMF_HEX && shape != NEKMF_SEG) ok = select_shape_fast(op);
    if (!ok && shape != NEKMF_SEG) ok = select_generic(op);
    if (!ok)
    {
        set_error("nekmf_op_create: no kernel for shape %d op %d nm %d nq %d", shape, optype, nm[0], nq[0]);
        nekmf_op_destroy(op);
        return NEKMF_ERR_UNSUPPORTED;
    }
    prism_maybe_wrap(op); // regular prism Helmholtz: DMMA kernel for extruded elements on top of the selected one
    dense_maybe_wrap(op); // regular Tri / Tet / Pyr Helmholtz: DMMA coefficient-space kernel on top of the selected one
    pyr_dmma_maybe_wrap(op); // pyramid BwdTrans / IProductWRTBase: tensor-core kernels on top of the runtime-sized one
    *out = op;
    return NEKMF_OK;
}

int nekmf_op_set_geom(nekmf_op_t op, const double *jac, const double *df, int memkind)
{
    if (!op) { set_error("nekmf_op_set_geom: null operator"); return NEKMF_ERR_ARG; }
    const cudaMemcpyKind k = memkind == NEKMF_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    // rows of `width` doubles in the caller's layout, stored with pitch `pitch` on the device
    const size_t width = op->deformed ? (size_t)op->nqTot : 1;
    const size_t pitch = op->deformed ? (size_t)op->geo_pitch : 1;
    const size_t rows  = (size_t)op->nElmt;
    if (jac)
    {
        if (!op->d_jac)
        {
            NEKMF_CUDA(cudaMalloc(&op->d_jac, (rows * pitch + 2) * 8));
            NEKMF_CUDA(cudaMemset(op->d_jac, 0, (rows * pitch + 2) * 8));
        }
        if (rows && pitch == width) NEKMF_CUDA(cudaMemcpy(op->d_jac, jac, rows * width * 8, k));
        else if (rows) NEKMF_CUDA(cudaMemcpy2D(op->d_jac, pitch * 8, jac, width * 8, width * 8, rows, k));
        op->has_jac = true;
    }
    if (df)
    {
        if (!op->d_df)
        {
            NEKMF_CUDA(cudaMalloc(&op->d_df, (rows * pitch * op->ndf + 2) * 8));
            NEKMF_CUDA(cudaMemset(op->d_df, 0, (rows * pitch * op->ndf + 2) * 8));
        }
        if (rows && pitch == width) NEKMF_CUDA(cudaMemcpy(op->d_df, df, rows * width * 8 * op->ndf, k));
        else if (rows) NEKMF_CUDA(cudaMemcpy2D(op->d_df, pitch * 8, df, width * 8, width * 8, rows * op->ndf, k));
        op->has_df = true;
    }
    return notify_geom_changed(op);
}

int nekmf_op_set_lambda(nekmf_op_t op, double lambda)
{
    if (!op) { set_error("nekmf_op_set_lambda: null operator"); return NEKMF_ERR_ARG; }
    op->lambda     = lambda;
    op->lambda_set = true;
    return NEKMF_OK;
}

int nekmf_op_set_stream(nekmf_op_t op, void *stream)
{
    if (!op) { set_error("nekmf_op_set_stream: null operator"); return NEKMF_ERR_ARG; }
    op->stream = static_cast<cudaStream_t>(stream);
    return NEKMF_OK;
}

int nekmf_op_ncoeff(nekmf_op_t op) { return op ? op->nmTot : -1; }
int nekmf_op_nphys(nekmf_op_t op) { return op ? op->nqTot : -1; }
const char *nekmf_op_kernel_name(nekmf_op_t op) { return op ? op->kname.c_str() : ""; }

int nekmf_op_enable_timing(nekmf_op_t op, int on)
{
    if (!op) { set_error("nekmf_op_enable_timing: null operator"); return NEKMF_ERR_ARG; }
    if (on && !op->ev0)
    {
        NEKMF_CUDA(cudaEventCreate(&op->ev0));
        NEKMF_CUDA(cudaEventCreate(&op->ev1));
    }
    op->timing = on != 0;
    return NEKMF_OK;
}

int nekmf_op_last_ms(nekmf_op_t op, float *ms)
{
    if (!op || !ms) { set_error("nekmf_op_last_ms: null argument"); return NEKMF_ERR_ARG; }
    if (!op->timing || !op->timed_once)
    {
        *ms = -1.0f;
        return NEKMF_OK;
    }
    NEKMF_CUDA(cudaEventSynchronize(op->ev1));
    NEKMF_CUDA(cudaEventElapsedTime(ms, op->ev0, op->ev1));
    return NEKMF_OK;
}

} // extern "C"
namespace nekmf
{
int op_check_ready(nekmf_op_s *op)
{
    const bool need_jac = op->optype == NEKMF_HELMHOLTZ || op->optype == NEKMF_IPRODUCTWRTBASE ||
                          op->optype == NEKMF_IPRODUCTWRTDERIVBASE;
    const bool need_df = op->optype == NEKMF_HELMHOLTZ || op->optype == NEKMF_PHYSDERIV ||
                         op->optype == NEKMF_IPRODUCTWRTDERIVBASE;
    if ((need_jac && !op->has_jac) || (need_df && !op->has_df))
    {
        set_error("nekmf_op_apply: geometric factors not set (SetJac/SetDF) for operator type %d", op->optype);
        return NEKMF_ERR_STATE;
    }
    if (op->optype == NEKMF_HELMHOLTZ && !op->lambda_set)
    {
        set_error("nekmf_op_apply: Helmholtz lambda not set (SetLambda)");
        return NEKMF_ERR_STATE;
    }
    return NEKMF_OK;
}

} // namespace nekmf
extern "C" {

// Device-side aliases of page-locked host arrays (cudaHostAlloc / cudaHostRegister); false if any array is
// pageable or not mapped into the device address space.
static bool host_arrays_device_pointers(const double *const ins[3], int nin, double *const outs[3], int nout,
                                        const double *din[3], double *dout[3])
{
    auto devptr = [](const void *p) -> void * {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
    };
    for (int a = 0; a < nin; ++a)
        if (!(din[a] = static_cast<const double *>(devptr(ins[a])))) return false;
    for (int a = 0; a < nout; ++a)
        if (!(dout[a] = static_cast<double *>(devptr(outs[a])))) return false;
    return true;
}

int nekmf_op_apply(nekmf_op_t op, const double *in0, const double *in1, const double *in2, double *out0,
                   double *out1, double *out2, int memkind)
{
    if (!op) { set_error("nekmf_op_apply: null operator"); return NEKMF_ERR_ARG; }
    int rc = op_check_ready(op);
    if (rc) return rc;
    const int nin   = op->optype == NEKMF_IPRODUCTWRTDERIVBASE ? op->coordim : 1;
    const int nout  = op->optype == NEKMF_PHYSDERIV ? op->coordim : 1;
    const bool cin  = op->optype == NEKMF_BWDTRANS || op->optype == NEKMF_HELMHOLTZ;
    const bool cout = op->optype != NEKMF_BWDTRANS && op->optype != NEKMF_PHYSDERIV;
    const size_t in_sz  = (size_t)op->nElmt * (cin ? op->nmTot : op->nqTot);
    const size_t out_sz = (size_t)op->nElmt * (cout ? op->nmTot : op->nqTot);
    const double *ins[3] = {in0, in1, in2};
    double *outs[3]      = {out0, out1, out2};
    for (int a = 0; a < nin; ++a)
        if (!ins[a]) { set_error("nekmf_op_apply: input %d is null", a); return NEKMF_ERR_ARG; }
    for (int a = 0; a < nout; ++a)
        if (!outs[a]) { set_error("nekmf_op_apply: output %d is null", a); return NEKMF_ERR_ARG; }
    if (op->nElmt == 0) return NEKMF_OK;

    if (memkind == NEKMF_DEVICE)
    {
        const double *din[3] = {ins[0], nin > 1 ? ins[1] : ins[0], nin > 2 ? ins[2] : ins[0]};
        double *dout[3]      = {outs[0], nout > 1 ? outs[1] : outs[0], nout > 2 ? outs[2] : outs[0]};
        op->run_e0     = 0;
        op->run_ne     = op->nElmt;
        op->run_stream = op->stream;
        if (op->timing) NEKMF_CUDA(cudaEventRecord(op->ev0, op->stream));
        rc = op->launch(op, din, dout);
        if (rc) return rc;
        if (op->timing)
        {
            NEKMF_CUDA(cudaEventRecord(op->ev1, op->stream));
            op->timed_once = true;
        }
        return NEKMF_OK;
    }
    if (memkind != NEKMF_HOST) { set_error("nekmf_op_apply: bad memkind %d", memkind); return NEKMF_ERR_ARG; }

    // Host arrays (the literal Array<OneD> drop-in).  The collection is cut into element chunks that flow
    // through three role streams: every H2D copy is queued back to back on the copy-in stream, the kernel on
    // chunk c waits (event) for its copy, the D2H copy of chunk c waits for its kernel.  The staging buffers
    // hold the whole collection, so no copy ever waits for a later stage: both PCIe directions stay busy and
    // the call costs about max(H2D, D2H) of a full-duplex link instead of H2D + kernel + D2H.
    // Synchronous: returns when `out` is complete.
    // Page-locked host arrays are addressable from the device (unified virtual addressing), so the kernels can
    // read their inputs from and write their results to host memory themselves: one launch over the whole
    // collection, no staging memory.  Opt-in (NEKMF_HOST_ZEROCOPY=1) for callers short of device memory:
    // measured on the 64^3 P=4 bench it is slower than the staged pipeline below (6.8 vs 6.3 ms; letting only
    // the outputs go direct while inputs use the copy engine is worse still, 7.3+ ms, the SM-issued PCIe
    // writes and the DMA reads throttle each other).
    const char *zc_env   = getenv("NEKMF_HOST_ZEROCOPY"); // read per call: a caller may switch it at run time
    const bool zero_copy = zc_env && atoi(zc_env) > 0;
    const double *zin[3] = {nullptr, nullptr, nullptr};
    double *zout[3]      = {nullptr, nullptr, nullptr};
    if (zero_copy && host_arrays_device_pointers(ins, nin, outs, nout, zin, zout))
    {
        for (int a = nin; a < 3; ++a) zin[a] = zin[0];
        for (int a = nout; a < 3; ++a) zout[a] = zout[0];
        op->run_e0     = 0;
        op->run_ne     = op->nElmt;
        op->run_stream = op->stream;
        if (op->timing) NEKMF_CUDA(cudaEventRecord(op->ev0, op->stream));
        rc = op->launch(op, zin, zout);
        if (rc) return rc;
        if (op->timing)
        {
            NEKMF_CUDA(cudaEventRecord(op->ev1, op->stream));
            op->timed_once = true;
        }
        NEKMF_CUDA(cudaStreamSynchronize(op->stream));
        return NEKMF_OK;
    }
    if (op->stage_in_sz < in_sz * nin)
    {
        if (op->d_stage_in) cudaFree(op->d_stage_in);
        op->d_stage_in = nullptr;
        NEKMF_CUDA(cudaMalloc(&op->d_stage_in, in_sz * nin * 8));
        op->stage_in_sz = in_sz * nin;
    }
    if (op->stage_out_sz < out_sz * nout)
    {
        if (op->d_stage_out) cudaFree(op->d_stage_out);
        op->d_stage_out = nullptr;
        NEKMF_CUDA(cudaMalloc(&op->d_stage_out, out_sz * nout * 8));
        op->stage_out_sz = out_sz * nout;
    }
    if (!op->pipe_stream[0])
    {
        for (int s = 0; s < 3; ++s) NEKMF_CUDA(cudaStreamCreateWithFlags(&op->pipe_stream[s], cudaStreamNonBlocking));
        NEKMF_CUDA(cudaEventCreateWithFlags(&op->pipe_done, cudaEventDisableTiming));
    }
    const size_t in_el  = cin ? op->nmTot : op->nqTot;
    const size_t out_el = cout ? op->nmTot : op->nqTot;
    const size_t el_bytes = 8 * (in_el * nin > out_el * nout ? in_el * nin : out_el * nout);
    static const size_t chunk_bytes = [] {
        const char *v = getenv("NEKMF_HOST_CHUNK_MB"); // tuning knob
        const long mb = v ? atol(v) : 0;
        return (size_t)(mb > 0 ? mb : 16) << 20;
    }();
    cudaStream_t sH = op->pipe_stream[0], sK = op->pipe_stream[1], sD = op->pipe_stream[2];

    // Chunk schedule: the first and the last chunks are small (the first H2D and the last D2H copy overlap
    // nothing, and the D2H stream trails the H2D stream by one chunk), sizes double towards the middle up to
    // NEKMF_HOST_CHUNK_MB (default 16 MB; every copy carries ~15 us of engine set-up, measured with chunk sweeps
    // both directly issued and replayed from a CUDA graph, so small uniform chunks lose).  Element counts are
    // even so that every chunk of every array stays 16-byte aligned for the TMA-fed kernels; an odd leftover
    // element goes to the last chunk.
    auto schedule = [&](size_t first_bytes, size_t max_bytes) {
        std::vector<int> sched, back;
        auto even      = [](size_t n) { int v = (int)n & ~1; return v < 2 ? 2 : v; };
        const int smax = even(max_bytes / el_bytes);
        int sz         = even(first_bytes / el_bytes);
        if (sz > smax) sz = smax;
        int remaining = op->nElmt & ~1;
        while (remaining > 0)
        {
            int t = sz < remaining ? sz : remaining;
            sched.push_back(t);
            remaining -= t;
            if (remaining > 0)
            {
                t = sz < remaining ? sz : remaining;
                back.push_back(t);
                remaining -= t;
            }
            sz = 2 * sz < smax ? 2 * sz : smax;
        }
        sched.insert(sched.end(), back.rbegin(), back.rend());
        if (op->nElmt & 1)
        {
            if (sched.empty()) sched.push_back(1);
            else sched.back() += 1;
        }
        return sched;
    };
    // queues the whole pipeline on (sH, sK, sD); used directly and under stream capture
    auto issue = [&](const std::vector<int> &sched) -> int {
        const int nChunks = (int)sched.size();
        while ((int)op->pipe_ev.size() < 2 * nChunks)
        {
            cudaEvent_t ev;
            NEKMF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            op->pipe_ev.push_back(ev);
        }
        int e0 = 0;
        for (int c = 0; c < nChunks; e0 += sched[c], ++c)
        {
            const int ne = sched[c];
            const double *din[3];
            double *dout[3];
            for (int a = 0; a < 3; ++a)
            {
                din[a]  = op->d_stage_in + (a < nin ? a : 0) * in_sz + (size_t)e0 * in_el;
                dout[a] = op->d_stage_out + (a < nout ? a : 0) * out_sz + (size_t)e0 * out_el;
            }
            for (int a = 0; a < nin; ++a)
                NEKMF_CUDA(cudaMemcpyAsync(const_cast<double *>(din[a]), ins[a] + (size_t)e0 * in_el,
                                           (size_t)ne * in_el * 8, cudaMemcpyHostToDevice, sH));
            NEKMF_CUDA(cudaEventRecord(op->pipe_ev[2 * c], sH));
            NEKMF_CUDA(cudaStreamWaitEvent(sK, op->pipe_ev[2 * c], 0));
            op->run_e0     = e0;
            op->run_ne     = ne;
            op->run_stream = sK;
            const int lrc  = op->launch(op, din, dout);
            if (lrc) return lrc;
            NEKMF_CUDA(cudaEventRecord(op->pipe_ev[2 * c + 1], sK));
            NEKMF_CUDA(cudaStreamWaitEvent(sD, op->pipe_ev[2 * c + 1], 0));
            for (int a = 0; a < nout; ++a)
                NEKMF_CUDA(cudaMemcpyAsync(outs[a] + (size_t)e0 * out_el, dout[a], (size_t)ne * out_el * 8,
                                           cudaMemcpyDeviceToHost, sD));
        }
        return NEKMF_OK;
    };

    // work queued earlier on the operator's stream (e.g. a device-array apply) stays ordered before us
    NEKMF_CUDA(cudaEventRecord(op->pipe_done, op->stream));
    for (int s = 0; s < 3; ++s) NEKMF_CUDA(cudaStreamWaitEvent(op->pipe_stream[s], op->pipe_done, 0));
    if (op->timing) NEKMF_CUDA(cudaEventRecord(op->ev0, sH));

    rc = issue(schedule((size_t)2 << 20, chunk_bytes));
    op->run_e0     = 0;
    op->run_ne     = op->nElmt;
    op->run_stream = op->stream;
    if (op->timing && rc == NEKMF_OK)
    {
        // ev0/ev1 bracket the whole pipelined call (copies included) for host-array applies
        NEKMF_CUDA(cudaEventRecord(op->pipe_done, sD));
        NEKMF_CUDA(cudaStreamWaitEvent(sH, op->pipe_done, 0));
        NEKMF_CUDA(cudaEventRecord(op->ev1, sH));
        op->timed_once = true;
    }
    for (int s = 0; s < 3; ++s)
    {
        cudaError_t e = cudaStreamSynchronize(op->pipe_stream[s]);
        if (e != cudaSuccess && rc == NEKMF_OK)
        {
            set_error("nekmf_op_apply: host pipeline failed: %s", cudaGetErrorString(e));
            rc = NEKMF_ERR_CUDA;
        }
    }
    return rc;
}

int nekmf_op_destroy(nekmf_op_t op)
{
    if (!op) return NEKMF_OK;
    if (op->kstate && op->kstate_free) op->kstate_free(op->kstate);
    cudaFree(op->d_tab);
    cudaFree(op->d_jac);
    cudaFree(op->d_df);
    cudaFree(op->d_stage_in);
    cudaFree(op->d_stage_out);
    for (int s = 0; s < 3; ++s)
        if (op->pipe_stream[s]) cudaStreamDestroy(op->pipe_stream[s]);
    if (op->pipe_done) cudaEventDestroy(op->pipe_done);
    for (cudaEvent_t ev : op->pipe_ev) cudaEventDestroy(ev);
    if (op->ev0) cudaEventDestroy(op->ev0);
    if (op->ev1) cudaEventDestroy(op->ev1);
    delete op;
    return NEKMF_OK;
}

} // extern "C"
