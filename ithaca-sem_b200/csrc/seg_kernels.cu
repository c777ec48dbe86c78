// seg_kernels.cu -- the 1-D (segment) operators: BwdTrans, IProductWRTBase, PhysDeriv, IProductWRTDerivBase for a
// segment embedded in coordim = 1, 2 or 3 space dimensions (boundary / trace expansions of 2-D and 3-D meshes).
//
// Reference semantics: MatrixFreeOps/BwdTransKernels.hpp:14-33, IProductKernels.hpp:39-75,
// PhysDerivKernels.hpp:13-38 + PhysDeriv.h:60-250, IProductWRTDerivBase.h:182-365.
//
// A segment is a handful of doubles: one thread computes one output value (a dot product of length nm or nq with
// a table row held in shared memory); consecutive threads write consecutive outputs, so loads of the element's
// input line are broadcast / coalesced and stores are coalesced.  Purely HBM-bound streaming.
#include "op_internal.h"

namespace nekmf
{

struct SegArgs
{
    const double *in0, *in1, *in2;
    double *out0, *out1, *out2;
    const double *jac, *df; // jac [nElmt|nElmt*nq]; df [coordim][nElmt|nElmt*nq]
    const double *tab;      // packed tables of direction 0 (nekmf_op_s::tab_off)
    int tab_len, off[5];
    int nm, nq, nElmt, deformed, optype, coordim;
    size_t dfStride;
};

__global__ void __launch_bounds__(256) seg_kernel(const __grid_constant__ SegArgs a)
{
    extern __shared__ __align__(16) double sTab[];
    for (int i = threadIdx.x; i < a.tab_len; i += blockDim.x) sTab[i] = __ldg(a.tab + i);
    __syncthreads();
    const double *b = sTab + a.off[0], *db = sTab + a.off[1], *D = sTab + a.off[2], *w = sTab + a.off[4];
    const int nm = a.nm, nq = a.nq;
    const bool def = a.deformed != 0;
    const bool cout = a.optype == NEKMF_IPRODUCTWRTBASE || a.optype == NEKMF_IPRODUCTWRTDERIVBASE;
    const int nout  = cout ? nm : nq;
    const size_t total = (size_t)a.nElmt * nout;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
    {
        const int e = (int)(t / nout), o = (int)(t - (size_t)e * nout);
        const size_t g = def ? (size_t)e * nq : (size_t)e; // geometry entry of (element, point 0)
        if (a.optype == NEKMF_BWDTRANS)
        {
            const double *x = a.in0 + (size_t)e * nm;
            double s = x[0] * b[o];
            for (int p = 1; p < nm; ++p) s = fma(x[p], b[p * nq + o], s);
            a.out0[t] = s;
        }
        else if (a.optype == NEKMF_PHYSDERIV)
        {
            const double *x = a.in0 + (size_t)e * nq;
            double d = 0.0;
            for (int k = 0; k < nq; ++k) d = fma(D[k * nq + o], x[k], d);
            const size_t gi = def ? g + o : g;
            a.out0[t] = d * __ldg(a.df + gi);
            if (a.coordim >= 2) a.out1[t] = d * __ldg(a.df + a.dfStride + gi);
            if (a.coordim == 3) a.out2[t] = d * __ldg(a.df + 2 * a.dfStride + gi);
        }
        else
        {
            // inner product with bdata (IProductWRTBase) or dbdata of sum_c df[c] in_c (IProductWRTDerivBase)
            const bool deriv = a.optype == NEKMF_IPRODUCTWRTDERIVBASE;
            const double *B  = deriv ? db : b;
            const size_t q0  = (size_t)e * nq;
            double s = 0.0;
            for (int i = 0; i < nq; ++i)
            {
                const size_t gi = def ? g + i : g;
                double v = a.in0[q0 + i];
                if (deriv)
                {
                    v *= __ldg(a.df + gi);
                    if (a.coordim >= 2) v = fma(__ldg(a.df + a.dfStride + gi), a.in1[q0 + i], v);
                    if (a.coordim == 3) v = fma(__ldg(a.df + 2 * a.dfStride + gi), a.in2[q0 + i], v);
                }
                s = fma(v * B[o * nq + i] * __ldg(a.jac + gi), w[i], s);
            }
            a.out0[t] = s;
        }
    }
}

static int seg_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    SegArgs a;
    a.in0 = in[0]; a.in1 = in[1]; a.in2 = in[2];
    a.out0 = out[0]; a.out1 = out[1]; a.out2 = out[2];
    const size_t gstep = op->deformed ? (size_t)op->nqTot : 1;
    a.jac = op->d_jac ? op->d_jac + (size_t)op->run_e0 * gstep : nullptr;
    a.df  = op->d_df ? op->d_df + (size_t)op->run_e0 * gstep : nullptr;
    a.dfStride = (size_t)op->nElmt * gstep;
    a.tab = op->d_tab; a.tab_len = op->tab_len;
    for (int t = 0; t < 5; ++t) a.off[t] = op->tab_off[0][t];
    a.nm = op->nm[0]; a.nq = op->nq[0]; a.nElmt = op->run_ne; a.deformed = op->deformed; a.optype = op->optype;
    a.coordim = op->coordim;
    const bool cout    = op->optype == NEKMF_IPRODUCTWRTBASE || op->optype == NEKMF_IPRODUCTWRTDERIVBASE;
    const size_t total = (size_t)op->run_ne * (cout ? op->nm[0] : op->nq[0]);
    if (total == 0) return NEKMF_OK;
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)NUM_SMS * 8) blocks = (size_t)NUM_SMS * 8;
    seg_kernel<<<(int)blocks, 256, (size_t)op->tab_len * 8, op->run_stream>>>(a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

bool select_seg(nekmf_op_s *op)
{
    if (op->shape != NEKMF_SEG) return false;
    char name[96];
    snprintf(name, sizeof(name), "seg_kernel(op=%d,nm=%d,nq=%d,coordim=%d)", op->optype, op->nm[0], op->nq[0], op->coordim);
    op->kname  = name;
    op->launch = seg_launch;
    return true;
}

} // namespace nekmf
