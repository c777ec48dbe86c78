// shape_inst.cu -- instantiates the Quad / Tri / Prism / Pyr / Tet kernels of shape_kernels.cuh for one polynomial
// order (NM = SHAPE_NM modes per direction, default quadrature) and provides their launchers.  Compiled once
// per order so the orders build in parallel.
#include "shape_kernels.cuh"
#include "op_internal.h"
#include <string.h>
#include <vector>

#ifndef SHAPE_NM
#error "compile with -DSHAPE_NM=<modes per direction>"
#endif
#define SHP_CAT2(a, b) a##b
#define SHP_CAT(a, b) SHP_CAT2(a, b)

namespace nekmf
{

template <int SHAPE, int NM> struct ShpState
{
    ShpTab<SHAPE, NM> tab;
    double *d_aux = nullptr;
    int blocks_per_sm = 0;
};

template <int SHAPE, int OP, int NM, bool DEF> static int shape_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    using Dm  = ShpDims<SHAPE, NM>;
    auto *st  = static_cast<ShpState<SHAPE, NM> *>(op->kstate);
    auto kern = shape_op_kernel<SHAPE, OP, NM, DEF>;
    // coefficient-input operators carry a second coefficient buffer behind the common layout, PhysDeriv a second
    // quadrature buffer (next-batch prefetch)
    constexpr size_t SMEM = Dm::SMEM + ((OP == NEKMF_BWDTRANS || OP == NEKMF_HELMHOLTZ) ? (size_t)Dm::CINSZ * 8 : 0) +
                            (OP == NEKMF_PHYSDERIV ? (size_t)Dm::BUF * 8 : 0); // second input buffer (next-batch prefetch)
    if (st->blocks_per_sm == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Dm::T, SMEM));
        if (nb < 1)
        {
            set_error("shape kernel <%d,%d,%d,%d> does not fit on an SM (smem %zu)", SHAPE, OP, NM, (int)DEF, (size_t)SMEM);
            return NEKMF_ERR_CUDA;
        }
        st->blocks_per_sm = nb;
    }
    ShpArgs a;
    const size_t gstep = DEF ? (size_t)op->nqTot : 1;
    a.in0 = in[0]; a.in1 = in[1]; a.in2 = in[2];
    a.out0 = out[0]; a.out1 = out[1]; a.out2 = out[2];
    a.jac = op->d_jac ? op->d_jac + (size_t)op->run_e0 * gstep : nullptr;
    a.df  = op->d_df ? op->d_df + (size_t)op->run_e0 * gstep : nullptr;
    a.aux = st->d_aux;
    a.dfStride = (size_t)op->nElmt * gstep;
    a.nElmt    = op->run_ne;
    a.lambda   = op->lambda;
    const int nBatches = (op->run_ne + Dm::E - 1) / Dm::E;
    int grid           = st->blocks_per_sm * NUM_SMS;
    if (grid > nBatches) grid = nBatches;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Dm::T, SMEM, op->run_stream>>>(st->tab, a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <int SHAPE, int NM> static bool shape_install(nekmf_op_s *op)
{
    using Dm = ShpDims<SHAPE, NM>;
    for (int d = 0; d < Dm::DIM; ++d)
        if (op->nm[d] != NM) return false;
    if (op->nq[0] != Dm::NQ0 || op->nq[1] != Dm::NQ1 || (Dm::DIM == 3 && op->nq[2] != Dm::NQ2)) return false;
    if (Dm::B1C_ROWS && op->rows[1] != Dm::B1C_ROWS) return false;
    if (Dm::B2C_ROWS && op->rows[2] != Dm::B2C_ROWS) return false;

    auto *st = new ShpState<SHAPE, NM>;
    memset(&st->tab, 0, sizeof(st->tab));
    memcpy(st->tab.b0, op->b[0].data(), sizeof(double) * NM * Dm::NQ0);
    if (!Dm::B1C_ROWS) memcpy(st->tab.b1t, op->b[1].data(), sizeof(double) * NM * Dm::NQ1);
    memcpy(st->tab.D0, op->D[0].data(), sizeof(double) * Dm::NQ0 * Dm::NQ0);
    memcpy(st->tab.D1, op->D[1].data(), sizeof(double) * Dm::NQ1 * Dm::NQ1);
    if (Dm::DIM == 3) memcpy(st->tab.D2, op->D[2].data(), sizeof(double) * Dm::NQ2 * Dm::NQ2);

    std::vector<double> aux(Dm::AUX_LEN, 0.0);
    if (Dm::B1C_ROWS) memcpy(aux.data() + Dm::OFF_B1C, op->b[1].data(), sizeof(double) * Dm::B1C_ROWS * Dm::NQ1);
    if (Dm::B2C_ROWS) memcpy(aux.data() + Dm::OFF_B2C, op->b[2].data(), sizeof(double) * Dm::B2C_ROWS * Dm::NQ2);
    for (int d = 0; d < Dm::DIM; ++d)
        for (int i = 0; i < op->nq[d]; ++i) aux[Dm::OFF_W + d * Dm::NQM + i] = op->ws[d][i];
    // collapsed-coordinate factors: Helmholtz.h:304-315 (Tri), 1017-1028 (Prism), 1490-1508 (Pyr), 1985-2004 (Tet)
    double *h0 = aux.data() + Dm::OFF_H, *h1 = h0 + Dm::NQM, *h2 = h1 + Dm::NQM, *h3 = h2 + Dm::NQM;
    for (int i = 0; i < Dm::NQ0; ++i) h0[i] = 0.5 * (1.0 + op->Z[0][i]);
    if (SHAPE == NEKMF_TRI)
        for (int j = 0; j < Dm::NQ1; ++j) h1[j] = 2.0 / (1.0 - op->Z[1][j]);
    if (SHAPE == NEKMF_PRISM || SHAPE == NEKMF_PYR)
        for (int k = 0; k < Dm::NQ2; ++k) h1[k] = 2.0 / (1.0 - op->Z[2][k]);
    if (SHAPE == NEKMF_PYR)
        for (int j = 0; j < Dm::NQ1; ++j) h2[j] = 0.5 * (1.0 + op->Z[1][j]);
    if (SHAPE == NEKMF_TET)
    {
        for (int j = 0; j < Dm::NQ1; ++j)
        {
            h1[j] = 0.5 * (1.0 + op->Z[1][j]);
            h2[j] = 2.0 / (1.0 - op->Z[1][j]);
        }
        for (int k = 0; k < Dm::NQ2; ++k) h3[k] = 2.0 / (1.0 - op->Z[2][k]);
    }
    if (cudaMalloc(&st->d_aux, aux.size() * 8) != cudaSuccess ||
        cudaMemcpy(st->d_aux, aux.data(), aux.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess)
    {
        cudaFree(st->d_aux);
        delete st;
        return false;
    }
    op->kstate      = st;
    op->geo_pitch   = op->nqTot;
    op->kstate_free = [](void *p) {
        auto *s = static_cast<ShpState<SHAPE, NM> *>(p);
        cudaFree(s->d_aux);
        delete s;
    };
    const char *sn[6] = {"Quad", "Tri", "Hex", "Prism", "Pyr", "Tet"};
    const char *opn[5] = {"bwd", "helm", "iprod", "ipwdb", "physderiv"};
    char name[96];
    snprintf(name, sizeof(name), "shape_op_kernel<%s,%s,nm=%d,%s>", sn[SHAPE], opn[op->optype], NM, op->deformed ? "deformed" : "regular");
    op->kname = name;
#define SHP_CASE(OPC)                                                                                           \
    case OPC:                                                                                                   \
        op->launch = op->deformed ? shape_launch<SHAPE, OPC, NM, true> : shape_launch<SHAPE, OPC, NM, false>;   \
        return true;
    switch (op->optype)
    {
        SHP_CASE(NEKMF_BWDTRANS)
        SHP_CASE(NEKMF_HELMHOLTZ)
        SHP_CASE(NEKMF_IPRODUCTWRTBASE)
        SHP_CASE(NEKMF_PHYSDERIV)
        SHP_CASE(NEKMF_IPRODUCTWRTDERIVBASE)
    }
#undef SHP_CASE
    op->kstate_free(st);
    op->kstate      = nullptr;
    op->kstate_free = nullptr;
    return false;
}

bool SHP_CAT(shape_try_nm, SHAPE_NM)(nekmf_op_s *op)
{
    if (op->nm[0] != SHAPE_NM) return false;
    switch (op->shape)
    {
        case NEKMF_QUAD: return shape_install<NEKMF_QUAD, SHAPE_NM>(op);
        case NEKMF_TRI: return shape_install<NEKMF_TRI, SHAPE_NM>(op);
        case NEKMF_PRISM: return shape_install<NEKMF_PRISM, SHAPE_NM>(op);
        case NEKMF_TET: return shape_install<NEKMF_TET, SHAPE_NM>(op);
        case NEKMF_PYR: return shape_install<NEKMF_PYR, SHAPE_NM>(op);
        default: return false;
    }
}

} // namespace nekmf
