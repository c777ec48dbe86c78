// hex_inst.cu -- instantiates the hexahedral kernels for one polynomial order (NM = HEX_NM modes per
// direction, NQ = NM+1 and NM+2 quadrature points) and provides their launchers.  Compiled once per order so the
// orders build in parallel.
#include "hex_kernels.cuh"
#include "hex_slab.cuh"
#include "op_internal.h"
#include <stdlib.h>
#include <string.h>

#ifndef HEX_NM
#error "compile with -DHEX_NM=<modes per direction>"
#endif
#define HEX_CAT2(a, b) a##b
#define HEX_CAT(a, b) HEX_CAT2(a, b)

namespace nekmf
{

template <int OP, int NM, int NQ, bool DEF> static int hex_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    using Cfg = HexCfg<OP, NM, NQ, DEF>;
    static int blocks_per_sm = 0;
    auto kern                = hex_op_kernel<OP, NM, NQ, DEF>;
    if (blocks_per_sm == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1)
        {
            set_error("hex kernel <%d,%d,%d,%d> does not fit on an SM (smem %zu)", OP, NM, NQ, (int)DEF, (size_t)Cfg::SMEM);
            return NEKMF_ERR_CUDA;
        }
        blocks_per_sm = nb;
    }
    HexArgs a;
    a.in0 = in[0]; a.in1 = in[1]; a.in2 = in[2];
    a.out0 = out[0]; a.out1 = out[1]; a.out2 = out[2];
    const size_t gstep = DEF ? (size_t)op->geo_pitch : 1; // geometry entries per element
    a.jac = op->d_jac ? op->d_jac + (size_t)op->run_e0 * gstep : nullptr;
    a.df  = op->d_df ? op->d_df + (size_t)op->run_e0 * gstep : nullptr;
    a.nElmt    = op->run_ne;
    a.dfStride = (size_t)op->nElmt * gstep;
    a.lambda   = op->lambda;
    a.in_aligned = (((uintptr_t)in[0] | (uintptr_t)in[1] | (uintptr_t)in[2]) & 15) == 0;
    const int nBatches = (op->run_ne + Cfg::E - 1) / Cfg::E;
    int grid           = blocks_per_sm * NUM_SMS;
    if (grid > nBatches) grid = nBatches;
    if (grid < 1) return NEKMF_OK;
    const HexTab<NM, NQ> *tab = static_cast<const HexTab<NM, NQ> *>(op->kstate);
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(*tab, a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

// register-slab kernels (hex_slab.cuh) for BwdTrans / IProductWRTBase (regular geometry) with the default
// quadrature.  Measured against the pencil kernels (fraction of HBM peak, nm = 2..6): BwdTrans .58->.86, .61->.70,
// .64->.93, .73->.93, .73->.83; IProductWRTBase .50->.88, .66->.81, .69->.85, .77->.81, .64->.70.  From nm = 7 the
// slab (nm^2 doubles plus a line) no longer fits the register file without spilling and the pencil kernels win.
#ifndef HEX_SLAB_MAX_NM
#define HEX_SLAB_MAX_NM 6
#endif
template <int OP, int NM> static int hex_slab_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    using Cfg = SlabCfg<OP, NM>;
    static int blocks_per_sm = 0;
    auto kern                = hex_slab_kernel<OP, NM>;
    if (blocks_per_sm == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1)
        {
            set_error("hex slab kernel <%d,%d> does not fit on an SM (smem %zu)", OP, NM, (size_t)Cfg::SMEM);
            return NEKMF_ERR_CUDA;
        }
        blocks_per_sm = nb;
    }
    SlabArgs a;
    a.in  = in[0];
    a.out = out[0];
    a.jac   = op->d_jac ? op->d_jac + (size_t)op->run_e0 : nullptr;
    a.nElmt = op->run_ne;
    a.io_aligned = (((uintptr_t)in[0] | (uintptr_t)out[0]) & 15) == 0;
    const int nBatches = (op->run_ne + Cfg::EPW * Cfg::WARPS - 1) / (Cfg::EPW * Cfg::WARPS);
    int grid           = blocks_per_sm * NUM_SMS;
    if (grid > nBatches) grid = nBatches;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(*static_cast<const HexTab<NM, NM + 1> *>(op->kstate), a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}
template <int NM> static int hex_pd_slab_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    using Cfg = PdSlabCfg<NM>;
    static int blocks_per_sm = 0;
    auto kern                = hex_pd_slab_kernel<NM>;
    if (blocks_per_sm == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("hex PhysDeriv slab kernel <%d> does not fit on an SM", NM); return NEKMF_ERR_CUDA; }
        blocks_per_sm = nb;
    }
    PdSlabArgs a;
    a.in = in[0]; a.out0 = out[0]; a.out1 = out[1]; a.out2 = out[2];
    a.df = op->d_df + (size_t)op->run_e0;
    a.dfStride = (size_t)op->nElmt;
    a.nElmt    = op->run_ne;
    a.io_aligned = (((uintptr_t)in[0] | (uintptr_t)out[0] | (uintptr_t)out[1] | (uintptr_t)out[2]) & 15) == 0;
    const int nBatches = (op->run_ne + Cfg::EPW * Cfg::WARPS - 1) / (Cfg::EPW * Cfg::WARPS);
    int grid           = blocks_per_sm * NUM_SMS;
    if (grid > nBatches) grid = nBatches;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(*static_cast<const HexTab<NM, NM + 1> *>(op->kstate), a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}
// the operator's constant tables: HexTab (B, D, w) followed by dbdata for the kernels that want it
template <int NM, int NQ> struct HexTabX : HexTab<NM, NQ>
{
    double dB[NM * NQ];
};

template <int NM> static int hex_ipwdb_slab_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    using Cfg = IpwdbSlabCfg<NM>;
    static int blocks_per_sm = 0;
    auto kern                = hex_ipwdb_slab_kernel<NM>;
    if (blocks_per_sm == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("hex IProductWRTDerivBase slab kernel <%d> does not fit on an SM", NM); return NEKMF_ERR_CUDA; }
        blocks_per_sm = nb;
    }
    const auto *tx = static_cast<const HexTabX<NM, NM + 1> *>(op->kstate);
    SlabDTab<NM> dt;
    memcpy(dt.dB, tx->dB, sizeof(dt.dB));
    IpwdbSlabArgs a;
    a.in0 = in[0]; a.in1 = in[1]; a.in2 = in[2]; a.out = out[0];
    a.jac = op->d_jac + (size_t)op->run_e0;
    a.df  = op->d_df + (size_t)op->run_e0;
    a.dfStride = (size_t)op->nElmt;
    a.nElmt    = op->run_ne;
    a.io_aligned = (((uintptr_t)in[0] | (uintptr_t)in[1] | (uintptr_t)in[2] | (uintptr_t)out[0]) & 15) == 0;
    const int nBatches = (op->run_ne + Cfg::EPW * Cfg::WARPS - 1) / (Cfg::EPW * Cfg::WARPS);
    int grid           = blocks_per_sm * NUM_SMS;
    if (grid > nBatches) grid = nBatches;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(*static_cast<const HexTab<NM, NM + 1> *>(tx), dt, a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <int NM, int NQ> static bool hex_slab_install(nekmf_op_s *op)
{
    if constexpr (NQ == NM + 1 && NM <= HEX_SLAB_MAX_NM)
    {
        const char *v = getenv("NEKMF_HEX_SLAB"); // NEKMF_HEX_SLAB=0: keep the pencil kernels (A/B comparisons)
        if (v && v[0] == '0') return false;
        // PhysDeriv slab kernel.  Measured against the pencil kernel (fraction of HBM peak, nm = 2..6): 0.70 -> 0.82,
        // 0.72 -> 0.81, 0.77 -> 0.91, 0.77 -> 0.73, 0.55 -> 0.72: the pencil kernel is kept at nm = 5.
        if (op->optype == HEX_PD && !op->deformed && NM != 5)
        {
            const char *vp = getenv("NEKMF_HEX_PD_SLAB"); // NEKMF_HEX_PD_SLAB=0: pencil kernel
            if (vp && vp[0] == '0') return false;
            char pname[96];
            snprintf(pname, sizeof(pname), "hex_pd_slab_kernel<nm=%d,nq=%d,regular>", NM, NQ);
            op->kname  = pname;
            op->launch = hex_pd_slab_launch<NM>;
            return true;
        }
        // IProductWRTDerivBase slab kernel: 0.47 -> 0.72, 0.55 -> 0.73, 0.55 -> 0.66 of the HBM peak at nm = 2, 3, 4;
        // at nm = 5, 6 the three passes need more than 255 registers (spills: 0.51 -> 0.43, 0.34 -> 0.21), pencil kept
        if (op->optype == HEX_IPWDB && !op->deformed && NM <= 4)
        {
            const char *vi = getenv("NEKMF_HEX_IPWDB_SLAB"); // NEKMF_HEX_IPWDB_SLAB=0: pencil kernel
            if (vi && vi[0] == '0') return false;
            char iname[96];
            snprintf(iname, sizeof(iname), "hex_ipwdb_slab_kernel<nm=%d,nq=%d,regular>", NM, NQ);
            op->kname  = iname;
            op->launch = hex_ipwdb_slab_launch<NM>;
            return true;
        }
        if (op->optype != HEX_BWD && !(op->optype == HEX_IPROD && !op->deformed)) return false;
        char name[96];
        snprintf(name, sizeof(name), "hex_slab_kernel<%s,nm=%d,nq=%d,%s>", op->optype == HEX_BWD ? "bwd" : "iprod", NM, NQ,
                 op->deformed ? "deformed" : "regular");
        op->kname = name;
        op->launch = op->optype == HEX_BWD ? hex_slab_launch<HEX_BWD, NM> : hex_slab_launch<HEX_IPROD, NM>;
        return true;
    }
    return false;
}

template <int NM, int NQ> static bool hex_install(nekmf_op_s *op)
{
    auto *tab = new HexTabX<NM, NQ>;
    memcpy(tab->dB, op->db[0].data(), sizeof(tab->dB));
    memcpy(tab->B, op->b[0].data(), sizeof(tab->B));
    memcpy(tab->D, op->D[0].data(), sizeof(tab->D));
    memcpy(tab->w, op->ws[0].data(), sizeof(tab->w));
    op->kstate      = tab;
    op->geo_pitch   = round_up(NQ * NQ * NQ, 2);
    op->kstate_free = [](void *p) { delete static_cast<HexTabX<NM, NQ> *>(p); };
    char name[96];
    const char *opn[5] = {"bwd", "helm", "iprod", "ipwdb", "physderiv"};
    snprintf(name, sizeof(name), "hex_op_kernel<%s,nm=%d,nq=%d,%s>", opn[op->optype], NM, NQ, op->deformed ? "deformed" : "regular");
    op->kname = name;
    if (hex_slab_install<NM, NQ>(op)) return true;
#define HEX_CASE(OPC)                                                                     \
    case OPC:                                                                             \
        op->launch = op->deformed ? hex_launch<OPC, NM, NQ, true> : hex_launch<OPC, NM, NQ, false>; \
        return true;
    switch (op->optype)
    {
        HEX_CASE(HEX_BWD)
        HEX_CASE(HEX_HELM)
        HEX_CASE(HEX_IPROD)
        HEX_CASE(HEX_IPWDB)
        HEX_CASE(HEX_PD)
    }
#undef HEX_CASE
    delete tab;
    op->kstate    = nullptr;
    op->geo_pitch = op->nqTot;
    return false;
}

bool HEX_CAT(hex_try_nm, HEX_NM)(nekmf_op_s *op)
{
    if (op->nm[0] != HEX_NM) return false;
    if (op->nq[0] == HEX_NM + 1) return hex_install<HEX_NM, HEX_NM + 1>(op);
#ifdef HEX_WITH_NQ2
    if (op->nq[0] == HEX_NM + 2) return hex_install<HEX_NM, HEX_NM + 2>(op);
#endif
    return false;
}

} // namespace nekmf
