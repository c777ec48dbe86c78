// quad_lane.cu -- BwdTrans, IProductWRTBase, PhysDeriv and (regular) IProductWRTDerivBase on quadrilaterals with ONE
// LANE PER ELEMENT.
//
// Reference semantics: MatrixFreeOps/BwdTransKernels.hpp:35-76, IProductKernels.hpp:76-133,
// PhysDerivKernels.hpp:39-90 + PhysDeriv.h (Quad): out_0 = df0 d0 + df1 d1, out_1 = df2 d0 + df3 d1.
//
// A 2-D element (nm^2 coefficients, nq^2 = (nm+1)^2 quadrature values) fits the registers of a single lane: both
// contractions happen there, no exchange between lanes, no barrier (the scheme of quad_kron.cu).  Every warp is an
// independent worker on batches of 32 elements.  Of nm^2 and nq^2 exactly one is odd: that side travels as ONE
// bulk TMA copy per batch (odd lane stride: conflict-free shared-memory accesses); the even side sits in slots
// padded by two doubles, filled by warp-wide 16-byte cp.async copies / drained by warp-wide 16-byte stores.
// Deformed IProductWRTBase stages the per-point Jacobian like an input; deformed PhysDeriv (four more arrays) keeps
// the CTA-level kernel of shape_kernels.cuh.
#include "hex_kernels.cuh"
#include "op_internal.h"
#include <stdlib.h>
#include <string.h>

namespace nekmf
{

enum { QL_BWD = 0, QL_IPROD = 1, QL_PD = 2, QL_IPWDB = 3 };

template <int NM> struct QLaneTab
{
    double B[NM * (NM + 1)];       // bdata[m*NQ+i]
    double D[(NM + 1) * (NM + 1)]; // D[k*NQ+i] = dh_k/dz(z_i)
    double w[NM + 1];
    double dB[NM * (NM + 1)];      // dbdata[m*NQ+i] (IProductWRTDerivBase)
};

struct QLaneArgs
{
    const double *in;
    const double *in1; // second input (IProductWRTDerivBase)
    double *out0, *out1;
    const double *jac; // [nElmt] | [nElmt][nq^2]
    const double *df;  // [4][dfStride] (regular)
    size_t dfStride;
    int nElmt;
    int io_aligned;
};

template <int OP, int NM, bool DEF> struct QLaneCfg
{
    static constexpr int NQ = NM + 1, NM2 = NM * NM, NQ2 = NQ * NQ;
    static constexpr int INL  = OP == QL_BWD ? NM2 : NQ2; // doubles per element, input side
    static constexpr int OUTL = (OP == QL_IPROD || OP == QL_IPWDB) ? NM2 : NQ2;
    static constexpr bool INPAD = (INL % 2) == 0, OUTPAD = (OUTL % 2) == 0;
    // lane stride of a slot: odd lengths are conflict free as they are; even lengths must stay even (16-byte copies)
    // and are best at 2 (mod 4) doubles -- 2-way bank conflicts; a multiple of 4 would be 4- to 16-way
    static constexpr int INS  = (INPAD && INL % 4 == 0) ? INL + 2 : INL, OUTS = (OUTPAD && OUTL % 4 == 0) ? OUTL + 2 : OUTL;
    static constexpr int NIN  = 1 + (((OP == QL_IPROD && DEF) || OP == QL_IPWDB) ? 1 : 0);
    static constexpr int NOUT = OP == QL_PD ? 2 : 1;
    static constexpr int INB  = round_up(32 * INS, 2), OUTB = round_up(32 * OUTS, 2);
    static constexpr int PER_WARP = NIN * INB + NOUT * OUTB + 2;
    static constexpr int W_FIT  = (200 * 1024) / (PER_WARP * 8);
    static constexpr int WARPS  = W_FIT >= 16 ? 16 : (W_FIT >= 12 ? 12 : (W_FIT >= 8 ? 8 : (W_FIT >= 4 ? 4 : (W_FIT >= 1 ? W_FIT : 1))));
    static constexpr int T      = WARPS * 32;
    static constexpr size_t SMEM = (size_t)WARPS * PER_WARP * 8 + 16;
};

__device__ __forceinline__ void ql_cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ql_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int OP, int NM, bool DEF>
__global__ void __launch_bounds__(QLaneCfg<OP, NM, DEF>::T, 1)
    quad_lane_kernel(const __grid_constant__ QLaneTab<NM> tab, const __grid_constant__ QLaneArgs args)
{
    using Cfg = QLaneCfg<OP, NM, DEF>;
    constexpr int NQ = Cfg::NQ, INL = Cfg::INL, OUTL = Cfg::OUTL, INS = Cfg::INS, OUTS = Cfg::OUTS;
    constexpr bool INPAD = Cfg::INPAD, OUTPAD = Cfg::OUTPAD;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *wbase  = reinterpret_cast<double *>(smem_raw) + (size_t)warp * Cfg::PER_WARP;
    double *sIn    = wbase;
    double *sJac   = sIn + Cfg::INB; // (deformed IProduct only)
    double *sO0    = sIn + Cfg::NIN * Cfg::INB;
    double *sO1    = sO0 + Cfg::OUTB; // (PhysDeriv only)
    uint64_t *bar  = reinterpret_cast<uint64_t *>(sO0 + Cfg::NOUT * Cfg::OUTB);

    const int nElmt = args.nElmt;
    const int nB    = (nElmt + 31) / 32;
    const int GW    = gridDim.x * Cfg::WARPS;
    const int gw    = blockIdx.x * Cfg::WARPS + warp;
    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto batch_ne = [&](int b) { int r = nElmt - b * 32; return r < 32 ? r : 32; };
    auto in_ok    = [&](int b) { return args.io_aligned && (INPAD || ((batch_ne(b) * INL) & 1) == 0); };
    auto out_ok   = [&](int b) { return args.io_aligned && (OUTPAD || ((batch_ne(b) * OUTL) & 1) == 0); };
    auto pad_in   = [&](int i2) { const int e = (2 * i2) / INL; return e * INS + (2 * i2 - e * INL); };
    auto pad_out  = [&](int i2) { const int e = (2 * i2) / OUTL; return e * OUTS + (2 * i2 - e * OUTL); };

    uint32_t phase = 0;
    for (int b = gw; b < nB; b += GW)
    {
        const int ne = batch_ne(b);
        const bool fin = in_ok(b), fout = out_ok(b);
        const size_t ioff = (size_t)b * 32 * INL, ooff = (size_t)b * 32 * OUTL;
        // the staging slots must no longer be read by the previous batch's bulk stores
        tma_store_wait_read0();
        __syncwarp();
        if (fin && INPAD)
        {
            for (int i2 = lane; i2 < ne * INL / 2; i2 += 32)
            {
                const int a = pad_in(i2);
                ql_cp_async16(sIn + a, args.in + ioff + 2 * i2);
                if (Cfg::NIN == 2) ql_cp_async16(sJac + a, (OP == QL_IPWDB ? args.in1 : args.jac) + ioff + 2 * i2);
            }
            ql_cp_async_wait_all();
        }
        else if (fin)
        {
            if (lane == 0)
            {
                const uint32_t bytes = (uint32_t)(ne * INL * 8);
                mbar_expect_tx(bar, bytes * Cfg::NIN);
                tma_load_1d(sIn, args.in + ioff, bytes, bar);
                if (Cfg::NIN == 2) tma_load_1d(sJac, (OP == QL_IPWDB ? args.in1 : args.jac) + ioff, bytes, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
        }
        else
        {
            for (int i = lane; i < ne * INL; i += 32)
            {
                const int a = (i / INL) * INS + (i % INL);
                sIn[a]      = __ldg(args.in + ioff + i);
                if (Cfg::NIN == 2) sJac[a] = __ldg((OP == QL_IPWDB ? args.in1 : args.jac) + ioff + i);
            }
        }
        __syncwarp();

        if (lane < ne)
        {
            const double *xe = sIn + lane * INS;
            double *o0       = sO0 + lane * OUTS;
            const size_t eg  = (size_t)b * 32 + lane;
            if (OP == QL_BWD)
            {
                double x[NM][NM];
#pragma unroll
                for (int q = 0; q < NM; ++q)
#pragma unroll
                    for (int p = 0; p < NM; ++p) x[q][p] = xe[q * NM + p];
#pragma unroll
                for (int i = 0; i < NQ; ++i)
                {
                    double t[NM];
#pragma unroll
                    for (int q = 0; q < NM; ++q)
                    {
                        double s = tab.B[i] * x[q][0];
#pragma unroll
                        for (int p = 1; p < NM; ++p) s = fma(tab.B[p * NQ + i], x[q][p], s);
                        t[q] = s;
                    }
#pragma unroll
                    for (int j = 0; j < NQ; ++j)
                    {
                        double s = tab.B[j] * t[0];
#pragma unroll
                        for (int q = 1; q < NM; ++q) s = fma(tab.B[q * NQ + j], t[q], s);
                        o0[j * NQ + i] = s;
                    }
                }
            }
            else if (OP == QL_IPROD)
            {
                double g[NQ][NQ];
                const double jr = DEF ? 1.0 : __ldg(args.jac + eg);
                const double *je = sJac + lane * INS;
#pragma unroll
                for (int j = 0; j < NQ; ++j)
#pragma unroll
                    for (int i = 0; i < NQ; ++i)
                    {
                        const double wji = tab.w[j] * tab.w[i];
                        g[j][i]          = xe[j * NQ + i] * ((DEF ? je[j * NQ + i] : jr) * wji);
                    }
#pragma unroll
                for (int p = 0; p < NM; ++p)
                {
                    double t[NQ];
#pragma unroll
                    for (int j = 0; j < NQ; ++j)
                    {
                        double s = tab.B[p * NQ] * g[j][0];
#pragma unroll
                        for (int i = 1; i < NQ; ++i) s = fma(tab.B[p * NQ + i], g[j][i], s);
                        t[j] = s;
                    }
#pragma unroll
                    for (int q = 0; q < NM; ++q)
                    {
                        double s = tab.B[q * NQ] * t[0];
#pragma unroll
                        for (int j = 1; j < NQ; ++j) s = fma(tab.B[q * NQ + j], t[j], s);
                        o0[q * NM + p] = s;
                    }
                }
            }
            else if (OP == QL_IPWDB)
            {
                // IProductWRTDerivBase.h:542-660 (regular): t_d = df[d] f_0 + df[2+d] f_1, out = IP(dB,B)[t_0] + IP(B,dB)[t_1]
                const double *ye = sJac + lane * INS; // second input
                const double jr  = __ldg(args.jac + eg);
                const double f0 = __ldg(args.df + eg), f1 = __ldg(args.df + args.dfStride + eg),
                             f2 = __ldg(args.df + 2 * args.dfStride + eg), f3 = __ldg(args.df + 3 * args.dfStride + eg);
#pragma unroll
                for (int d = 0; d < 2; ++d)
                {
                    double g[NQ][NQ];
#pragma unroll
                    for (int j = 0; j < NQ; ++j)
#pragma unroll
                        for (int i = 0; i < NQ; ++i)
                        {
                            const double t = (d == 0 ? f0 : f1) * xe[j * NQ + i] + (d == 0 ? f2 : f3) * ye[j * NQ + i];
                            g[j][i]        = t * (jr * (tab.w[j] * tab.w[i]));
                        }
#pragma unroll
                    for (int p = 0; p < NM; ++p)
                    {
                        double t[NQ];
#pragma unroll
                        for (int j = 0; j < NQ; ++j)
                        {
                            double s = (d == 0 ? tab.dB[p * NQ] : tab.B[p * NQ]) * g[j][0];
#pragma unroll
                            for (int i = 1; i < NQ; ++i) s = fma(d == 0 ? tab.dB[p * NQ + i] : tab.B[p * NQ + i], g[j][i], s);
                            t[j] = s;
                        }
#pragma unroll
                        for (int q = 0; q < NM; ++q)
                        {
                            double s = (d == 1 ? tab.dB[q * NQ] : tab.B[q * NQ]) * t[0];
#pragma unroll
                            for (int j = 1; j < NQ; ++j) s = fma(d == 1 ? tab.dB[q * NQ + j] : tab.B[q * NQ + j], t[j], s);
                            if (d == 0) o0[q * NM + p] = s;
                            else o0[q * NM + p] += s;
                        }
                    }
                }
            }
            else
            {
                double u[NQ][NQ];
#pragma unroll
                for (int j = 0; j < NQ; ++j)
#pragma unroll
                    for (int i = 0; i < NQ; ++i) u[j][i] = xe[j * NQ + i];
                const double f0 = __ldg(args.df + eg), f1 = __ldg(args.df + args.dfStride + eg),
                             f2 = __ldg(args.df + 2 * args.dfStride + eg), f3 = __ldg(args.df + 3 * args.dfStride + eg);
                double *o1 = sO1 + lane * OUTS;
#pragma unroll
                for (int j = 0; j < NQ; ++j)
#pragma unroll
                    for (int i = 0; i < NQ; ++i)
                    {
                        double d0 = tab.D[i] * u[j][0], d1 = tab.D[j] * u[0][i];
#pragma unroll
                        for (int m = 1; m < NQ; ++m)
                        {
                            d0 = fma(tab.D[m * NQ + i], u[j][m], d0);
                            d1 = fma(tab.D[m * NQ + j], u[m][i], d1);
                        }
                        o0[j * NQ + i] = f0 * d0 + f1 * d1;
                        o1[j * NQ + i] = f2 * d0 + f3 * d1;
                    }
            }
        }
        if (fout && OUTPAD)
        {
            __syncwarp();
            for (int i2 = lane; i2 < ne * OUTL / 2; i2 += 32)
            {
                const int a = pad_out(i2);
                *reinterpret_cast<double2 *>(args.out0 + ooff + 2 * i2) = *reinterpret_cast<const double2 *>(sO0 + a);
                if (OP == QL_PD)
                    *reinterpret_cast<double2 *>(args.out1 + ooff + 2 * i2) = *reinterpret_cast<const double2 *>(sO1 + a);
            }
        }
        else if (fout)
        {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0)
            {
                tma_store_1d(args.out0 + ooff, sO0, (uint32_t)(ne * OUTL * 8));
                if (OP == QL_PD) tma_store_1d(args.out1 + ooff, sO1, (uint32_t)(ne * OUTL * 8));
            }
            tma_store_commit();
        }
        else
        {
            __syncwarp();
            for (int i = lane; i < ne * OUTL; i += 32)
            {
                const int a           = (i / OUTL) * OUTS + (i % OUTL);
                args.out0[ooff + i] = sO0[a];
                if (OP == QL_PD) args.out1[ooff + i] = sO1[a];
            }
        }
        __syncwarp();
    }
    tma_store_wait0();
}

template <int OP, int NM, bool DEF> static int quad_lane_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    using Cfg = QLaneCfg<OP, NM, DEF>;
    static int blocks_per_sm = 0;
    auto kern                = quad_lane_kernel<OP, NM, DEF>;
    if (blocks_per_sm == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("quad lane kernel <%d,%d,%d> does not fit on an SM", OP, NM, (int)DEF); return NEKMF_ERR_CUDA; }
        blocks_per_sm = nb;
    }
    QLaneArgs a;
    a.in = in[0]; a.in1 = in[1]; a.out0 = out[0]; a.out1 = out[1];
    const size_t gstep = DEF ? (size_t)op->nqTot : 1;
    a.jac = op->d_jac ? op->d_jac + (size_t)op->run_e0 * gstep : nullptr;
    a.df  = op->d_df ? op->d_df + (size_t)op->run_e0 * gstep : nullptr;
    a.dfStride = (size_t)op->nElmt * gstep;
    a.nElmt    = op->run_ne;
    uintptr_t al = (uintptr_t)in[0] | (uintptr_t)out[0];
    if (OP == QL_PD) al |= (uintptr_t)out[1];
    if (OP == QL_IPROD && DEF) al |= (uintptr_t)a.jac;
    if (OP == QL_IPWDB) al |= (uintptr_t)in[1];
    a.io_aligned = (al & 15) == 0;
    const int nBatches = (op->run_ne + 32 * Cfg::WARPS - 1) / (32 * Cfg::WARPS);
    int grid           = blocks_per_sm * NUM_SMS;
    if (grid > nBatches) grid = nBatches;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(*static_cast<const QLaneTab<NM> *>(op->kstate), a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <int NM> static bool quad_lane_install(nekmf_op_s *op)
{
    int kind = -1;
    if (op->optype == NEKMF_BWDTRANS) kind = QL_BWD;
    else if (op->optype == NEKMF_IPRODUCTWRTBASE) kind = QL_IPROD;
    else if (op->optype == NEKMF_PHYSDERIV && !op->deformed) kind = QL_PD;
    else if (op->optype == NEKMF_IPRODUCTWRTDERIVBASE && !op->deformed && NM <= 5) kind = QL_IPWDB; // nm 6, 7: shape_op_kernel<ipwdb> is 1.2-1.8x faster
    if (kind < 0) return false;
    auto *tab = new QLaneTab<NM>;
    memcpy(tab->B, op->b[0].data(), sizeof(tab->B));
    memcpy(tab->D, op->D[0].data(), sizeof(tab->D));
    memcpy(tab->w, op->ws[0].data(), sizeof(tab->w));
    memcpy(tab->dB, op->db[0].data(), sizeof(tab->dB));
    op->kstate      = tab;
    op->kstate_free = [](void *p) { delete static_cast<QLaneTab<NM> *>(p); };
    op->geo_pitch   = op->nqTot;
    const char *kn[4] = {"bwd", "iprod", "physderiv", "ipwdb"};
    char name[96];
    snprintf(name, sizeof(name), "quad_lane_kernel<%s,nm=%d,%s>", kn[kind], NM, op->deformed ? "deformed" : "regular");
    op->kname = name;
    if (kind == QL_BWD) op->launch = quad_lane_launch<QL_BWD, NM, false>;
    else if (kind == QL_PD) op->launch = quad_lane_launch<QL_PD, NM, false>;
    else if (kind == QL_IPWDB) op->launch = quad_lane_launch<QL_IPWDB, NM, false>;
    else op->launch = op->deformed ? quad_lane_launch<QL_IPROD, NM, true> : quad_lane_launch<QL_IPROD, NM, false>;
    return true;
}

// called first by select_shape_fast for quadrilaterals with the default quadrature
bool select_quad_lane(nekmf_op_s *op)
{
    if (op->shape != NEKMF_QUAD || op->nq[0] != op->nm[0] + 1 || op->nm[1] != op->nm[0] || op->nq[1] != op->nq[0] ||
        op->b[1] != op->b[0] || op->D[1] != op->D[0] || op->ws[1] != op->ws[0])
        return false;
    const char *v = getenv("NEKMF_QUAD_LANE"); // NEKMF_QUAD_LANE=0: CTA-level kernels of shape_kernels.cuh
    if (v && v[0] == '0') return false;
    switch (op->nm[0])
    {
        case 2: return quad_lane_install<2>(op);
        case 3: return quad_lane_install<3>(op);
        case 4: return quad_lane_install<4>(op);
        case 5: return quad_lane_install<5>(op);
        case 6: return quad_lane_install<6>(op);
        case 7: return quad_lane_install<7>(op); // nm = 8: the element no longer fits 255 registers (spills), CTA-level kernels
    }
    return false;
}

} // namespace nekmf
