// tri_lane.cu -- BwdTrans, IProductWRTBase, (regular) PhysDeriv and (regular) IProductWRTDerivBase on triangles
// with ONE LANE PER ELEMENT.
//
// Reference semantics: MatrixFreeOps/BwdTransKernels.hpp:78-126, IProductKernels.hpp:135-234 (incl. the CORRECT
// term of the singular vertex, mode 1), PhysDerivKernels.hpp:153-217 (collapsed-coordinate chain rule
// d/dxi_0 = 2/(1-eta_1) d/deta_0, d/dxi_1 = (1+eta_0)/(1-eta_1) d/deta_0 + d/deta_1).
//
// A triangle has nm(nm+1)/2 coefficients and (nm+1) nm quadrature values (Gauss-Lobatto x Gauss-Radau): both fit
// the registers of a single lane, so the collapsed sum-factorisation runs there with every table entry a
// kernel-parameter constant (the eModified_B rows depend on (p,q), which is a compile-time index after unrolling).
// Same batch / copy scheme as quad_lane.cu: warps are independent workers on 32 elements; odd-length blocks
// travel as one bulk TMA copy per batch, even-length blocks sit in padded slots filled by warp-wide 16-byte
// cp.async copies and drained by warp-wide 16-byte stores.
#include "hex_kernels.cuh"
#include "op_internal.h"
#include <stdlib.h>
#include <string.h>

namespace nekmf
{

enum { TL_BWD = 0, TL_IPROD = 1, TL_PD = 2, TL_IPWDB = 3 };

template <int NM> struct TLaneTab
{
    static constexpr int NQ0 = NM + 1, NQ1 = NM, NP = NM * (NM + 1) / 2;
    double b0[NM * NQ0];   // eModified_A, [p][i]
    double b1[NP * NQ1];   // eModified_B, [(p,q)][j], rows in mode order (p outer, q < nm-p)
    double D0[NQ0 * NQ0];  // D[k*nq+i] = dh_k/dz(z_i)
    double D1[NQ1 * NQ1];
    double w0[NQ0], w1[NQ1]; // w1 carries the 0.5 of the collapsed Jacobian (Operator.hpp:244-258)
    double h0[NQ0], h1[NQ1]; // 0.5 (1 + z0_i),  2 / (1 - z1_j)
    double db0[NM * NQ0];    // dbdata of both directions (IProductWRTDerivBase)
    double db1[NP * NQ1];
};

struct TLaneArgs
{
    const double *in;
    const double *in1; // second input (IProductWRTDerivBase)
    double *out0, *out1;
    const double *jac;
    const double *df;
    size_t dfStride;
    int nElmt;
    int io_aligned;
};

template <int OP, int NM, bool DEF> struct TLaneCfg
{
    static constexpr int NQ0 = NM + 1, NQ1 = NM, NQT = NQ0 * NQ1, NP = NM * (NM + 1) / 2;
    static constexpr int INL  = OP == TL_BWD ? NP : NQT;
    static constexpr int OUTL = (OP == TL_IPROD || OP == TL_IPWDB) ? NP : NQT;
    static constexpr bool INPAD = (INL % 2) == 0, OUTPAD = (OUTL % 2) == 0;
    // lane stride of a slot: odd lengths are conflict free as they are; even lengths must stay even (16-byte copies)
    // and are best at 2 (mod 4) doubles -- 2-way bank conflicts; a multiple of 4 would be 4- to 16-way
    static constexpr int INS  = (INPAD && INL % 4 == 0) ? INL + 2 : INL, OUTS = (OUTPAD && OUTL % 4 == 0) ? OUTL + 2 : OUTL;
    static constexpr int NIN  = 1 + (((OP == TL_IPROD && DEF) || OP == TL_IPWDB) ? 1 : 0);
    static constexpr int NOUT = OP == TL_PD ? 2 : 1;
    static constexpr int INB  = round_up(32 * INS, 2), OUTB = round_up(32 * OUTS, 2);
    static constexpr int PER_WARP = NIN * INB + NOUT * OUTB + 2;
    static constexpr int W_FIT  = (200 * 1024) / (PER_WARP * 8);
    static constexpr int WARPS  = W_FIT >= 16 ? 16 : (W_FIT >= 12 ? 12 : (W_FIT >= 8 ? 8 : (W_FIT >= 4 ? 4 : (W_FIT >= 1 ? W_FIT : 1))));
    static constexpr int T      = WARPS * 32;
    static constexpr size_t SMEM = (size_t)WARPS * PER_WARP * 8 + 16;
};

__device__ __forceinline__ void tl_cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void tl_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// first mode of block p (modes are ordered p outer, q < nm - p)
__host__ __device__ constexpr int tl_off(int p, int nm) { return p * nm - p * (p - 1) / 2; }

template <int OP, int NM, bool DEF>
__global__ void __launch_bounds__(TLaneCfg<OP, NM, DEF>::T, 1)
    tri_lane_kernel(const __grid_constant__ TLaneTab<NM> tab, const __grid_constant__ TLaneArgs args)
{
    using Cfg = TLaneCfg<OP, NM, DEF>;
    constexpr int NQ0 = Cfg::NQ0, NQ1 = Cfg::NQ1, NP = Cfg::NP, INL = Cfg::INL, OUTL = Cfg::OUTL, INS = Cfg::INS, OUTS = Cfg::OUTS;
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
        tma_store_wait_read0(); // the staging slots are free of the previous batch's bulk stores
        __syncwarp();
        if (fin && INPAD)
        {
            for (int i2 = lane; i2 < ne * INL / 2; i2 += 32)
            {
                const int a = pad_in(i2);
                tl_cp_async16(sIn + a, args.in + ioff + 2 * i2);
                if (Cfg::NIN == 2) tl_cp_async16(sJac + a, (OP == TL_IPWDB ? args.in1 : args.jac) + ioff + 2 * i2);
            }
            tl_cp_async_wait_all();
        }
        else if (fin)
        {
            if (lane == 0)
            {
                const uint32_t bytes = (uint32_t)(ne * INL * 8);
                mbar_expect_tx(bar, bytes * Cfg::NIN);
                tma_load_1d(sIn, args.in + ioff, bytes, bar);
                if (Cfg::NIN == 2) tma_load_1d(sJac, (OP == TL_IPWDB ? args.in1 : args.jac) + ioff, bytes, bar);
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
                if (Cfg::NIN == 2) sJac[a] = __ldg((OP == TL_IPWDB ? args.in1 : args.jac) + ioff + i);
            }
        }
        __syncwarp();

        if (lane < ne)
        {
            const double *xe = sIn + lane * INS;
            double *o0       = sO0 + lane * OUTS;
            const size_t eg  = (size_t)b * 32 + lane;
            if (OP == TL_BWD)
            {
                double c[NP];
#pragma unroll
                for (int m = 0; m < NP; ++m) c[m] = xe[m];
#pragma unroll
                for (int j = 0; j < NQ1; ++j)
                {
                    double fp[NM];
#pragma unroll
                    for (int p = 0; p < NM; ++p)
                    {
                        double s = tab.b1[tl_off(p, NM) * NQ1 + j] * c[tl_off(p, NM)];
#pragma unroll
                        for (int q = 1; q < NM - p; ++q) s = fma(tab.b1[(tl_off(p, NM) + q) * NQ1 + j], c[tl_off(p, NM) + q], s);
                        fp[p] = s;
                    }
                    const double corr = c[1] * tab.b1[NQ1 + j]; // CORRECT: singular vertex (mode 1) x b0[1][i]
#pragma unroll
                    for (int i = 0; i < NQ0; ++i)
                    {
                        double s = tab.b0[i] * fp[0];
#pragma unroll
                        for (int p = 1; p < NM; ++p) s = fma(tab.b0[p * NQ0 + i], fp[p], s);
                        o0[j * NQ0 + i] = fma(corr, tab.b0[NQ0 + i], s);
                    }
                }
            }
            else if (OP == TL_IPROD)
            {
                double g[NQ1][NQ0];
                const double jr  = DEF ? 1.0 : __ldg(args.jac + eg);
                const double *je = sJac + lane * INS;
#pragma unroll
                for (int j = 0; j < NQ1; ++j)
#pragma unroll
                    for (int i = 0; i < NQ0; ++i)
                        g[j][i] = xe[j * NQ0 + i] * ((DEF ? je[j * NQ0 + i] : jr) * (tab.w1[j] * tab.w0[i]));
                double t1[NQ1]; // the p = 1 line, needed again by the CORRECT term
#pragma unroll
                for (int p = 0; p < NM; ++p)
                {
                    double t[NQ1];
#pragma unroll
                    for (int j = 0; j < NQ1; ++j)
                    {
                        double s = tab.b0[p * NQ0] * g[j][0];
#pragma unroll
                        for (int i = 1; i < NQ0; ++i) s = fma(tab.b0[p * NQ0 + i], g[j][i], s);
                        t[j] = s;
                        if (p == 1) t1[j] = s;
                    }
#pragma unroll
                    for (int q = 0; q < NM - p; ++q)
                    {
                        double s = tab.b1[(tl_off(p, NM) + q) * NQ1] * t[0];
#pragma unroll
                        for (int j = 1; j < NQ1; ++j) s = fma(tab.b1[(tl_off(p, NM) + q) * NQ1 + j], t[j], s);
                        o0[tl_off(p, NM) + q] = s;
                    }
                }
                if (NM > 1)
                {
                    // CORRECT: mode 1 also collects the p = 1 line against b1 row 1
                    double s = o0[1];
#pragma unroll
                    for (int j = 0; j < NQ1; ++j) s = fma(tab.b1[NQ1 + j], t1[j], s);
                    o0[1] = s;
                }
            }
            else if (OP == TL_IPWDB)
            {
                // IProductWRTDerivBase.h:891-1060 (regular): t_d = df[d] f_0 + df[2+d] f_1, collapsed chain rule on t_0,
                // out = IPTri(dB0, B1)[t_0] + IPTri(B0, dB1)[t_1], each with its CORRECT term
                const double *ye = sJac + lane * INS;
                const double jr  = __ldg(args.jac + eg);
                const double f0 = __ldg(args.df + eg), f1 = __ldg(args.df + args.dfStride + eg),
                             f2 = __ldg(args.df + 2 * args.dfStride + eg), f3 = __ldg(args.df + 3 * args.dfStride + eg);
#pragma unroll
                for (int d = 0; d < 2; ++d)
                {
                    double g[NQ1][NQ0];
#pragma unroll
                    for (int j = 0; j < NQ1; ++j)
#pragma unroll
                        for (int i = 0; i < NQ0; ++i)
                        {
                            const double x = xe[j * NQ0 + i], y = ye[j * NQ0 + i];
                            const double t1 = f1 * x + f3 * y;
                            double t        = t1;
                            if (d == 0)
                            {
                                const double t0 = f0 * x + f2 * y;
                                t               = t0 * tab.h1[j] + (tab.h0[i] * t1) * tab.h1[j];
                            }
                            g[j][i] = t * (jr * (tab.w1[j] * tab.w0[i]));
                        }
                    double tl1[NQ1];
#pragma unroll
                    for (int p = 0; p < NM; ++p)
                    {
                        double t[NQ1];
#pragma unroll
                        for (int j = 0; j < NQ1; ++j)
                        {
                            double s = (d == 0 ? tab.db0[p * NQ0] : tab.b0[p * NQ0]) * g[j][0];
#pragma unroll
                            for (int i = 1; i < NQ0; ++i) s = fma(d == 0 ? tab.db0[p * NQ0 + i] : tab.b0[p * NQ0 + i], g[j][i], s);
                            t[j] = s;
                            if (p == 1) tl1[j] = s;
                        }
#pragma unroll
                        for (int q = 0; q < NM - p; ++q)
                        {
                            const int m = tl_off(p, NM) + q;
                            double s    = (d == 1 ? tab.db1[m * NQ1] : tab.b1[m * NQ1]) * t[0];
#pragma unroll
                            for (int j = 1; j < NQ1; ++j) s = fma(d == 1 ? tab.db1[m * NQ1 + j] : tab.b1[m * NQ1 + j], t[j], s);
                            if (d == 0) o0[m] = s;
                            else o0[m] += s;
                        }
                    }
                    if (NM > 1)
                    {
                        double s = o0[1];
#pragma unroll
                        for (int j = 0; j < NQ1; ++j) s = fma(d == 1 ? tab.db1[NQ1 + j] : tab.b1[NQ1 + j], tl1[j], s);
                        o0[1] = s;
                    }
                }
            }
            else
            {
                double u[NQ1][NQ0];
#pragma unroll
                for (int j = 0; j < NQ1; ++j)
#pragma unroll
                    for (int i = 0; i < NQ0; ++i) u[j][i] = xe[j * NQ0 + i];
                const double f0 = __ldg(args.df + eg), f1 = __ldg(args.df + args.dfStride + eg),
                             f2 = __ldg(args.df + 2 * args.dfStride + eg), f3 = __ldg(args.df + 3 * args.dfStride + eg);
                double *o1 = sO1 + lane * OUTS;
#pragma unroll
                for (int j = 0; j < NQ1; ++j)
#pragma unroll
                    for (int i = 0; i < NQ0; ++i)
                    {
                        double d0 = tab.D0[i] * u[j][0], d1 = tab.D1[j] * u[0][i];
#pragma unroll
                        for (int m = 1; m < NQ0; ++m) d0 = fma(tab.D0[m * NQ0 + i], u[j][m], d0);
#pragma unroll
                        for (int m = 1; m < NQ1; ++m) d1 = fma(tab.D1[m * NQ1 + j], u[m][i], d1);
                        const double a  = tab.h1[j] * d0;
                        const double bb = fma(a, tab.h0[i], d1);
                        o0[j * NQ0 + i] = a * f0 + bb * f1;
                        o1[j * NQ0 + i] = a * f2 + bb * f3;
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
                if (OP == TL_PD)
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
                if (OP == TL_PD) tma_store_1d(args.out1 + ooff, sO1, (uint32_t)(ne * OUTL * 8));
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
                if (OP == TL_PD) args.out1[ooff + i] = sO1[a];
            }
        }
        __syncwarp();
    }
    tma_store_wait0();
}

template <int OP, int NM, bool DEF> static int tri_lane_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    using Cfg = TLaneCfg<OP, NM, DEF>;
    static int blocks_per_sm = 0;
    auto kern                = tri_lane_kernel<OP, NM, DEF>;
    if (blocks_per_sm == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("tri lane kernel <%d,%d,%d> does not fit on an SM", OP, NM, (int)DEF); return NEKMF_ERR_CUDA; }
        blocks_per_sm = nb;
    }
    TLaneArgs a;
    a.in = in[0]; a.in1 = in[1]; a.out0 = out[0]; a.out1 = out[1];
    const size_t gstep = DEF ? (size_t)op->nqTot : 1;
    a.jac = op->d_jac ? op->d_jac + (size_t)op->run_e0 * gstep : nullptr;
    a.df  = op->d_df ? op->d_df + (size_t)op->run_e0 * gstep : nullptr;
    a.dfStride = (size_t)op->nElmt * gstep;
    a.nElmt    = op->run_ne;
    uintptr_t al = (uintptr_t)in[0] | (uintptr_t)out[0];
    if (OP == TL_PD) al |= (uintptr_t)out[1];
    if (OP == TL_IPROD && DEF) al |= (uintptr_t)a.jac;
    if (OP == TL_IPWDB) al |= (uintptr_t)in[1];
    a.io_aligned = (al & 15) == 0;
    const int nBatches = (op->run_ne + 32 * Cfg::WARPS - 1) / (32 * Cfg::WARPS);
    int grid           = blocks_per_sm * NUM_SMS;
    if (grid > nBatches) grid = nBatches;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(*static_cast<const TLaneTab<NM> *>(op->kstate), a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <int NM> static bool tri_lane_install(nekmf_op_s *op)
{
    using Tab = TLaneTab<NM>;
    int kind = -1;
    if (op->optype == NEKMF_BWDTRANS) kind = TL_BWD;
    else if (op->optype == NEKMF_IPRODUCTWRTBASE) kind = TL_IPROD;
    else if (op->optype == NEKMF_PHYSDERIV && !op->deformed) kind = TL_PD;
    else if (op->optype == NEKMF_IPRODUCTWRTDERIVBASE && !op->deformed && NM <= 6) kind = TL_IPWDB; // nm 7: shape_op_kernel<ipwdb> is as fast
    if (kind < 0) return false;
    if (op->rows[1] != Tab::NP || op->nq[1] != Tab::NQ1) return false;
    auto *tab = new Tab;
    memcpy(tab->b0, op->b[0].data(), sizeof(tab->b0));
    memcpy(tab->b1, op->b[1].data(), sizeof(tab->b1));
    memcpy(tab->D0, op->D[0].data(), sizeof(tab->D0));
    memcpy(tab->D1, op->D[1].data(), sizeof(tab->D1));
    memcpy(tab->w0, op->ws[0].data(), sizeof(tab->w0));
    memcpy(tab->w1, op->ws[1].data(), sizeof(tab->w1));
    memcpy(tab->db0, op->db[0].data(), sizeof(tab->db0));
    memcpy(tab->db1, op->db[1].data(), sizeof(tab->db1));
    for (int i = 0; i < Tab::NQ0; ++i) tab->h0[i] = 0.5 * (1.0 + op->Z[0][i]);
    for (int j = 0; j < Tab::NQ1; ++j) tab->h1[j] = 2.0 / (1.0 - op->Z[1][j]);
    op->kstate      = tab;
    op->kstate_free = [](void *p) { delete static_cast<Tab *>(p); };
    op->geo_pitch   = op->nqTot;
    const char *kn[4] = {"bwd", "iprod", "physderiv", "ipwdb"};
    char name[96];
    snprintf(name, sizeof(name), "tri_lane_kernel<%s,nm=%d,%s>", kn[kind], NM, op->deformed ? "deformed" : "regular");
    op->kname = name;
    if (kind == TL_BWD) op->launch = tri_lane_launch<TL_BWD, NM, false>;
    else if (kind == TL_PD) op->launch = tri_lane_launch<TL_PD, NM, false>;
    else if (kind == TL_IPWDB) op->launch = tri_lane_launch<TL_IPWDB, NM, false>;
    else op->launch = op->deformed ? tri_lane_launch<TL_IPROD, NM, true> : tri_lane_launch<TL_IPROD, NM, false>;
    return true;
}

// called first by select_shape_fast for triangles with the default quadrature (nq0 = nm + 1, nq1 = nm)
bool select_tri_lane(nekmf_op_s *op)
{
    if (op->shape != NEKMF_TRI || op->nq[0] != op->nm[0] + 1 || op->nm[1] != op->nm[0]) return false;
    const char *v = getenv("NEKMF_TRI_LANE"); // NEKMF_TRI_LANE=0: CTA-level kernels of shape_kernels.cuh
    if (v && v[0] == '0') return false;
    switch (op->nm[0])
    {
        case 2: return tri_lane_install<2>(op);
        case 3: return tri_lane_install<3>(op);
        case 4: return tri_lane_install<4>(op);
        case 5: return tri_lane_install<5>(op);
        case 6: return tri_lane_install<6>(op);
        case 7: return tri_lane_install<7>(op);
    }
    return false;
}

} // namespace nekmf
