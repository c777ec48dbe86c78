// tet_dmma.cu -- BwdTrans and IProductWRTBase on tetrahedra at nm = 3..7 (default quadrature nq = (nm+1, nm, nm)) with
// FP64 tensor-core tiles (DMMA, mma.sync.m8n8k4.f64) for the one tensor-product contraction and a lane-per-mode-pair
// scheme for the two collapsed ones.
//
// Reference semantics: MatrixFreeOps/BwdTransKernels.hpp:374-484 (BwdTransTetKernel, CORRECT = true for the modified
// basis), IProductKernels.hpp:600-761 (IProductTetKernel); results differ from shape_kernels.cuh by summation order only.
//
// phi_pqr = A_p(xi_0) B_pq(xi_1) C_pqr(xi_2): only the xi_0 contraction is a plain tensor contraction.
//   BwdTrans:  f[(p,q)][k]  = sum_r c[pqr] C_pqr(k)          lane = mode pair (p,q) (nm(nm+1)/2 <= 28 lanes), r-line of
//                                                             the pair in registers, result to a per-warp scratch
//              g_k[p][j]    = sum_q f[(p,q)][k] B_pq(j)       lane (g, t) builds ITS A-operand entries (p = t | 4 + t,
//                                                             j = g) from the scratch; its B_pq(g) values are registers
//              out[k][j][i] = sum_p g_k[p][j] A_p(i)          DMMA, B operand = fixed fragments of A; 16-byte stores
//   IProduct:  s_k[p][j]    = sum_i A_p(i) (w J F)[k][j][i]   DMMA, result to a per-warp scratch
//              h[(p,q)][k]  = sum_j B_pq(j) s_k[p][j]         lane = mode pair
//              out[pqr]     = sum_k C_pqr(k) h[(p,q)][k]      same lane, r-line staged for coalesced stores
// The corrections of the modified basis (top vertex, bottom vertex, singular edge: all expressed through the rows
// B_00, B_01 and C_0, C_1.. of the collapsed tables) are computed by two otherwise idle lanes.
// Every warp is an independent worker: elements (pairs where one block is an odd number of doubles) arrive by TMA bulk
// copies into the warp's own double buffer.
#include "op_internal.h"
#include <stdlib.h>
#include <string.h>

namespace nekmf
{

template <int NM> struct TetDmmaTab
{
    static constexpr int NQ0 = NM + 1, NQ1 = NM, NQ2 = NM, NPAIR = NM * (NM + 1) / 2, NMT = NM * (NM + 1) * (NM + 2) / 6;
    double b0[NM * NQ0];    // [p][i]
    double b1[NPAIR * NQ1]; // rows (p, q)
    double b2[NMT * NQ2];   // rows (p, q, r)
    double w0[NQ0], w1[NQ1], w2[NQ2]; // weights with the collapsed-coordinate factors folded in
    int pairP[32], pairStart[32], pairLen[32]; // per mode pair: p, first mode, number of modes (nm - p - q)
};

struct TetDmmaArgs
{
    const double *in;
    double *out;
    const double *jac; // IProduct: [nElmt] | [nElmt][nqTot]
    int nElmt;
    int in_aligned, out_aligned; // 16-byte aligned
};

__device__ __forceinline__ void tt_mma(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// OP 0: BwdTrans, 1: IProductWRTBase
template <int OP, int NM, bool DEF> struct TetDmmaCfg
{
    static constexpr int NQ0 = NM + 1, NQ1 = NM, NQ2 = NM, NPAIR = NM * (NM + 1) / 2, NMT = NM * (NM + 1) * (NM + 2) / 6;
    static constexpr int NQT = NQ0 * NQ1 * NQ2;
    static constexpr int IN_EL = OP == 0 ? NMT : NQT;
    static constexpr int EPB   = (IN_EL % 2) ? 2 : 1; // elements per buffer: a whole number of 16-byte units
    static constexpr int BUF   = EPB * IN_EL;
    static constexpr int SLOT  = BUF + ((OP == 1 && DEF) ? EPB * NQT : 0);
    // per-warp scratch: BwdTrans f[(pair | X1 | X2)][k]; IProduct s[k][p (8)][j (8)] and the output staging block
    static constexpr int SCR   = OP == 0 ? (((NPAIR + 2) * NQ2 + 1) & ~1) : (NQ2 * 64 + ((NMT + 1) & ~1));
    static constexpr int PER_WARP = 2 * SLOT + SCR + 2;
    static constexpr int B2S   = (NMT * NQ2 + 1) & ~1; // shared copy of the xi_2 table
    static constexpr int WARPS = 8, T = WARPS * 32;
    static constexpr size_t SMEM = (size_t)(B2S + WARPS * PER_WARP) * 8 + 16;
};

template <int OP, int NM, bool DEF>
__global__ void __launch_bounds__(TetDmmaCfg<OP, NM, DEF>::T, 2)
    tet_dmma_kernel(const __grid_constant__ TetDmmaTab<NM> tab, const __grid_constant__ TetDmmaArgs args)
{
    using Cfg = TetDmmaCfg<OP, NM, DEF>;
    constexpr int NQ0 = Cfg::NQ0, NQ1 = Cfg::NQ1, NQ2 = Cfg::NQ2, NPAIR = Cfg::NPAIR, NMT = Cfg::NMT, NQT = Cfg::NQT;
    constexpr int IN_EL = Cfg::IN_EL, BUF = Cfg::BUF, SLOT = Cfg::SLOT, EPB = Cfg::EPB;
    constexpr bool JSM = OP == 1 && DEF;
    static_assert(NPAIR + 2 <= 32, "one lane per mode pair plus two correction lanes");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = lane & 3, g = lane >> 2;
    double *sB2   = reinterpret_cast<double *>(smem_raw);
    double *wbase = sB2 + Cfg::B2S + (size_t)warp * Cfg::PER_WARP;
    double *sScr  = wbase + 2 * SLOT;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sScr + Cfg::SCR);

    // xi_2 table (IProduct: weights are applied to the input instead)
    for (int i = threadIdx.x; i < NMT * NQ2; i += Cfg::T) sB2[i] = tab.b2[i];
    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int nBlk = (args.nElmt + EPB - 1) / EPB; // blocks of EPB elements
    const int GW = gridDim.x * Cfg::WARPS, gw = blockIdx.x * Cfg::WARPS + warp;
    auto blk_ne = [&](int b) { return args.nElmt - EPB * b >= EPB ? EPB : 1; };
    auto tma_ok = [&](int b) { return args.in_aligned && blk_ne(b) == EPB; };
    auto issue  = [&](int b, int slot) { // lane 0
        if (!tma_ok(b)) return;
        mbar_expect_tx(bar + slot, (uint32_t)((BUF + (JSM ? EPB * NQT : 0)) * 8));
        tma_load_1d(wbase + slot * SLOT, args.in + (size_t)b * BUF, (uint32_t)(BUF * 8), bar + slot);
        if (JSM) tma_load_1d(wbase + slot * SLOT + BUF, args.jac + (size_t)b * EPB * NQT, (uint32_t)(EPB * NQT * 8), bar + slot);
    };

    // ---- lane roles of the pair phase: lanes 0..NPAIR-1 own one mode pair, lanes NPAIR, NPAIR+1 the corrections.
    //   the modes of a lane are in[ibase + r], r < len, and meet the xi_2 rows rbase + r (+1 from r = 1 on for the
    //   singular-edge lane: modes (0,1,r) use the rows of (0,0,r+1), the bottom vertex (0,1,0) row 0)
    const bool is_pair = lane < NPAIR, is_top = lane == NPAIR, is_edge = lane == NPAIR + 1;
    const int len   = is_pair ? tab.pairLen[lane] : (is_top ? 1 : (is_edge ? NM - 1 : 0));
    const int ibase = is_pair ? tab.pairStart[lane] : (is_top ? 1 : NM);
    const int rbase = is_pair ? tab.pairStart[lane] : (is_top ? 1 : 0);
    const int bump  = is_edge ? 1 : 0;
    const int myp   = is_pair ? tab.pairP[lane] : (is_top ? 0 : 1);
    auto xi2row = [&](int r) { const int row = rbase + r + (r >= 1 ? bump : 0); return row < NMT ? row : NMT - 1; };
    auto mpr    = [](int p) { return p * NM - p * (p - 1) / 2; }; // pair index of (p, 0)

    // ---- fragment phase constants
    const int p0 = t, p1 = 4 + t;
    double Bf[2]; // fixed operand fragments of A_p(i): BwdTrans B[kk = p][n = i = g]; IProduct A[m = p = g][kk = i]
#pragma unroll
    for (int s = 0; s < 2; ++s)
    {
        const int con = 4 * s + t;
        double v = 0.0;
        if (OP == 0) { if (con < NM && g < NQ0) v = tab.b0[con * NQ0 + g]; }
        else { if (g < NM && con < NQ0) v = tab.b0[g * NQ0 + con]; }
        Bf[s] = v;
    }
    // BwdTrans: B_pq(j = g) of the lane's two p; the correction coefficients of its p
    double b1r0[NM], b1r1[NM > 4 ? NM - 4 : 1], cA = 0.0, cB = 0.0;
    // IProduct: the pair lane's B row with the xi_1 part of the correction rows
    double b1row[NQ1], b1row2[NQ1];
    if (OP == 0)
    {
#pragma unroll
        for (int q = 0; q < NM; ++q) b1r0[q] = (p0 < NM && q < NM - p0 && g < NQ1) ? tab.b1[(mpr(p0) + q) * NQ1 + g] : 0.0;
#pragma unroll
        for (int q = 0; q < NM - 4; ++q) b1r1[q] = (p1 < NM && q < NM - p1 && g < NQ1) ? tab.b1[(mpr(p1) + q) * NQ1 + g] : 0.0;
        if (g < NQ1)
        {
            const double r0 = tab.b1[g], r1 = tab.b1[NQ1 + g];
            if (t == 0) cA = r1;
            if (t == 1) { cA = r0 + r1; cB = r1; }
        }
    }
    else
    {
#pragma unroll
        for (int j = 0; j < NQ1; ++j)
        {
            const double r0 = tab.b1[j], r1 = tab.b1[NQ1 + j];
            b1row[j]  = is_pair ? tab.b1[lane * NQ1 + j] : ((is_top || is_edge) ? r1 : 0.0);
            b1row2[j] = is_top ? r0 + r1 : 0.0; // top vertex: p = 1 part
        }
    }

    uint32_t phase[2] = {0, 0};
    int slot          = 0;
    if (lane == 0 && gw < nBlk) issue(gw, 0);

    for (int b = gw; b < nBlk; b += GW, slot ^= 1)
    {
        const int ne = blk_ne(b);
        double *sIn  = wbase + slot * SLOT;
        if (lane == 0 && b + GW < nBlk) issue(b + GW, slot ^ 1); // the other buffer was consumed one trip ago
        double jpre = 1.0;
        if (OP == 1 && !DEF) jpre = __ldg(args.jac + (size_t)b * EPB);
        if (tma_ok(b))
        {
            mbar_wait(bar + slot, phase[slot]);
            phase[slot] ^= 1;
        }
        else
        {
            const double *src = args.in + (size_t)b * BUF;
            for (int i = lane; i < ne * IN_EL; i += 32) sIn[i] = __ldg(src + i);
            if (JSM)
            {
                const double *sj = args.jac + (size_t)b * EPB * NQT;
                for (int i = lane; i < ne * NQT; i += 32) sIn[BUF + i] = __ldg(sj + i);
            }
        }
        __syncwarp();

#pragma unroll 1
        for (int e = 0; e < ne; ++e)
        {
            const size_t el = (size_t)b * EPB + e;
            const double *U = sIn + e * IN_EL;
            if (OP == 0)
            {
                // ---------------------------------------------------------------- BwdTrans: c[pqr] -> out[k][j][i]
                // pair phase: f[(p,q)][k] (and the two correction lines) into the scratch
                {
                    double cin[NM];
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                    {
                        const bool ok  = r < len;
                        const double x = U[ok ? ibase + r : 0];
                        cin[r]         = ok ? x : 0.0;
                    }
                    if (lane < NPAIR + 2)
                    {
#pragma unroll
                        for (int k = 0; k < NQ2; ++k)
                        {
                            double f = cin[0] * sB2[xi2row(0) * NQ2 + k];
#pragma unroll
                            for (int r = 1; r < NM; ++r) f = fma(cin[r], sB2[xi2row(r) * NQ2 + k], f);
                            sScr[lane * NQ2 + k] = f;
                        }
                    }
                }
                __syncwarp();
                // fragment phase
                double *o = args.out + el * NQT + g * NQ0 + 2 * t;
                const int pb0 = p0 < NM ? mpr(p0) : 0, pb1 = p1 < NM ? mpr(p1) : 0;
#pragma unroll
                for (int k = 0; k < NQ2; ++k)
                {
                    // A-operand entries g_k[p][j = g], p = t | 4 + t, with the corrections of p = 0, 1
                    double a0 = fma(sScr[NPAIR * NQ2 + k], cA, sScr[(NPAIR + 1) * NQ2 + k] * cB), a1 = 0.0;
#pragma unroll
                    for (int q = 0; q < NM; ++q)
                    {
                        const int row = pb0 + q < NPAIR ? pb0 + q : NPAIR - 1; // clamped: b1r0[q] is zero there
                        a0            = fma(sScr[row * NQ2 + k], b1r0[q], a0);
                    }
#pragma unroll
                    for (int q = 0; q < NM - 4; ++q)
                    {
                        const int row = pb1 + q < NPAIR ? pb1 + q : NPAIR - 1;
                        a1            = fma(sScr[row * NQ2 + k], b1r1[q], a1);
                    }
                    double d0 = 0.0, d1 = 0.0;
                    tt_mma(d0, d1, a0, Bf[0]); // C[j = g][i = 2t, 2t + 1]
                    tt_mma(d0, d1, a1, Bf[1]);
                    if (g < NQ1 && 2 * t < NQ0)
                    {
                        if ((NQ0 % 2 == 0) && args.out_aligned)
                            *reinterpret_cast<double2 *>(o + k * (NQ0 * NQ1)) = make_double2(d0, d1);
                        else
                        {
                            o[k * (NQ0 * NQ1)] = d0;
                            if (2 * t + 1 < NQ0) o[k * (NQ0 * NQ1) + 1] = d1;
                        }
                    }
                }
                __syncwarp(); // scratch free for the next element
            }
            else
            {
                // ---------------------------------------------------------------- IProductWRTBase: F[k][j][i] -> out[pqr]
                double *sS   = sScr;             // [k][p (8)][j (8)]
                double *sStg = sScr + NQ2 * 64;  // [NMT]
                const double *Jp  = sIn + BUF + e * NQT; // DEF only
                const double jreg = DEF ? 1.0 : (e == 0 ? jpre : __ldg(args.jac + el));
                // fragment phase: s_k[p = g][j = 2t, 2t+1] = sum_i A_p(i) (w J F)[k][j][i]; the lane supplies (j = g, i = t | 4 + t)
                {
                    const bool v0 = g < NQ1 && t < NQ0, v1 = g < NQ1 && 4 + t < NQ0;
                    const int ix0 = v0 ? g * NQ0 + t : 0, ix1 = v1 ? g * NQ0 + 4 + t : 0;
                    const double wj  = g < NQ1 ? tab.w1[g] * jreg : 0.0;
                    const double w00 = v0 ? tab.w0[t] * wj : 0.0, w01 = v1 ? tab.w0[4 + t < NQ0 ? 4 + t : 0] * wj : 0.0;
#pragma unroll
                    for (int k = 0; k < NQ2; ++k)
                    {
                        double f0 = U[k * (NQ0 * NQ1) + ix0] * (w00 * tab.w2[k]), f1 = U[k * (NQ0 * NQ1) + ix1] * (w01 * tab.w2[k]);
                        if (DEF)
                        {
                            f0 *= Jp[k * (NQ0 * NQ1) + ix0];
                            f1 *= Jp[k * (NQ0 * NQ1) + ix1];
                        }
                        double d0 = 0.0, d1 = 0.0;
                        tt_mma(d0, d1, Bf[0], f0);
                        tt_mma(d0, d1, Bf[1], f1);
                        *reinterpret_cast<double2 *>(sS + k * 64 + g * 8 + 2 * t) = make_double2(d0, d1);
                    }
                }
                __syncwarp();
                // pair phase: h[k] = sum_j B(j) s_k[p][j], then the xi_2 contraction for the lane's modes
                double acc[NM];
#pragma unroll
                for (int r = 0; r < NM; ++r) acc[r] = 0.0;
                if (lane < NPAIR + 2)
                {
#pragma unroll
                    for (int k = 0; k < NQ2; ++k)
                    {
                        const double *sp = sS + k * 64 + myp * 8;
                        double h = b1row[0] * sp[0];
#pragma unroll
                        for (int j = 1; j < NQ1; ++j) h = fma(b1row[j], sp[j], h);
                        if (is_top)
                        {
                            // top vertex: the p = 1 rows with B_00 + B_01
#pragma unroll
                            for (int j = 0; j < NQ1; ++j) h = fma(b1row2[j], sp[8 + j], h);
                        }
#pragma unroll
                        for (int r = 0; r < NM; ++r) acc[r] = fma(sB2[xi2row(r) * NQ2 + k], h, acc[r]);
                    }
                }
                if (is_pair)
                {
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                        if (r < len) sStg[ibase + r] = acc[r];
                }
                __syncwarp();
                if (is_top || is_edge)
                {
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                        if (r < len) sStg[ibase + r] += acc[r];
                }
                __syncwarp();
                {
                    double *o = args.out + el * NMT;
                    if ((NMT % 2 == 0) && args.out_aligned)
                        for (int i2 = lane; i2 < NMT / 2; i2 += 32)
                            *reinterpret_cast<double2 *>(o + 2 * i2) = *reinterpret_cast<const double2 *>(sStg + 2 * i2);
                    else
                        for (int i = lane; i < NMT; i += 32) o[i] = sStg[i];
                }
                __syncwarp(); // scratch and staging free for the next element
            }
        }
        __syncwarp(); // every lane is done with this buffer before lane 0 refills it
    }
}

template <int NM> struct TetDmmaState
{
    TetDmmaTab<NM> tab;
    void *fallback_state          = nullptr;
    void (*fallback_free)(void *) = nullptr;
    int bps[2][2]                 = {{0, 0}, {0, 0}};
};

template <int OP, int NM, bool DEF> static int tet_dmma_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    auto *st  = static_cast<TetDmmaState<NM> *>(op->kstate);
    using Cfg = TetDmmaCfg<OP, NM, DEF>;
    auto kern = tet_dmma_kernel<OP, NM, DEF>;
    int &bps  = st->bps[OP][DEF ? 1 : 0];
    if (bps == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("tet DMMA kernel <%d> does not fit on an SM", NM); return NEKMF_ERR_CUDA; }
        bps = nb;
    }
    TetDmmaArgs a;
    a.in = in[0]; a.out = out[0]; a.nElmt = op->run_ne;
    a.jac = nullptr;
    if (OP == 1) a.jac = DEF ? op->d_jac + (size_t)op->run_e0 * op->geo_pitch : op->d_jac + op->run_e0;
    a.in_aligned  = ((((uintptr_t)in[0]) | (OP == 1 && DEF ? (uintptr_t)a.jac : 0)) & 15) == 0;
    a.out_aligned = (((uintptr_t)out[0]) & 15) == 0;
    const int nBlk = (op->run_ne + Cfg::EPB - 1) / Cfg::EPB;
    int grid       = bps * NUM_SMS;
    const int need = (nBlk + Cfg::WARPS - 1) / Cfg::WARPS;
    if (grid > need) grid = need;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(st->tab, a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <int NM> static void tet_dmma_wrap(nekmf_op_s *op)
{
    auto *st = new TetDmmaState<NM>;
    memcpy(st->tab.b0, op->b[0].data(), sizeof(st->tab.b0));
    memcpy(st->tab.b1, op->b[1].data(), sizeof(st->tab.b1));
    memcpy(st->tab.b2, op->b[2].data(), sizeof(st->tab.b2));
    memcpy(st->tab.w0, op->ws[0].data(), sizeof(st->tab.w0));
    memcpy(st->tab.w1, op->ws[1].data(), sizeof(st->tab.w1));
    memcpy(st->tab.w2, op->ws[2].data(), sizeof(st->tab.w2));
    int pi = 0, mode = 0;
    for (int i = 0; i < 32; ++i) st->tab.pairP[i] = st->tab.pairStart[i] = st->tab.pairLen[i] = 0;
    for (int p = 0; p < NM; ++p)
        for (int q = 0; q < NM - p; ++q, ++pi)
        {
            st->tab.pairP[pi]     = p;
            st->tab.pairStart[pi] = mode;
            st->tab.pairLen[pi]   = NM - p - q;
            mode += NM - p - q;
        }
    st->fallback_state = op->kstate;
    st->fallback_free  = op->kstate_free;
    op->kstate         = st;
    op->kstate_free    = [](void *p) {
        auto *s = static_cast<TetDmmaState<NM> *>(p);
        if (s->fallback_state && s->fallback_free) s->fallback_free(s->fallback_state);
        delete s;
    };
    char name[128];
    if (op->optype == NEKMF_BWDTRANS)
    {
        op->launch = tet_dmma_launch<0, NM, false>;
        snprintf(name, sizeof(name), "tet_dmma_kernel<bwd,nm=%d>(DMMA m8n8k4 + lane per mode pair)", NM);
    }
    else
    {
        op->launch = op->deformed ? tet_dmma_launch<1, NM, true> : tet_dmma_launch<1, NM, false>;
        snprintf(name, sizeof(name), "tet_dmma_kernel<iprod,nm=%d,%s>(DMMA m8n8k4 + lane per mode pair)", NM, op->deformed ? "deformed" : "regular");
    }
    op->kname = name;
}

// called from select_shape_fast after the pencil launcher is installed: default quadrature, BwdTrans or
// IProductWRTBase on tetrahedra.  NEKMF_TET_DMMA=0 keeps the pencil kernels (the other arm of the A/B), =all takes the
// tensor-core kernels at every instantiated order.
void tet_dmma_maybe_wrap(nekmf_op_s *op)
{
    if (op->shape != NEKMF_TET) return;
    if (op->optype != NEKMF_BWDTRANS && op->optype != NEKMF_IPRODUCTWRTBASE) return;
    const int nm = op->nm[0];
    if (op->nm[1] != nm || op->nm[2] != nm || op->nq[0] != nm + 1 || op->nq[1] != nm || op->nq[2] != nm) return;
    if (op->deformed && op->geo_pitch != op->nqTot) return;
    const char *v = getenv("NEKMF_TET_DMMA");
    if (v && v[0] == '0') return;
    // measured A/B (profiles/r02_sweep_tet_dmma_*.jsonl, fraction of the HBM peak, this kernel / pencil kernel):
    //   nm = 7: BwdTrans 0.37 / 0.28, IProductWRTBase 0.28 / 0.25 (regular), 0.41 / 0.33 (deformed)
    //   nm = 5, 6: BwdTrans 0.34 / 0.28, 0.37 / 0.35; IProductWRTBase within 5 % either way (pencil kept)
    // both are bound by shared-memory wavefronts of the collapsed contractions, not by the tensor pipe
    // (tet_gemm.cu takes over where it is instantiated and faster; this policy covers what it leaves: deformed
    // IProductWRTBase at nm = 6, and everything when NEKMF_TET_GEMM=0)
    const bool faster = nm == 7 || ((nm == 5 || nm == 6) && op->optype == NEKMF_BWDTRANS) || (nm == 6 && op->deformed);
    if (!(v && v[0] == 'a') && !faster) return;
    switch (nm)
    {
        case 3: tet_dmma_wrap<3>(op); break;
        case 4: tet_dmma_wrap<4>(op); break;
        case 5: tet_dmma_wrap<5>(op); break;
        case 6: tet_dmma_wrap<6>(op); break;
        case 7: tet_dmma_wrap<7>(op); break;
        default: break;
    }
}

} // namespace nekmf
