// prism_dmma.cu -- BwdTrans and IProductWRTBase on prisms and pyramids at nm = 3..7 (default quadrature
// nq = (nm+1, nm+1, nm)) with FP64 tensor-core tiles (DMMA, mma.sync.m8n8k4.f64).
//
// Reference semantics: MatrixFreeOps/BwdTransKernels.hpp:128-300 (BwdTransPrismKernel, CORRECT = true for the
// modified basis) and its pyramid sibling, IProductKernels.hpp:316-450 (IProductPrismKernel) and the pyramid one;
// results differ from shape_kernels.cuh / generic_kernels.cu by summation order only.
//
// A prism is (triangle in xi_0, xi_2) x (segment in xi_1): phi_pqr = A_p(xi_0) A_q(xi_1) B_pr(xi_2).  Only the xi_2
// contraction is collapsed (its basis rows depend on p); for a fixed quadrature plane k the other two are the same
// pair of tensor contractions as on a hexahedron, and nq_0 = nq_1 = nm + 1 <= 8 fits ONE 8-row tile:
//   BwdTrans:  f_k[p][q] = sum_r c[p][q][r] B_pr(k)         in the lane that needs it as its B-operand entry (DFMA,
//                                                            the r-lines of its two (p, q) entries sit in registers)
//              C1[i][q] = sum_p A[i][p] f_k[p][q]            DMMA, tile columns ordered q = (t | 4 + t)
//              C2[j][i] = sum_q A[j][q] C1[i][q]             DMMA, the C fragment of pass 1 is the B operand as it is
//              out[k][j][i] = C2                              16-byte stores straight from the accumulator registers
//   IProduct:  the transposed chain: two chained DMMA passes per plane k (i -> p, j -> q), then the collapsed xi_2
//              contraction accumulated on the fly by the lane that owns (p, q).
// The singular-edge correction of the modified basis (mode (0, q, 1) also acts as (1, q, .) with the B_01 row) is one
// extra FMA in the lane that holds p = 1.  A PYRAMID has the same two tensor directions; its xi_2 rows depend on
// (p, q) (mode lines of length nm - max(p, q)) and its one correction is the top vertex: mode (0,0,1) also acts through
// the entries (0,1), (1,0), (1,1) with the table row 1 -- the `PYR` variant of the same kernel.
// Shared memory is touched to read the input only.  Every warp is an
// independent worker: elements (pairs where one block is an odd number of doubles) arrive by TMA bulk copies into the
// warp's own double buffer.
#include "op_internal.h"
#include <stdlib.h>
#include <string.h>

namespace nekmf
{

// rows of the xi_2 table: prism (p, r) pairs, pyramid all modes (p, q, r)
template <bool PYR, int NM> constexpr int prdm_rows() { return PYR ? NM * (NM + 1) * (2 * NM + 1) / 6 : NM * (NM + 1) / 2; }
template <bool PYR, int NM> constexpr int prdm_nmt() { return PYR ? NM * (NM + 1) * (2 * NM + 1) / 6 : NM * NM * (NM + 1) / 2; }

template <bool PYR, int NM> struct PrismDmmaTab
{
    static constexpr int NQ0 = NM + 1, NQ2 = NM, NROWS = prdm_rows<PYR, NM>();
    double b0[NM * NQ0];    // bdata of direction 0 (and 1: same basis and points), [p][i]
    double b2[NROWS * NQ2]; // bdata of direction 2
    double w0[NQ0], w2[NQ2]; // weights (collapsed-coordinate factor folded into w2)
    int start[NM * NM];     // first coefficient of the mode line (p, q)
    int row[NM * NM];       // its first row of the xi_2 table
    int len[NM * NM];       // its length: nm - p (prism), nm - max(p, q) (pyramid)
};

struct PrismDmmaArgs
{
    const double *in;
    double *out;
    const double *jac; // IProduct: [nElmt] | [nElmt][nqTot]
    int nElmt;
    int in_aligned, out_aligned; // 16-byte aligned
};

__device__ __forceinline__ void pr_mma(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// OP 0: BwdTrans, 1: IProductWRTBase
template <bool PYR, int OP, int NM, bool DEF> struct PrismDmmaCfg
{
    static constexpr int NQ0 = NM + 1, NQ1 = NM + 1, NQ2 = NM, NROWS = prdm_rows<PYR, NM>();
    static constexpr int NMT = prdm_nmt<PYR, NM>(), NQT = NQ0 * NQ1 * NQ2;
    static constexpr int IN_EL = OP == 0 ? NMT : NQT;
    static constexpr int EPB   = (IN_EL % 2) ? 2 : 1;                  // elements per buffer: a whole number of 16-byte units
    static constexpr int BUF   = EPB * IN_EL;
    static constexpr int SLOT  = BUF + ((OP == 1 && DEF) ? EPB * NQT : 0); // input block (+ its Jacobians)
    static constexpr int STG   = OP == 1 ? ((NMT + 1) & ~1) : 0;       // IProduct: output staging (coalesced stores)
    static constexpr int PER_WARP = 2 * SLOT + STG + 2;                // double buffer + staging + two mbarriers
    static constexpr int B2S   = (NROWS * NQ2 + 1) & ~1;               // shared copy of the collapsed table
    static constexpr int WARPS = (B2S + 8 * PER_WARP) * 8 + 16 <= 224 * 1024 ? 8 : 6, T = WARPS * 32;
    static constexpr size_t SMEM = (size_t)(B2S + WARPS * PER_WARP) * 8 + 16;
};

// MINB: resident CTAs per SM the register allocation aims at (2: 128 registers, 3: 80, 4: 64)
template <bool PYR, int OP, int NM, bool DEF, int MINB>
__global__ void __launch_bounds__(PrismDmmaCfg<PYR, OP, NM, DEF>::T, MINB)
    prism_dmma_kernel(const __grid_constant__ PrismDmmaTab<PYR, NM> tab, const __grid_constant__ PrismDmmaArgs args)
{
    using Cfg = PrismDmmaCfg<PYR, OP, NM, DEF>;
    constexpr int NQ0 = Cfg::NQ0, NQ1 = Cfg::NQ1, NQ2 = Cfg::NQ2, NROWS = Cfg::NROWS, NMT = Cfg::NMT, NQT = Cfg::NQT;
    constexpr int IN_EL = Cfg::IN_EL, BUF = Cfg::BUF, SLOT = Cfg::SLOT, EPB = Cfg::EPB;
    constexpr bool JSM = OP == 1 && DEF;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = lane & 3, g = lane >> 2;
    double *sB2   = reinterpret_cast<double *>(smem_raw);
    double *wbase = sB2 + Cfg::B2S + (size_t)warp * Cfg::PER_WARP;
    double *sStg  = wbase + 2 * SLOT;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase + 2 * SLOT + Cfg::STG);

    // collapsed table, with the xi_2 weights folded in for IProduct
    for (int i = threadIdx.x; i < NROWS * NQ2; i += Cfg::T) sB2[i] = OP == 0 ? tab.b2[i] : tab.b2[i] * tab.w2[i % NQ2];
    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int nPairs = (args.nElmt + EPB - 1) / EPB; // blocks of EPB elements
    const int GW = gridDim.x * Cfg::WARPS, gw = blockIdx.x * Cfg::WARPS + warp;
    auto pair_ne = [&](int pr) { return args.nElmt - EPB * pr >= EPB ? EPB : 1; };
    auto tma_ok  = [&](int pr) { return args.in_aligned && pair_ne(pr) == EPB; };
    auto issue   = [&](int pr, int slot) { // lane 0
        if (!tma_ok(pr)) return;
        mbar_expect_tx(bar + slot, (uint32_t)((BUF + (JSM ? EPB * NQT : 0)) * 8));
        tma_load_1d(wbase + slot * SLOT, args.in + (size_t)pr * BUF, (uint32_t)(BUF * 8), bar + slot);
        if (JSM) tma_load_1d(wbase + slot * SLOT + BUF, args.jac + (size_t)pr * EPB * NQT, (uint32_t)(EPB * NQT * 8), bar + slot);
    };

    // A fragments of the basis matrix A (directions 0 and 1 share it): rows g of the output index, columns 4 s + t of
    // the contracted index.  BwdTrans: A[i][p] = b0[p*NQ0 + i]; IProduct: A^T[p][i] = b0[p*NQ0 + i].
    double A[2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
    {
        const int con = 4 * s + t;
        double v = 0.0;
        if (OP == 0) { if (g < NQ0 && con < NM) v = tab.b0[con * NQ0 + g]; }
        else { if (g < NM && con < NQ0) v = tab.b0[g * NQ0 + con]; }
        A[s] = v;
    }
    // tile column g of pass 1 <-> second index (q resp. j) = g/2 for even g, 4 + g/2 for odd g
    const int col0 = (g & 1) ? 4 + (g >> 1) : (g >> 1);
    // mode line (p, q): first coefficient, first row of the xi_2 table, length.  Prisms: closed forms (the lane-dependent
    // lookups in the kernel-parameter tables that the pyramid needs cost the prism kernels 0.19 -> 0.28 ms at nm = 7 when
    // both shapes went through them); pyramids: tables.
    auto mpr     = [](int p) { return p * NM - p * (p - 1) / 2; }; // first row of the (p, .) block of the prism table
    auto l_start = [&](int p, int q) { return PYR ? tab.start[p * NM + q] : NM * mpr(p) + q * (NM - p); };
    auto l_len   = [&](int p, int q) { return PYR ? tab.len[p * NM + q] : NM - p; };
    auto l_row   = [&](int p, int q) { return PYR ? tab.row[p * NM + q] : mpr(p); };

    uint32_t phase[2] = {0, 0};
    int slot          = 0;
    if (lane == 0 && gw < nPairs) issue(gw, 0);

    for (int pr = gw; pr < nPairs; pr += GW, slot ^= 1)
    {
        const int ne = pair_ne(pr);
        double *sIn  = wbase + slot * SLOT;
        if (lane == 0 && pr + GW < nPairs) issue(pr + GW, slot ^ 1); // the other buffer was consumed one trip ago
        // regular IProduct: the element's Jacobian is requested before the wait, its latency hides behind the copy
        double jpre = 1.0;
        if (OP == 1 && !DEF) jpre = __ldg(args.jac + (size_t)pr * EPB);
        if (tma_ok(pr))
        {
            mbar_wait(bar + slot, phase[slot]);
            phase[slot] ^= 1;
        }
        else
        {
            // odd tail element or 8-byte aligned caller arrays: plain loads by the warp
            const double *src = args.in + (size_t)pr * BUF;
            for (int i = lane; i < ne * IN_EL; i += 32) sIn[i] = __ldg(src + i);
            if (JSM)
            {
                const double *sj = args.jac + (size_t)pr * EPB * NQT;
                for (int i = lane; i < ne * NQT; i += 32) sIn[BUF + i] = __ldg(sj + i);
            }
        }
        __syncwarp();

#pragma unroll 1
        for (int e = 0; e < ne; ++e)
        {
            const size_t el = (size_t)pr * EPB + e;
            const double *U = sIn + e * IN_EL;
            if (OP == 0)
            {
                // ------------------------------------------------------------ BwdTrans: c[p][q][r] -> out[k][j][i]
                // the lane's B-operand entries of pass 1: (p, q) = (t, col0) and (4 + t, col0); their r-lines
                const int q  = col0;
                const int p0 = t, p1 = 4 + t;
                const bool v0 = q < NM && p0 < NM, v1 = q < NM && p1 < NM;
                double c0[NM], c1[NM > 4 ? NM - 4 : 1];
                const int base0 = v0 ? l_start(p0, q) : 0, len0 = v0 ? l_len(p0, q) : 0;
                const int base1 = v1 ? l_start(p1, q) : 0, len1 = v1 ? l_len(p1, q) : 0;
#pragma unroll
                for (int r = 0; r < NM; ++r)
                {
                    const bool ok = r < len0;
                    const double x = U[ok ? base0 + r : 0];
                    c0[r] = ok ? x : 0.0;
                }
#pragma unroll
                for (int r = 0; r < NM - 4; ++r)
                {
                    const bool ok = r < len1;
                    const double x = U[ok ? base1 + r : 0];
                    c1[r] = ok ? x : 0.0;
                }
                // prism, singular edge: mode (0, q, 1) also contributes to f[1][q] through the row (0, 1) of the table;
                // pyramid, top vertex: mode (0, 0, 1) contributes to f[0][1], f[1][0], f[1][1] through the row 1
                const bool cor  = v0 && (PYR ? ((p0 == 0 && q == 1) || (p0 == 1 && q <= 1)) : p0 == 1);
                const double xc = U[cor ? (PYR ? 1 : q * NM + 1) : 0];
                const double cc = cor ? xc : 0.0;
                const int row0 = v0 ? l_row(p0, q) : 0, row1 = v1 ? l_row(p1, q) : 0;
                const int i0 = 2 * t;
                double *o = args.out + el * NQT + g * NQ0 + i0;
#pragma unroll
                for (int k = 0; k < NQ2; ++k)
                {
                    double f0 = cc * sB2[NQ2 + k], f1 = 0.0;
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                    {
                        const int row = row0 + r < NROWS ? row0 + r : NROWS - 1; // clamped: c0[r] is zero there
                        f0 = fma(c0[r], sB2[row * NQ2 + k], f0);
                    }
#pragma unroll
                    for (int r = 0; r < NM - 4; ++r)
                    {
                        const int row = row1 + r < NROWS ? row1 + r : NROWS - 1;
                        f1 = fma(c1[r], sB2[row * NQ2 + k], f1);
                    }
                    double x0 = 0.0, x1 = 0.0, d0 = 0.0, d1 = 0.0;
                    pr_mma(x0, x1, A[0], f0); // C1[i = g][q = t | 4 + t]
                    pr_mma(x0, x1, A[1], f1);
                    pr_mma(d0, d1, A[0], x0); // C2[j = g][i = 2t, 2t + 1]
                    pr_mma(d0, d1, A[1], x1);
                    if (g < NQ1 && i0 < NQ0)
                    {
                        if ((NQ0 % 2 == 0) && args.out_aligned)
                            *reinterpret_cast<double2 *>(o + k * (NQ0 * NQ1)) = make_double2(d0, d1);
                        else
                        {
                            o[k * (NQ0 * NQ1)] = d0;
                            if (i0 + 1 < NQ0) o[k * (NQ0 * NQ1) + 1] = d1;
                        }
                    }
                }
            }
            else
            {
                // ------------------------------------------------------------ IProductWRTBase: F[k][j][i] -> out[p][q][r]
                const double *Jp  = sIn + BUF + e * NQT; // DEF only
                const double jreg = DEF ? 1.0 : (e == 0 ? jpre : __ldg(args.jac + el));
                // the lane's B-operand entries of pass 1: (j, i) = (col0, t) and (col0, 4 + t)
                const int j = col0;
                const bool v0 = j < NQ1 && t < NQ0, v1 = j < NQ1 && 4 + t < NQ0;
                const int ix0 = v0 ? j * NQ0 + t : 0, ix1 = v1 ? j * NQ0 + 4 + t : 0;
                const double wj  = j < NQ1 ? tab.w0[j] * jreg : 0.0; // directions 0 and 1 share the weights
                const double w00 = v0 ? tab.w0[t] * wj : 0.0, w01 = v1 ? tab.w0[(4 + t) < NQ0 ? 4 + t : 0] * wj : 0.0;
                // the lane owns (p, q) = (2t, g) and (2t + 1, g) of the result
                const int p0 = 2 * t, p1 = 2 * t + 1;
                const bool o0 = g < NM && p0 < NM, o1 = g < NM && p1 < NM;
                const int row0 = o0 ? l_row(p0, g) : 0, row1 = o1 ? l_row(p1, g) : 0;
                // corrections: prism, singular edge: the (1, q) value also feeds mode (0, q, 1) through the row (0, 1);
                // pyramid, top vertex: the (0,1), (1,0), (1,1) values feed mode (0,0,1) through the row 1 (lane (g,t) = (1,0)
                // adds its share to the staged result afterwards)
                const bool own1 = PYR ? (t == 0 && g == 0) : t == 0;
                const bool ext1 = PYR && t == 0 && g == 1;
                double ext = 0.0;
                double acc0[NM], acc1[NM > 1 ? NM - 1 : 1];
#pragma unroll
                for (int r = 0; r < NM; ++r) acc0[r] = 0.0;
#pragma unroll
                for (int r = 0; r < NM - 1; ++r) acc1[r] = 0.0;
#pragma unroll
                for (int k = 0; k < NQ2; ++k)
                {
                    double f0 = U[k * (NQ0 * NQ1) + ix0] * w00, f1 = U[k * (NQ0 * NQ1) + ix1] * w01;
                    if (DEF)
                    {
                        f0 *= Jp[k * (NQ0 * NQ1) + ix0];
                        f1 *= Jp[k * (NQ0 * NQ1) + ix1];
                    }
                    double x0 = 0.0, x1 = 0.0, d0 = 0.0, d1 = 0.0;
                    pr_mma(x0, x1, A[0], f0); // C1[p = g][j = t | 4 + t]
                    pr_mma(x0, x1, A[1], f1);
                    pr_mma(d0, d1, A[0], x0); // C2[q = g][p = 2t, 2t + 1]
                    pr_mma(d0, d1, A[1], x1);
                    // collapsed xi_2 contraction on the fly (weights already in the table)
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                    {
                        const int row = row0 + r < NROWS ? row0 + r : NROWS - 1;
                        acc0[r]       = fma(sB2[row * NQ2 + k], d0, acc0[r]);
                    }
#pragma unroll
                    for (int r = 0; r < NM - 1; ++r)
                    {
                        const int row = row1 + r < NROWS ? row1 + r : NROWS - 1;
                        acc1[r]       = fma(sB2[row * NQ2 + k], d1, acc1[r]);
                    }
                    if (NM > 1) acc0[1] = fma(own1 ? sB2[NQ2 + k] : 0.0, d1, acc0[1]);
                    if (PYR) ext = fma(ext1 ? sB2[NQ2 + k] : 0.0, d0 + d1, ext);
                }
                // the (p, q) owners write their mode lines: staged in shared memory, stored coalesced
                if (o0)
                {
                    double *o    = sStg + l_start(p0, g);
                    const int ln = l_len(p0, g);
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                        if (r < ln) o[r] = acc0[r];
                }
                if (o1)
                {
                    double *o    = sStg + l_start(p1, g);
                    const int ln = l_len(p1, g);
#pragma unroll
                    for (int r = 0; r < NM - 1; ++r)
                        if (r < ln) o[r] = acc1[r];
                }
                __syncwarp();
                if (PYR)
                {
                    if (ext1) sStg[1] += ext;
                    __syncwarp();
                }
                {
                    double *o = args.out + el * NMT;
                    if ((NMT % 2 == 0) && args.out_aligned)
                        for (int i2 = lane; i2 < NMT / 2; i2 += 32)
                            *reinterpret_cast<double2 *>(o + 2 * i2) = *reinterpret_cast<const double2 *>(sStg + 2 * i2);
                    else
                        for (int i = lane; i < NMT; i += 32) o[i] = sStg[i];
                }
                __syncwarp(); // staging block free for the next element
            }
        }
        __syncwarp(); // every lane is done with this buffer before lane 0 refills it
    }
}

template <bool PYR, int NM> struct PrismDmmaState
{
    PrismDmmaTab<PYR, NM> tab;
    void *fallback_state          = nullptr;
    void (*fallback_free)(void *) = nullptr;
    int bps[2][2]                 = {{0, 0}, {0, 0}};
    int minb                      = 2;
};

template <bool PYR, int OP, int NM, bool DEF> static int prism_dmma_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    auto *st  = static_cast<PrismDmmaState<PYR, NM> *>(op->kstate);
    using Cfg = PrismDmmaCfg<PYR, OP, NM, DEF>;
    auto kern = st->minb == 3 ? prism_dmma_kernel<PYR, OP, NM, DEF, 3> : (st->minb == 4 ? prism_dmma_kernel<PYR, OP, NM, DEF, 4> : prism_dmma_kernel<PYR, OP, NM, DEF, 2>);
    int &bps  = st->bps[OP][DEF ? 1 : 0];
    if (bps == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("prism / pyramid DMMA kernel <%d> does not fit on an SM", NM); return NEKMF_ERR_CUDA; }
        bps = nb;
    }
    PrismDmmaArgs a;
    a.in = in[0]; a.out = out[0]; a.nElmt = op->run_ne;
    a.jac = nullptr;
    if (OP == 1) a.jac = DEF ? op->d_jac + (size_t)op->run_e0 * op->geo_pitch : op->d_jac + op->run_e0;
    a.in_aligned  = ((((uintptr_t)in[0]) | (OP == 1 && DEF ? (uintptr_t)a.jac : 0)) & 15) == 0;
    a.out_aligned = (((uintptr_t)out[0]) & 15) == 0;
    const int nPairs = (op->run_ne + Cfg::EPB - 1) / Cfg::EPB;
    int grid         = bps * NUM_SMS;
    const int need   = (nPairs + Cfg::WARPS - 1) / Cfg::WARPS;
    if (grid > need) grid = need;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(st->tab, a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <bool PYR, int NM> static bool prism_dmma_wrap(nekmf_op_s *op)
{
    using Tab = PrismDmmaTab<PYR, NM>;
    if (op->b[0].size() != sizeof(Tab::b0) / 8 || op->b[2].size() != sizeof(Tab::b2) / 8 || op->ws[0].size() != sizeof(Tab::w0) / 8 ||
        op->ws[2].size() != sizeof(Tab::w2) / 8 || op->nmTot != prdm_nmt<PYR, NM>())
        return false; // tables of another size than the kernel is compiled for: keep the installed kernel
    auto *st  = new PrismDmmaState<PYR, NM>;
    memcpy(st->tab.b0, op->b[0].data(), sizeof(st->tab.b0));
    memcpy(st->tab.b2, op->b[2].data(), sizeof(st->tab.b2));
    memcpy(st->tab.w0, op->ws[0].data(), sizeof(st->tab.w0));
    memcpy(st->tab.w2, op->ws[2].data(), sizeof(st->tab.w2));
    {
        // mode lines in the reference's order: p outer, q, r inner
        int mode = 0;
        for (int p = 0; p < NM; ++p)
            for (int q = 0; q < NM; ++q)
            {
                const int ln          = PYR ? NM - (p > q ? p : q) : NM - p;
                st->tab.start[p * NM + q] = mode;
                st->tab.len[p * NM + q]   = ln;
                st->tab.row[p * NM + q]   = PYR ? mode : p * NM - p * (p - 1) / 2;
                mode += ln;
            }
    }
    if (const char *vb = getenv("NEKMF_PRISM_DMMA_MINB")) // A/B knob: register budget (resident CTAs per SM aimed at)
        if (vb[0] == '3' || vb[0] == '4') st->minb = vb[0] - '0'; // measured: 2 is the fastest from nm = 5 on
    st->fallback_state = op->kstate;
    st->fallback_free  = op->kstate_free;
    op->kstate         = st;
    op->kstate_free    = [](void *p) {
        auto *s = static_cast<PrismDmmaState<PYR, NM> *>(p);
        if (s->fallback_state && s->fallback_free) s->fallback_free(s->fallback_state);
        delete s;
    };
    char name[128];
    if (op->optype == NEKMF_BWDTRANS)
    {
        op->launch = prism_dmma_launch<PYR, 0, NM, false>;
        snprintf(name, sizeof(name), "%s_dmma_kernel<bwd,nm=%d>(DMMA m8n8k4, collapsed direction in the owning lane)", PYR ? "pyr" : "prism", NM);
    }
    else
    {
        op->launch = op->deformed ? prism_dmma_launch<PYR, 1, NM, true> : prism_dmma_launch<PYR, 1, NM, false>;
        snprintf(name, sizeof(name), "%s_dmma_kernel<iprod,nm=%d,%s>(DMMA m8n8k4)", PYR ? "pyr" : "prism", NM, op->deformed ? "deformed" : "regular");
    }
    op->kname = name;
    return true;
}

// called from select_shape_fast after the pencil launcher is installed: default quadrature, BwdTrans or
// IProductWRTBase on prisms.  NEKMF_PRISM_DMMA=0 keeps the pencil kernels (the other arm of the A/B), =all takes the
// tensor-core kernels at every instantiated order, the default is the set of orders where they measured faster.
void prism_dmma_maybe_wrap(nekmf_op_s *op)
{
    if (op->shape != NEKMF_PRISM) return;
    if (op->optype != NEKMF_BWDTRANS && op->optype != NEKMF_IPRODUCTWRTBASE) return;
    const int nm = op->nm[0];
    if (op->nm[1] != nm || op->nm[2] != nm || op->nq[0] != nm + 1 || op->nq[1] != nm + 1 || op->nq[2] != nm) return;
    // directions 0 and 1 must share basis and weights (the default GLL / GLL / Gauss-Radau quadrature does)
    if (op->b[1] != op->b[0] || op->ws[1] != op->ws[0]) return;
    if (op->deformed && op->geo_pitch != op->nqTot) return;
    const char *v = getenv("NEKMF_PRISM_DMMA");
    if (v && v[0] == '0') return;
    // measured A/B (profiles/r02_sweep_prism_dmma_*.jsonl): faster from nm = 5 on (three-quarters of the tile rows in
    // use); NEKMF_PRISM_DMMA=all takes it at every instantiated order
    if (!(v && v[0] == 'a') && nm < 5) return;
    switch (nm)
    {
        case 3: prism_dmma_wrap<false, 3>(op); break;
        case 4: prism_dmma_wrap<false, 4>(op); break;
        case 5: prism_dmma_wrap<false, 5>(op); break;
        case 6: prism_dmma_wrap<false, 6>(op); break;
        case 7: prism_dmma_wrap<false, 7>(op); break;
        default: break;
    }
}

// the same for pyramids (called from nekmf_op_create after the runtime-sized kernel is installed: there is no
// compile-time sized pencil family for pyramids, so the tensor-core kernels are taken at every instantiated order).
// NEKMF_PYR_DMMA=0 keeps the runtime-sized kernel.
void pyr_dmma_maybe_wrap(nekmf_op_s *op)
{
    if (op->shape != NEKMF_PYR) return;
    if (op->optype != NEKMF_BWDTRANS && op->optype != NEKMF_IPRODUCTWRTBASE) return;
    const int nm = op->nm[0];
    if (op->nm[1] != nm || op->nm[2] != nm || op->nq[0] != nm + 1 || op->nq[1] != nm + 1 || op->nq[2] != nm) return;
    if (op->b[1] != op->b[0] || op->ws[1] != op->ws[0]) return;
    if (op->deformed && op->geo_pitch != op->nqTot) return;
    const char *v = getenv("NEKMF_PYR_DMMA");
    if (v && v[0] == '0') return;
    // default = the cells where the tensor tiles beat the compile-time sized pencil kernels that pyramids got later in round 2
    // (profiles/r02_final_sweep_collapsed.jsonl against r02_sweep_pyr_shape_1.jsonl): nm = 7 (0.36 / 0.55 / 0.67 against 0.54 /
    // 0.59 / 0.79 ms), BwdTrans and deformed IProductWRTBase at nm = 5; the pencil kernels win at nm = 3, 4, 6 (0.37-0.45
    // against 0.46-0.91 ms).  NEKMF_PYR_DMMA=1 takes the tiles at every instantiated order.
    const bool forced = v && v[0] == '1';
    const bool faster = nm == 7 || (nm == 5 && (op->optype == NEKMF_BWDTRANS || op->deformed));
    if (!forced && !faster) return;
    switch (nm)
    {
        case 3: prism_dmma_wrap<true, 3>(op); break;
        case 4: prism_dmma_wrap<true, 4>(op); break;
        case 5: prism_dmma_wrap<true, 5>(op); break;
        case 6: prism_dmma_wrap<true, 6>(op); break;
        case 7: prism_dmma_wrap<true, 7>(op); break;
        default: break;
    }
}

} // namespace nekmf
