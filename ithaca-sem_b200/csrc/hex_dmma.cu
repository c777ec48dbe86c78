// hex_dmma.cu -- BwdTrans and IProductWRTBase on hexahedra at nm = 7..11 (default quadrature nq = nm + 1) with FP64
// tensor-core tiles (DMMA, mma.sync.m8n8k4.f64): the DMMA arm of the DMMA-vs-DFMA comparison of BASELINE.json
// configs[2].
//
// Reference semantics: MatrixFreeOps/BwdTransKernels.hpp:302-372, IProductKernels.hpp:236-314 (three 1-D
// contractions with the basis matrix B[p][i]); results differ from the DFMA kernels by summation order only.
//
// On B200 the DMMA pipe has the DFMA pipe's peak (37.1 vs 33.8 TFLOP/s measured, tools/fp64_peak.cu), so tensor
// tiles cannot raise the ceiling; what they can do is remove the shared-memory traffic that bounds the pencil
// kernels (hex_kernels.cuh: every pass reads and writes the whole intermediate), by CHAINING two contractions in
// registers:
//   pass 1 (contract p):  C1[i][q] = sum_p B[i][p] U[p][q]      A = basis matrix (rows i), B operand = the
//                          coefficient block read from shared memory, tile columns n <-> q
//   pass 2 (contract q):  C2[j][i] = sum_q B[j][q] C1[i][q]     B operand = the C fragment of pass 1 AS IT IS
// The m8n8k4 C fragment gives lane (t = lane % 4, g = lane / 4) row g, columns 2t and 2t+1; the B fragment of
// k-step s wants from that lane row k = t, column g.  With the columns of a pass-1 tile ordered q(2t) = t,
// q(2t+1) = 4 + t (+ 8 per column tile) the lane already holds C1[i = g][q = t] and C1[i = g][q = 4 + t] -- exactly
// the B-operand entries of two k-steps of pass 2, whose A fragments take the matching columns of the basis matrix:
// the SAME registers as in pass 1 (k-step s = 2 ct + s').  No shuffle, no shared memory between the passes.  The
// third contraction runs over the index the tiles are enumerated by (r resp. k), whose full line sits in ONE
// lane's registers: plain DFMA, accumulated on the fly, and the results go to global memory straight from
// registers.  Shared memory is touched only to read the input.
// nq = 8 (nm = 7) fits exactly: M = 8 rows, K = 7 -> 8.  At nq = 9..12 a second row tile is 13-50 % full and the
// pipe does up to twice the useful work; whether that still beats the pencil kernels is what the A/B measures
// (DESIGN.md 4.2c).
//
// Mapping: every warp is an independent worker (as hex_kron.cu): element pairs (16-byte aligned blocks) arrive by
// TMA bulk copies into the warp's own buffer(s), completion on the warp's own mbarrier(s).
#include "op_internal.h"
#include <stdlib.h>
#include <string.h>

namespace nekmf
{

template <int NM> struct HexDmmaTab
{
    static constexpr int NQ = NM + 1;
    double B[NM * NQ]; // bdata[m*NQ + i]
    double w[NQ];      // quadrature weights
};

struct HexDmmaArgs
{
    const double *in;
    double *out;
    const double *jac; // IProduct: [nElmt] | [nElmt][pitch]
    int nElmt;
    int jpitch;                  // deformed IProduct: element pitch of jac
    int in_aligned, out_aligned; // 16-byte aligned
};

__device__ __forceinline__ void dm_mma(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// OP 0: BwdTrans (NM^3 -> NQ^3 per element), 1: IProductWRTBase (NQ^3 -> NM^3)
template <int OP, int NM, bool DEF> struct HexDmmaCfg
{
    static constexpr int NQ = NM + 1, NM3 = NM * NM * NM, NQ3 = NQ * NQ * NQ;
    static constexpr int NROW = OP == 0 ? NQ : NM; // rows of the basis matrix as the A operand (output index)
    static constexpr int NCON = OP == 0 ? NM : NQ; // contracted index
    static constexpr int MT = (NROW + 7) / 8, KS = (NCON + 3) / 4, CT = (NCON + 7) / 8;
    static constexpr int IN_EL = OP == 0 ? NM3 : NQ3;
    static constexpr int BUF   = 2 * IN_EL;                               // one element PAIR (even number of doubles)
    static constexpr int JP    = (NQ3 + 1) & ~1;                          // element pitch of the deformed Jacobian
    static constexpr int SLOT  = BUF + ((OP == 1 && DEF) ? 2 * JP : 0);   // input pair (+ its Jacobians)
    static constexpr int NBUF  = ((2 * SLOT + 2) * 8 * 6 <= 224 * 1024) ? 2 : 1; // double buffer when 6 warps' worth fits
    static constexpr int PER_WARP = NBUF * SLOT + 2;                      // + two mbarriers
    static constexpr int W_FIT = (224 * 1024) / (PER_WARP * 8);
    static constexpr int WARPS = W_FIT >= 12 ? 12 : (W_FIT >= 8 ? 8 : (W_FIT >= 6 ? 6 : 4));
    static constexpr int T     = WARPS * 32;
    static constexpr size_t SMEM = (size_t)WARPS * PER_WARP * 8 + 16;
};

template <int OP, int NM, bool DEF>
__global__ void __launch_bounds__(HexDmmaCfg<OP, NM, DEF>::T, 1)
    hex_dmma_kernel(const __grid_constant__ HexDmmaTab<NM> tab, const __grid_constant__ HexDmmaArgs args)
{
    using Cfg = HexDmmaCfg<OP, NM, DEF>;
    constexpr int NQ = Cfg::NQ, NM3 = Cfg::NM3, NQ3 = Cfg::NQ3, IN_EL = Cfg::IN_EL, BUF = Cfg::BUF, NBUF = Cfg::NBUF;
    constexpr int SLOT = Cfg::SLOT, JP = Cfg::JP;
    constexpr bool JSM = OP == 1 && DEF; // the deformed Jacobian travels with the input (TMA into shared memory)
    constexpr int MT = Cfg::MT, KS = Cfg::KS, CT = Cfg::CT, NROW = Cfg::NROW, NCON = Cfg::NCON;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = lane & 3, g = lane >> 2;
    double *wbase = reinterpret_cast<double *>(smem_raw) + (size_t)warp * Cfg::PER_WARP;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase + NBUF * SLOT);

    const int nPairs = (args.nElmt + 1) / 2;
    const int GW = gridDim.x * Cfg::WARPS, gw = blockIdx.x * Cfg::WARPS + warp;

    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto pair_ne = [&](int pr) { return args.nElmt - 2 * pr >= 2 ? 2 : 1; };
    auto tma_ok  = [&](int pr) { return args.in_aligned && pair_ne(pr) == 2 && (!JSM || args.jpitch == JP); };
    auto issue   = [&](int pr, int slot) { // lane 0
        if (!tma_ok(pr)) return;
        mbar_expect_tx(bar + slot, (uint32_t)((2 * IN_EL + (JSM ? 2 * JP : 0)) * 8));
        tma_load_1d(wbase + slot * SLOT, args.in + (size_t)pr * 2 * IN_EL, (uint32_t)(2 * IN_EL * 8), bar + slot);
        if (JSM) tma_load_1d(wbase + slot * SLOT + BUF, args.jac + (size_t)pr * 2 * JP, (uint32_t)(2 * JP * 8), bar + slot);
    };

    // A fragments of the basis matrix, rows g + 8 mt of the output index, columns 4 s + t of the contracted index:
    // BwdTrans B[i][p] = bdata[p*NQ + i], IProduct B^T[p][i] = bdata[p*NQ + i].  Shared by pass 1 and pass 2.
    double A[MT][KS];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int s = 0; s < KS; ++s)
        {
            const int row = g + 8 * mt, con = 4 * s + t;
            double v = 0.0;
            if (row < NROW && con < NCON) v = OP == 0 ? tab.B[con * NQ + row] : tab.B[row * NQ + con];
            A[mt][s] = v;
        }
    // tile column g of pass 1 <-> second index (q resp. j) = 8 ct + (g/2 for even g, 4 + g/2 for odd g)
    const int col0 = (g & 1) ? 4 + (g >> 1) : (g >> 1);

    uint32_t phase[2] = {0, 0};
    int slot          = 0;
    if (lane == 0 && gw < nPairs) issue(gw, 0);

    for (int pr = gw; pr < nPairs; pr += GW)
    {
        const int ne = pair_ne(pr);
        double *sIn  = wbase + slot * SLOT;
        if (NBUF == 2 && lane == 0 && pr + GW < nPairs) issue(pr + GW, slot ^ 1); // consumed one trip ago
        if (tma_ok(pr))
        {
            mbar_wait(bar + slot, phase[slot]);
            phase[slot] ^= 1;
        }
        else
        {
            // odd tail element or 8-byte aligned caller arrays: plain loads by the warp
            const double *src = args.in + (size_t)pr * 2 * IN_EL;
            for (int i = lane; i < ne * IN_EL; i += 32) sIn[i] = __ldg(src + i);
            if (JSM)
                for (int e = 0; e < ne; ++e)
                {
                    const double *sj = args.jac + ((size_t)pr * 2 + e) * args.jpitch;
                    for (int i = lane; i < NQ3; i += 32) sIn[BUF + e * JP + i] = __ldg(sj + i);
                }
        }
        __syncwarp();

#pragma unroll 1
        for (int e = 0; e < ne; ++e)
        {
            const size_t el = (size_t)pr * 2 + e;
            const double *U = sIn + e * IN_EL;
            if (OP == 0)
            {
                // ------------------------------------------------------------ BwdTrans: U[r][q][p] -> out[k][j][i]
#pragma unroll
                for (int it = 0; it < MT; ++it)
                {
                    double acc[NQ][MT][2];
#pragma unroll
                    for (int k = 0; k < NQ; ++k)
#pragma unroll
                        for (int jt = 0; jt < MT; ++jt) acc[k][jt][0] = acc[k][jt][1] = 0.0;
#pragma unroll 1
                    for (int r = 0; r < NM; ++r)
                    {
                        double c[CT][2];
#pragma unroll
                        for (int ct = 0; ct < CT; ++ct)
                        {
                            const int q = 8 * ct + col0;
                            c[ct][0] = c[ct][1] = 0.0;
#pragma unroll
                            for (int s = 0; s < KS; ++s)
                            {
                                // branch-free operand fetch (clamped index + select): mma.sync needs the warp converged
                                const int p    = 4 * s + t;
                                const bool ok  = q < NM && p < NM;
                                const double v = U[ok ? r * (NM * NM) + q * NM + p : 0];
                                const double b = ok ? v : 0.0;
                                dm_mma(c[ct][0], c[ct][1], A[it][s], b); // C1[i = g + 8 it][q = 8 ct + t | 4 + t]
                            }
                        }
#pragma unroll
                        for (int jt = 0; jt < MT; ++jt)
                        {
                            double d0 = 0.0, d1 = 0.0;
#pragma unroll
                            for (int ct = 0; ct < CT; ++ct)
#pragma unroll
                                for (int sp = 0; sp < 2; ++sp)
                                    if (2 * ct + sp < KS) dm_mma(d0, d1, A[jt][2 * ct + sp], c[ct][sp]);
                            // C2[j = g + 8 jt][i = 8 it + 2t, + 1]: third contraction on the fly
#pragma unroll
                            for (int k = 0; k < NQ; ++k)
                            {
                                const double bk = tab.B[r * NQ + k];
                                acc[k][jt][0]   = fma(bk, d0, acc[k][jt][0]);
                                acc[k][jt][1]   = fma(bk, d1, acc[k][jt][1]);
                            }
                        }
                    }
                    const int i0 = 8 * it + 2 * t;
#pragma unroll
                    for (int jt = 0; jt < MT; ++jt)
                    {
                        const int j = g + 8 * jt;
                        if (j < NQ && i0 < NQ)
                        {
                            double *o = args.out + el * NQ3 + j * NQ + i0;
#pragma unroll
                            for (int k = 0; k < NQ; ++k)
                            {
                                if ((NQ % 2 == 0) && args.out_aligned)
                                    *reinterpret_cast<double2 *>(o + k * (NQ * NQ)) = make_double2(acc[k][jt][0], acc[k][jt][1]);
                                else
                                {
                                    o[k * (NQ * NQ)] = acc[k][jt][0];
                                    if (i0 + 1 < NQ) o[k * (NQ * NQ) + 1] = acc[k][jt][1];
                                }
                            }
                        }
                    }
                }
            }
            else
            {
                // ------------------------------------------------------------ IProductWRTBase: F[k][j][i] -> out[r][q][p]
                const double *Jp  = sIn + BUF + e * JP; // DEF only
                const double jreg = DEF ? 0.0 : __ldg(args.jac + el);
                double wij[CT][KS];
#pragma unroll
                for (int ct = 0; ct < CT; ++ct)
#pragma unroll
                    for (int s = 0; s < KS; ++s)
                    {
                        const int j = 8 * ct + col0, i = 4 * s + t;
                        wij[ct][s]  = (j < NQ && i < NQ) ? tab.w[j] * tab.w[i] : 0.0;
                    }
#pragma unroll
                for (int pt = 0; pt < MT; ++pt)
                {
                    double acc[NM][MT][2];
#pragma unroll
                    for (int r = 0; r < NM; ++r)
#pragma unroll
                        for (int qt = 0; qt < MT; ++qt) acc[r][qt][0] = acc[r][qt][1] = 0.0;
#pragma unroll 1
                    for (int k = 0; k < NQ; ++k)
                    {
                        const double sk = DEF ? tab.w[k] : jreg * tab.w[k];
                        double c[CT][2];
#pragma unroll
                        for (int ct = 0; ct < CT; ++ct)
                        {
                            const int j = 8 * ct + col0;
                            c[ct][0] = c[ct][1] = 0.0;
#pragma unroll
                            for (int s = 0; s < KS; ++s)
                            {
                                // branch-free operand fetch (clamped index + select): a lane-divergent branch with a
                                // global load in front of mma.sync hung the warp (nm >= 8, deformed)
                                const int i   = 4 * s + t;
                                const bool ok = j < NQ && i < NQ;
                                const int idx = ok ? k * (NQ * NQ) + j * NQ + i : 0;
                                double v      = U[idx] * (wij[ct][s] * sk);
                                if (DEF) v *= Jp[idx];
                                const double b = ok ? v : 0.0;
                                dm_mma(c[ct][0], c[ct][1], A[pt][s], b); // C1[p = g + 8 pt][j = 8 ct + t | 4 + t]
                            }
                        }
#pragma unroll
                        for (int qt = 0; qt < MT; ++qt)
                        {
                            double d0 = 0.0, d1 = 0.0;
#pragma unroll
                            for (int ct = 0; ct < CT; ++ct)
#pragma unroll
                                for (int sp = 0; sp < 2; ++sp)
                                    if (2 * ct + sp < KS) dm_mma(d0, d1, A[qt][2 * ct + sp], c[ct][sp]);
                            // C2[q = g + 8 qt][p = 8 pt + 2t, + 1]
#pragma unroll
                            for (int r = 0; r < NM; ++r)
                            {
                                const double bk = tab.B[r * NQ + k];
                                acc[r][qt][0]   = fma(bk, d0, acc[r][qt][0]);
                                acc[r][qt][1]   = fma(bk, d1, acc[r][qt][1]);
                            }
                        }
                    }
                    const int p0 = 8 * pt + 2 * t;
#pragma unroll
                    for (int qt = 0; qt < MT; ++qt)
                    {
                        const int q = g + 8 * qt;
                        if (q < NM && p0 < NM)
                        {
                            double *o = args.out + el * NM3 + q * NM + p0;
#pragma unroll
                            for (int r = 0; r < NM; ++r)
                            {
                                o[r * (NM * NM)] = acc[r][qt][0];
                                if (p0 + 1 < NM) o[r * (NM * NM) + 1] = acc[r][qt][1];
                            }
                        }
                    }
                }
            }
        }
        __syncwarp(); // every lane is done with this buffer before lane 0 refills it
        if (NBUF == 2) slot ^= 1;
        else if (lane == 0 && pr + GW < nPairs) issue(pr + GW, 0);
    }
}

template <int NM> struct HexDmmaState
{
    HexDmmaTab<NM> tab;
    void *fallback_state          = nullptr;
    void (*fallback_free)(void *) = nullptr;
    int bps[2][2]                 = {{0, 0}, {0, 0}};
};

template <int OP, int NM, bool DEF> static int hex_dmma_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    auto *st  = static_cast<HexDmmaState<NM> *>(op->kstate);
    using Cfg = HexDmmaCfg<OP, NM, DEF>;
    auto kern = hex_dmma_kernel<OP, NM, DEF>;
    int &bps  = st->bps[OP][DEF ? 1 : 0];
    if (bps == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("hex DMMA kernel <%d> does not fit on an SM", NM); return NEKMF_ERR_CUDA; }
        bps = nb;
    }
    HexDmmaArgs a;
    a.in = in[0]; a.out = out[0]; a.nElmt = op->run_ne;
    a.jac = nullptr; a.jpitch = op->geo_pitch;
    if (OP == 1) a.jac = DEF ? op->d_jac + (size_t)op->run_e0 * op->geo_pitch : op->d_jac + op->run_e0;
    a.in_aligned  = ((((uintptr_t)in[0]) | (OP == 1 && DEF ? (uintptr_t)a.jac : 0)) & 15) == 0;
    a.out_aligned = (((uintptr_t)out[0]) & 15) == 0;
    const int nPairs = (op->run_ne + 1) / 2;
    int grid         = bps * NUM_SMS;
    const int need   = (nPairs + Cfg::WARPS - 1) / Cfg::WARPS;
    if (grid > need) grid = need;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(st->tab, a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <int NM> static void hex_dmma_wrap(nekmf_op_s *op)
{
    auto *st = new HexDmmaState<NM>;
    memcpy(st->tab.B, op->b[0].data(), sizeof(st->tab.B));
    memcpy(st->tab.w, op->ws[0].data(), sizeof(st->tab.w));
    st->fallback_state = op->kstate;
    st->fallback_free  = op->kstate_free;
    op->kstate         = st;
    op->kstate_free    = [](void *p) {
        auto *s = static_cast<HexDmmaState<NM> *>(p);
        if (s->fallback_state && s->fallback_free) s->fallback_free(s->fallback_state);
        delete s;
    };
    char name[128];
    if (op->optype == NEKMF_BWDTRANS)
    {
        op->launch = hex_dmma_launch<0, NM, false>;
        snprintf(name, sizeof(name), "hex_dmma_kernel<bwd,nm=%d,nq=%d>(DMMA m8n8k4, two contractions chained in registers)", NM, NM + 1);
    }
    else
    {
        op->launch = op->deformed ? hex_dmma_launch<1, NM, true> : hex_dmma_launch<1, NM, false>;
        snprintf(name, sizeof(name), "hex_dmma_kernel<iprod,nm=%d,nq=%d,%s>(DMMA m8n8k4)", NM, NM + 1, op->deformed ? "deformed" : "regular");
    }
    op->kname = name;
}

// ------------------------------------------------------------------------------------------------ PhysDeriv, nq = 8
// PhysDeriv on REGULAR hexahedra at nm = 7 (MatrixFreeOps/PhysDerivKernels.hpp:219-372): out_c = sum_d df[3c+d] du/dxi_d.
// Lane (g, t) reads the element's values in the C-fragment layout, u[k][j = g][i = 2t, 2t+1] for all eight planes k
// (one 16-byte shared-memory load per plane).  Per plane:
//   d/dxi_0 (contract i): the lane's two values ARE the A operand of two k-steps whose contracted index is ordered
//            i = (2t | 2t + 1); the B operand is the matching pair of rows of D, fixed registers.  No extra load.
//   d/dxi_1 (contract j): A = D^T (fixed registers), B = the plane in the fragment layout u[j = 4s + t][i = g]
//            (two 8-byte loads, 32 consecutive doubles per k-step: conflict-free).
//   d/dxi_2 (contract k): the whole k-line of the lane's two points is in its registers: plain DFMA.
// All three results come out in the SAME layout (j = g, i = 2t, 2t+1), so the 3 x 3 constant factors are applied in
// registers and the three outputs leave as 16-byte stores, 512 contiguous bytes per warp instruction.  Shared memory
// is touched to read the input only (3 loads per plane and lane; the pencil kernel makes three passes).
struct HexDmmaPdTab
{
    double D[64]; // D[a*8 + b] = dh_a/dz(z_b)
};
struct HexDmmaPdArgs
{
    const double *in;
    double *out0, *out1, *out2;
    const double *df; // [9][dfStride]
    size_t dfStride;
    int nElmt;
    int in_aligned, out_aligned; // 16-byte aligned
};
struct HexDmmaPdCfg
{
    static constexpr int NQ = 8, NQ3 = 512, BUF = 2 * NQ3, PER_WARP = 2 * BUF + 2, WARPS = 12, T = WARPS * 32;
    static constexpr size_t SMEM = (size_t)WARPS * PER_WARP * 8 + 16;
};

__global__ void __launch_bounds__(HexDmmaPdCfg::T, 1)
    hex_dmma_pd_kernel(const __grid_constant__ HexDmmaPdTab tab, const __grid_constant__ HexDmmaPdArgs args)
{
    using Cfg = HexDmmaPdCfg;
    constexpr int NQ = Cfg::NQ, NQ3 = Cfg::NQ3, BUF = Cfg::BUF;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = lane & 3, g = lane >> 2;
    double *wbase = reinterpret_cast<double *>(smem_raw) + (size_t)warp * Cfg::PER_WARP;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase + 2 * BUF);

    const int nPairs = (args.nElmt + 1) / 2;
    const int GW = gridDim.x * Cfg::WARPS, gw = blockIdx.x * Cfg::WARPS + warp;
    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_fence_init();
    }
    __syncwarp();
    auto pair_ne = [&](int pr) { return args.nElmt - 2 * pr >= 2 ? 2 : 1; };
    auto tma_ok  = [&](int pr) { return args.in_aligned && pair_ne(pr) == 2; };
    auto issue   = [&](int pr, int slot) { // lane 0
        if (!tma_ok(pr)) return;
        mbar_expect_tx(bar + slot, (uint32_t)(BUF * 8));
        tma_load_1d(wbase + slot * BUF, args.in + (size_t)pr * BUF, (uint32_t)(BUF * 8), bar + slot);
    };
    // fixed fragments of the collocation derivative matrix
    const double b00 = tab.D[(2 * t) * NQ + g], b01 = tab.D[(2 * t + 1) * NQ + g]; // d/dxi_0: B[kk = t][n = g], i = 2t | 2t+1
    const double a10 = tab.D[t * NQ + g], a11 = tab.D[(4 + t) * NQ + g];           // d/dxi_1: A[m = g][kk = t], j = t | 4+t

    uint32_t phase[2] = {0, 0};
    int slot          = 0;
    if (lane == 0 && gw < nPairs) issue(gw, 0);
    for (int pr = gw; pr < nPairs; pr += GW, slot ^= 1)
    {
        const int ne = pair_ne(pr);
        double *sIn  = wbase + slot * BUF;
        if (lane == 0 && pr + GW < nPairs) issue(pr + GW, slot ^ 1); // the other buffer was consumed one trip ago
        if (tma_ok(pr))
        {
            mbar_wait(bar + slot, phase[slot]);
            phase[slot] ^= 1;
        }
        else
        {
            const double *src = args.in + (size_t)pr * BUF;
            for (int i = lane; i < ne * NQ3; i += 32) sIn[i] = __ldg(src + i);
        }
        __syncwarp();
#pragma unroll 1
        for (int e = 0; e < ne; ++e)
        {
            const size_t el = (size_t)pr * 2 + e;
            const double *U = sIn + e * NQ3;
            double f[9];
#pragma unroll
            for (int n = 0; n < 9; ++n) f[n] = __ldg(args.df + (size_t)n * args.dfStride + el);
            double2 u[NQ];
#pragma unroll
            for (int k = 0; k < NQ; ++k) u[k] = *reinterpret_cast<const double2 *>(U + k * (NQ * NQ) + g * NQ + 2 * t);
            const size_t o = el * NQ3 + g * NQ + 2 * t;
#pragma unroll
            for (int k = 0; k < NQ; ++k)
            {
                double d0x = 0.0, d0y = 0.0, d1x = 0.0, d1y = 0.0;
                dm_mma(d0x, d0y, u[k].x, b00);
                dm_mma(d0x, d0y, u[k].y, b01);
                const double q0 = U[k * (NQ * NQ) + t * NQ + g], q1 = U[k * (NQ * NQ) + (4 + t) * NQ + g];
                dm_mma(d1x, d1y, a10, q0);
                dm_mma(d1x, d1y, a11, q1);
                double d2x = tab.D[k] * u[0].x, d2y = tab.D[k] * u[0].y;
#pragma unroll
                for (int m = 1; m < NQ; ++m)
                {
                    d2x = fma(tab.D[m * NQ + k], u[m].x, d2x);
                    d2y = fma(tab.D[m * NQ + k], u[m].y, d2y);
                }
                const double2 r0 = make_double2(fma(f[2], d2x, fma(f[1], d1x, f[0] * d0x)), fma(f[2], d2y, fma(f[1], d1y, f[0] * d0y)));
                const double2 r1 = make_double2(fma(f[5], d2x, fma(f[4], d1x, f[3] * d0x)), fma(f[5], d2y, fma(f[4], d1y, f[3] * d0y)));
                const double2 r2 = make_double2(fma(f[8], d2x, fma(f[7], d1x, f[6] * d0x)), fma(f[8], d2y, fma(f[7], d1y, f[6] * d0y)));
                const size_t ok = o + k * (NQ * NQ);
                if (args.out_aligned)
                {
                    *reinterpret_cast<double2 *>(args.out0 + ok) = r0;
                    *reinterpret_cast<double2 *>(args.out1 + ok) = r1;
                    *reinterpret_cast<double2 *>(args.out2 + ok) = r2;
                }
                else
                {
                    args.out0[ok] = r0.x; args.out0[ok + 1] = r0.y;
                    args.out1[ok] = r1.x; args.out1[ok + 1] = r1.y;
                    args.out2[ok] = r2.x; args.out2[ok + 1] = r2.y;
                }
            }
        }
        __syncwarp(); // every lane is done with this buffer before lane 0 refills it
    }
}

struct HexDmmaPdState
{
    HexDmmaPdTab tab;
    void *fallback_state          = nullptr;
    void (*fallback_free)(void *) = nullptr;
    int bps                       = 0;
};

static int hex_dmma_pd_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    auto *st  = static_cast<HexDmmaPdState *>(op->kstate);
    using Cfg = HexDmmaPdCfg;
    auto kern = hex_dmma_pd_kernel;
    if (st->bps == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("hex DMMA PhysDeriv kernel does not fit on an SM"); return NEKMF_ERR_CUDA; }
        st->bps = nb;
    }
    HexDmmaPdArgs a;
    a.in = in[0]; a.out0 = out[0]; a.out1 = out[1]; a.out2 = out[2];
    a.df = op->d_df + (size_t)op->run_e0; a.dfStride = (size_t)op->nElmt; a.nElmt = op->run_ne;
    a.in_aligned  = (((uintptr_t)in[0]) & 15) == 0;
    a.out_aligned = (((uintptr_t)out[0] | (uintptr_t)out[1] | (uintptr_t)out[2]) & 15) == 0;
    const int nPairs = (op->run_ne + 1) / 2;
    int grid         = st->bps * NUM_SMS;
    const int need   = (nPairs + Cfg::WARPS - 1) / Cfg::WARPS;
    if (grid > need) grid = need;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(st->tab, a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

static void hex_dmma_pd_wrap(nekmf_op_s *op)
{
    auto *st = new HexDmmaPdState;
    memcpy(st->tab.D, op->D[0].data(), sizeof(st->tab.D));
    st->fallback_state = op->kstate;
    st->fallback_free  = op->kstate_free;
    op->kstate         = st;
    op->kstate_free    = [](void *p) {
        auto *s = static_cast<HexDmmaPdState *>(p);
        if (s->fallback_state && s->fallback_free) s->fallback_free(s->fallback_state);
        delete s;
    };
    op->launch = hex_dmma_pd_launch;
    op->kname  = "hex_dmma_pd_kernel<nm=7,nq=8,regular>(DMMA m8n8k4: xi_0 / xi_1 on tensor tiles, xi_2 in the owning lane)";
}

// called from select_hex_fast after the DFMA launcher is installed: default quadrature, BwdTrans or IProductWRTBase.
// NEKMF_HEX_DMMA=0 keeps the DFMA kernels (the other arm of the A/B), NEKMF_HEX_DMMA=all takes the tensor-core
// kernel at every instantiated order; the default is the set of (operator, order) cells where it measured faster.
void hex_dmma_maybe_wrap(nekmf_op_s *op)
{
    if (op->shape != NEKMF_HEX || op->nq[0] != op->nm[0] + 1) return;
    const int nm  = op->nm[0];
    const char *v = getenv("NEKMF_HEX_DMMA");
    if (v && v[0] == '0') return;
    if (op->optype == NEKMF_PHYSDERIV && !op->deformed && nm == 7)
    {
        hex_dmma_pd_wrap(op); // nq = 8 fills the tiles exactly; deformed collections stream geometry at the HBM peak already
        return;
    }
    if (op->optype != NEKMF_BWDTRANS && op->optype != NEKMF_IPRODUCTWRTBASE) return;
    const bool all = v && v[0] == 'a';
    // measured A/B (profiles/r02_sweep_hex_dmma_{all,0}.jsonl): every nm = 7 cell (0.77-0.90 against 0.47-0.78 of the
    // HBM peak); the half-empty second row tile loses to the pencil kernels everywhere else (regular IProductWRTBase at
    // nm = 11 was 0.39 against 0.36 until the pencil kernel went to one element per CTA: 0.56,
    // profiles/r02_sweep_hex_nodmma_v2.jsonl)
    const bool faster = nm == 7;
    if (!all && !faster) return;
    switch (nm)
    {
        case 7: hex_dmma_wrap<7>(op); break;
        case 8: hex_dmma_wrap<8>(op); break;
        case 9: hex_dmma_wrap<9>(op); break;
        case 10: hex_dmma_wrap<10>(op); break;
        case 11: hex_dmma_wrap<11>(op); break;
        default: break;
    }
}

} // namespace nekmf
