// hex_dmma.cu -- BwdTrans and IProductWRTBase on hexahedra at nm = 7, nq = 8 with FP64 tensor-core tiles
// (DMMA, mma.sync.m8n8k4.f64): the DMMA arm of the DMMA-vs-DFMA comparison of BASELINE.json configs[2].
//
// Reference semantics: MatrixFreeOps/BwdTransKernels.hpp:302-372, IProductKernels.hpp:236-314 (three 1-D
// contractions with the basis matrix B[p][i]); results differ from the DFMA kernels by summation order only.
//
// Why only nq = 8.  On B200 the DMMA pipe has the DFMA pipe's peak (37.1 vs 33.8 TFLOP/s measured,
// tools/fp64_peak.cu), so tensor tiles cannot raise the ceiling; what they can do is halve the shared-memory
// traffic that bounds the pencil kernels (hex_kernels.cuh: every pass reads and writes the whole intermediate),
// by CHAINING two contractions in registers:
//   pass 1 (contract p):  C1[i][q] = sum_p B[i][p] U[p][q]      A = basis matrix (8 x 8, rows i), B operand = the
//                          coefficient block read from shared memory, tile columns n <-> q
//   pass 2 (contract q):  C2[j][i] = sum_q B[j][q] C1[i][q]     B operand = the C fragment of pass 1 AS IT IS
// The m8n8k4 C fragment gives lane (t = lane % 4, g = lane / 4) row g, columns 2t and 2t+1; the B fragment of
// k-step s wants from that lane row k = t, column g.  With the columns of pass 1 ordered q(2t) = t, q(2t+1) = 4 + t
// the lane already holds C1[i = g][q = t] and C1[i = g][q = 4 + t] -- exactly the B-operand entries of the two
// k-steps of pass 2 (whose A fragments take the matching columns q = t, 4 + t of the basis matrix: the SAME two
// registers as in pass 1).  No shuffle, no shared memory between the passes.  The third contraction runs over the
// index the tiles are enumerated by (r resp. k), whose full line sits in ONE lane's registers: plain DFMA with the
// matrix entries as constant-bank operands, and the results go to global memory straight from registers
// (BwdTrans: one 512-byte contiguous warp store per k-plane).  Shared memory is touched once, for the input.
// That fits nq = 8 (nm = 7) exactly -- M = 8 rows, K = 7 -> 8 (12 % padding); at nq = 9..12 two row tiles waste
// 25-44 % of the pipe that is already the limiter there (DESIGN.md 4.2c has the arithmetic and the measurements).
//
// Mapping: every warp is an independent worker (as hex_kron.cu): element pairs (16-byte aligned 5488 B / 8192 B
// blocks) arrive by TMA bulk copies into the warp's own double buffer, completion on the warp's own mbarrier.
#include "op_internal.h"
#include <stdlib.h>
#include <string.h>

namespace nekmf
{

constexpr int DM_NM = 7, DM_NQ = 8, DM_NM3 = 343, DM_NQ3 = 512;

struct HexDmmaTab
{
    double B[DM_NM * DM_NQ]; // bdata[m*NQ + i]
    double w[DM_NQ];         // quadrature weights
};

struct HexDmmaArgs
{
    const double *in;
    double *out;
    const double *jac; // IProduct: [nElmt] | [nElmt][512]
    int nElmt;
    int in_aligned, out_aligned; // 16-byte aligned
};

__device__ __forceinline__ void dm_mma(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// OP 0: BwdTrans (343 -> 512 per element), 1: IProductWRTBase (512 -> 343); DEF: per-point Jacobian (IProduct only)
template <int OP, bool DEF> struct HexDmmaCfg
{
    static constexpr int IN_EL  = OP == 0 ? DM_NM3 : DM_NQ3;
    static constexpr int BUF    = 2 * IN_EL + ((2 * IN_EL) & 1);          // one element PAIR
    static constexpr int NARR   = (OP == 1 && DEF) ? 2 : 1;               // input (+ jacobian)
    static constexpr int PER_WARP = 2 * NARR * BUF + 2;                   // double buffer + mbarriers (2 x 8 B)
    static constexpr int WARPS  = NARR == 2 ? 6 : 12;                     // 33 KB / 11-16 KB of shared memory per warp
    static constexpr int T      = WARPS * 32;
    static constexpr size_t SMEM = (size_t)WARPS * PER_WARP * 8 + 16;
};

template <int OP, bool DEF>
__global__ void __launch_bounds__(HexDmmaCfg<OP, DEF>::T, 1)
    hex_dmma_kernel(const __grid_constant__ HexDmmaTab tab, const __grid_constant__ HexDmmaArgs args)
{
    using Cfg = HexDmmaCfg<OP, DEF>;
    constexpr int IN_EL = Cfg::IN_EL, BUF = Cfg::BUF, NARR = Cfg::NARR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = lane & 3, g = lane >> 2;
    double *wbase = reinterpret_cast<double *>(smem_raw) + (size_t)warp * Cfg::PER_WARP;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase + 2 * NARR * BUF);

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
        double *dst          = wbase + slot * NARR * BUF;
        const uint32_t bytes = 2 * IN_EL * 8;
        mbar_expect_tx(bar + slot, bytes * NARR);
        tma_load_1d(dst, args.in + (size_t)pr * 2 * IN_EL, bytes, bar + slot);
        if (NARR == 2) tma_load_1d(dst + BUF, args.jac + (size_t)pr * 2 * IN_EL, bytes, bar + slot);
    };

    // A fragments of the basis matrix: BwdTrans rows are quadrature indices (B[i = g][p]), IProduct rows are modes
    // (B^T[p = g][i]); columns t and 4 + t of the contracted index.  The same two registers serve pass 1 and pass 2.
    double a0, a1;
    if (OP == 0)
    {
        a0 = tab.B[t * DM_NQ + g];
        a1 = 4 + t < DM_NM ? tab.B[(4 + t) * DM_NQ + g] : 0.0;
    }
    else
    {
        a0 = g < DM_NM ? tab.B[g * DM_NQ + t] : 0.0;
        a1 = g < DM_NM ? tab.B[g * DM_NQ + 4 + t] : 0.0;
    }
    // tile column g of pass 1 <-> second index (q resp. j) = g/2 for even g, 4 + g/2 for odd g
    const int col = (g & 1) ? 4 + (g >> 1) : (g >> 1);

    uint32_t phase[2] = {0, 0};
    int slot          = 0;
    if (lane == 0 && gw < nPairs) issue(gw, 0);

    for (int pr = gw; pr < nPairs; pr += GW, slot ^= 1)
    {
        const int ne = pair_ne(pr);
        double *sIn  = wbase + slot * NARR * BUF;
        if (lane == 0 && pr + GW < nPairs) issue(pr + GW, slot ^ 1); // the other buffer was consumed one trip ago
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
            if (NARR == 2)
            {
                const double *sj = args.jac + (size_t)pr * 2 * IN_EL;
                for (int i = lane; i < ne * IN_EL; i += 32) sIn[BUF + i] = __ldg(sj + i);
            }
        }
        __syncwarp();

#pragma unroll 1
        for (int e = 0; e < ne; ++e)
        {
            const size_t el = (size_t)pr * 2 + e;
            if (OP == 0)
            {
                // ---------------------------------------------------------------- BwdTrans
                const double *U = sIn + e * DM_NM3;
                double T2[DM_NM][2];
#pragma unroll
                for (int r = 0; r < DM_NM; ++r)
                {
                    const double *row = U + r * (DM_NM * DM_NM) + col * DM_NM;
                    const double b0   = col < DM_NM ? row[t] : 0.0;
                    const double b1   = (col < DM_NM && 4 + t < DM_NM) ? row[4 + t] : 0.0;
                    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                    dm_mma(c0, c1, a0, b0); // C1[i = g][q = t], C1[i = g][q = 4 + t]
                    dm_mma(c0, c1, a1, b1);
                    dm_mma(d0, d1, a0, c0); // C2[j = g][i = 2t], C2[j = g][i = 2t + 1]
                    dm_mma(d0, d1, a1, c1);
                    T2[r][0] = d0;
                    T2[r][1] = d1;
                }
                double *o = args.out + el * DM_NQ3 + g * DM_NQ + 2 * t;
#pragma unroll 2 // (full unrolling makes the compiler hoist all 56 matrix entries into registers: spills)
                for (int k = 0; k < DM_NQ; ++k)
                {
                    double o0 = tab.B[k] * T2[0][0], o1 = tab.B[k] * T2[0][1];
#pragma unroll
                    for (int r = 1; r < DM_NM; ++r)
                    {
                        o0 = fma(tab.B[r * DM_NQ + k], T2[r][0], o0);
                        o1 = fma(tab.B[r * DM_NQ + k], T2[r][1], o1);
                    }
                    if (args.out_aligned)
                        *reinterpret_cast<double2 *>(o + k * (DM_NQ * DM_NQ)) = make_double2(o0, o1);
                    else
                    {
                        o[k * (DM_NQ * DM_NQ)]     = o0;
                        o[k * (DM_NQ * DM_NQ) + 1] = o1;
                    }
                }
            }
            else
            {
                // ---------------------------------------------------------------- IProductWRTBase
                const double *F = sIn + e * DM_NQ3;
                const double *J = sIn + BUF + e * DM_NQ3; // DEF only
                const double jreg = DEF ? 0.0 : __ldg(args.jac + el);
                const double w0 = tab.w[t] * tab.w[col], w1 = tab.w[4 + t] * tab.w[col];
                double acc[DM_NM][2];
#pragma unroll
                for (int r = 0; r < DM_NM; ++r) acc[r][0] = acc[r][1] = 0.0;
#pragma unroll 2
                for (int k = 0; k < DM_NQ; ++k)
                {
                    const int base = k * (DM_NQ * DM_NQ) + col * DM_NQ;
                    double b0 = F[base + t] * w0, b1 = F[base + 4 + t] * w1;
                    if (DEF)
                    {
                        b0 *= J[base + t] * tab.w[k];
                        b1 *= J[base + 4 + t] * tab.w[k];
                    }
                    else
                    {
                        const double s = jreg * tab.w[k];
                        b0 *= s;
                        b1 *= s;
                    }
                    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                    dm_mma(c0, c1, a0, b0); // C1[p = g][j = t], C1[p = g][j = 4 + t]
                    dm_mma(c0, c1, a1, b1);
                    dm_mma(d0, d1, a0, c0); // C2[q = g][p = 2t], C2[q = g][p = 2t + 1]
                    dm_mma(d0, d1, a1, c1);
#pragma unroll
                    for (int r = 0; r < DM_NM; ++r)
                    {
                        acc[r][0] = fma(tab.B[r * DM_NQ + k], d0, acc[r][0]);
                        acc[r][1] = fma(tab.B[r * DM_NQ + k], d1, acc[r][1]);
                    }
                }
                if (g < DM_NM)
                {
                    double *o = args.out + el * DM_NM3 + g * DM_NM + 2 * t;
#pragma unroll
                    for (int r = 0; r < DM_NM; ++r)
                    {
                        o[r * (DM_NM * DM_NM)] = acc[r][0];
                        if (2 * t + 1 < DM_NM) o[r * (DM_NM * DM_NM) + 1] = acc[r][1];
                    }
                }
            }
        }
        __syncwarp(); // every lane is done with this buffer before lane 0 refills it two trips later
    }
}

struct HexDmmaState
{
    HexDmmaTab tab;
    int (*fallback)(nekmf_op_s *, const double *const in[3], double *const out[3]) = nullptr;
    void *fallback_state                                                           = nullptr;
    void (*fallback_free)(void *)                                                  = nullptr;
    int bps = 0;
};

template <int OP, bool DEF> static int hex_dmma_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    HexDmmaState *st = static_cast<HexDmmaState *>(op->kstate);
    using Cfg        = HexDmmaCfg<OP, DEF>;
    auto kern        = hex_dmma_kernel<OP, DEF>;
    if (st->bps == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("hex DMMA kernel does not fit on an SM"); return NEKMF_ERR_CUDA; }
        st->bps = nb;
    }
    HexDmmaArgs a;
    a.in = in[0]; a.out = out[0]; a.nElmt = op->run_ne;
    a.jac = nullptr;
    if (OP == 1) a.jac = DEF ? op->d_jac + (size_t)op->run_e0 * op->geo_pitch : op->d_jac + op->run_e0;
    a.in_aligned  = ((((uintptr_t)in[0]) | (DEF && OP == 1 ? (uintptr_t)a.jac : 0)) & 15) == 0;
    a.out_aligned = (((uintptr_t)out[0]) & 15) == 0;
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

// called from select_hex_fast after the DFMA launcher is installed: nm = 7, default quadrature, BwdTrans or
// IProductWRTBase.  NEKMF_HEX_DMMA=0 keeps the DFMA kernel (the other arm of the A/B).
void hex_dmma_maybe_wrap(nekmf_op_s *op)
{
    if (op->shape != NEKMF_HEX || op->nm[0] != DM_NM || op->nq[0] != DM_NQ) return;
    if (op->optype != NEKMF_BWDTRANS && op->optype != NEKMF_IPRODUCTWRTBASE) return;
    const char *v = getenv("NEKMF_HEX_DMMA");
    if (v && v[0] == '0') return;
    // deformed IProduct reads the Jacobian with the element pitch of the TMA-fed kernels: 512 is even, pitch == nq^3
    if (op->optype == NEKMF_IPRODUCTWRTBASE && op->deformed && op->geo_pitch != DM_NQ3) return;
    HexDmmaState *st = new HexDmmaState;
    memcpy(st->tab.B, op->b[0].data(), sizeof(st->tab.B));
    memcpy(st->tab.w, op->ws[0].data(), sizeof(st->tab.w));
    st->fallback       = op->launch;
    st->fallback_state = op->kstate;
    st->fallback_free  = op->kstate_free;
    op->kstate         = st;
    op->kstate_free    = [](void *p) {
        HexDmmaState *s = static_cast<HexDmmaState *>(p);
        if (s->fallback_state && s->fallback_free) s->fallback_free(s->fallback_state);
        delete s;
    };
    if (op->optype == NEKMF_BWDTRANS)
    {
        op->launch = hex_dmma_launch<0, false>;
        op->kname  = "hex_dmma_kernel<bwd,nm=7,nq=8>(DMMA m8n8k4, two contractions chained in registers)";
    }
    else
    {
        op->launch = op->deformed ? hex_dmma_launch<1, true> : hex_dmma_launch<1, false>;
        op->kname  = op->deformed ? "hex_dmma_kernel<iprod,nm=7,nq=8,deformed>(DMMA m8n8k4)"
                                  : "hex_dmma_kernel<iprod,nm=7,nq=8,regular>(DMMA m8n8k4)";
    }
}

} // namespace nekmf
