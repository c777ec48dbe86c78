// hex_slab.cuh -- BwdTrans and IProductWRTBase on hexahedra as warp-independent register-slab kernels.
//
// Reference semantics (what is computed, not how):
//   BwdTrans         MatrixFreeOps/BwdTransKernels.hpp:302-372   u[k][j][i] = sum_pqr c[r][q][p] B[p][i] B[q][j] B[r][k]
//   IProductWRTBase  MatrixFreeOps/IProductKernels.hpp:236-314   c[r][q][p] = sum_kji f[k][j][i] J w_i w_j w_k B[p][i] B[q][j] B[r][k]
//
// The pencil kernels of hex_kernels.cuh are bound by shared-memory wavefronts (three passes, each value read and
// written once per pass, 35-40 % bank conflicts).  Here a LANE owns a whole 2-D slab of its element in registers
// and does TWO of the three contractions there; one transposing exchange through a conflict-free padded layout
// feeds the third:
//   BwdTrans : lane (e,r) holds c[r][.][.], contracts p->i and q->j, scatters w[r][j][i]; lane (e,j) then gathers
//              the r-lines of its (j,.) row and contracts r->k into the output staging block;
//   IProduct : lane (e,k) holds (f J w)[k][.][.], contracts i->p and j->q, scatters v[k][q][p]; lane (e,q) gathers
//              the k-lines of its (q,.) row and contracts k->r.
// Every warp is an independent worker (own TMA-fed input buffer, own mbarrier, own bulk stores; no CTA barrier).
// Slabs whose length is even (coefficient slabs for even nm, quadrature slabs for odd nm) live in slots padded by
// two doubles, filled by warp-wide 16-byte cp.async copies (inputs) and drained by per-slab bulk stores (BwdTrans)
// or warp-wide 16-byte stores (IProduct); odd-length slabs travel as one bulk copy per batch.  In both cases the
// lane-strided slab accesses are at most 2-way bank conflicted.  Matrix entries are kernel-parameter constants.
#pragma once
#include "hex_kernels.cuh"

namespace nekmf
{

constexpr int slab_pad16(int minimum, int residue) // smallest v >= minimum with v % 16 == residue % 16
{
    int v = minimum;
    while (v % 16 != residue % 16) ++v;
    return v;
}

// 16-byte asynchronous global -> shared copy (LDGSTS.128): one warp-wide instruction moves 512 contiguous bytes
// into arbitrary 16-byte aligned shared-memory slots -- the way to fill PADDED slab slots without one bulk-copy
// issue per lane (per-lane cp.async.bulk instructions serialise: ~50 cycles each, measured on the PhysDeriv kernel)
__device__ __forceinline__ void slab_cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void slab_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int OP, int NM> struct SlabCfg
{
    static constexpr int NQ = NM + 1, NM2 = NM * NM, NM3 = NM2 * NM, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ;
    // elements per warp step: the wider stage uses NQ lanes per element; kept even because one of nm^3, nq^3 is
    // always odd and a batch must be a whole number of 16-byte units to travel by TMA
    static constexpr int EPW = (32 / NQ) >= 2 ? ((32 / NQ) & ~1) : 1;
    // coefficient block [e][r][q][p]: slab stride CS, element stride CE
    static constexpr bool CPAD = (NM % 2) == 0;
    static constexpr int CS = CPAD ? NM2 + 2 : NM2, CE = NM * CS, CBUF = round_up(EPW * CE, 2);
    // quadrature block [e][k][j][i]: slab stride PS, element stride PE
    static constexpr bool PPAD = (NQ % 2) == 0;
    static constexpr int PS = PPAD ? NQ2 + 2 : NQ2, PE = NQ * PS, PBUF = round_up(EPW * PE, 2);
    // exchange block.  BwdTrans: X[e*EX + j*LS + i*NM + r] (r-lines of length NM, rows j of NQ lines);
    //                  IProduct: X[e*EX + q*LS + p*NQ + k] (k-lines of length NQ, rows q of NM lines)
    static constexpr int LINE = OP == HEX_BWD ? NM : NQ;   // line length = lanes per element of stage I
    static constexpr int ROWS = OP == HEX_BWD ? NQ : NM;   // rows = lanes per element of stage II
    static constexpr int LPR  = OP == HEX_BWD ? NQ : NM;   // lines per row
    static constexpr int LS   = (LPR * LINE) | 1;          // odd row stride: stage II lanes hit distinct banks
    static constexpr int EX   = slab_pad16(ROWS * LS, LINE); // stage I: the elements of a half-warp tile the banks
    static constexpr int XBUF = round_up(EPW * EX, 2);
    static constexpr int INBUF  = OP == HEX_BWD ? CBUF : PBUF;
    static constexpr int OUTBUF = OP == HEX_BWD ? PBUF : CBUF;
    static constexpr int PER_WARP = INBUF + XBUF + OUTBUF + 2; // doubles (+2: mbarrier slot)
    static constexpr int W_FIT  = (216 * 1024) / (PER_WARP * 8);
    static constexpr int WARPS  = W_FIT >= 12 ? 12 : (W_FIT >= 8 ? 8 : (W_FIT >= 1 ? W_FIT : 1));
    static constexpr int T      = WARPS * 32;
    static constexpr size_t SMEM = (size_t)WARPS * PER_WARP * 8 + 16;
};

struct SlabArgs
{
    const double *in;
    double *out;
    const double *jac; // IProduct: [nElmt] (regular geometry; deformed collections keep the pencil kernel, which
                       // already streams the per-point Jacobian at 0.9-1.0 of the HBM peak)
    int nElmt;
    int io_aligned; // in and out 16-byte aligned
};

template <int OP, int NM>
__global__ void __launch_bounds__(SlabCfg<OP, NM>::T, 1)
    hex_slab_kernel(const __grid_constant__ HexTab<NM, NM + 1> tab, const __grid_constant__ SlabArgs args)
{
    using Cfg = SlabCfg<OP, NM>;
    constexpr int NQ = Cfg::NQ, NM2 = Cfg::NM2, NM3 = Cfg::NM3, NQ2 = Cfg::NQ2, NQ3 = Cfg::NQ3, EPW = Cfg::EPW;
    constexpr int CS = Cfg::CS, CE = Cfg::CE, PS = Cfg::PS, PE = Cfg::PE, LS = Cfg::LS, EX = Cfg::EX;
    constexpr bool BWD = OP == HEX_BWD;
    // input side / output side geometry of the shared-memory blocks
    constexpr int IN_SLABS = BWD ? NM : NQ, IN_SLEN = BWD ? NM2 : NQ2, IN_SS = BWD ? CS : PS, IN_ES = BWD ? CE : PE;
    constexpr int IN_EL = BWD ? NM3 : NQ3;
    constexpr bool IN_PAD = BWD ? Cfg::CPAD : Cfg::PPAD;
    constexpr int OUT_SLABS = BWD ? NQ : NM, OUT_SLEN = BWD ? NQ2 : NM2, OUT_SS = BWD ? PS : CS, OUT_ES = BWD ? PE : CE;
    constexpr int OUT_EL = BWD ? NQ3 : NM3;
    constexpr bool OUT_PAD = BWD ? Cfg::PPAD : Cfg::CPAD;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *wbase  = reinterpret_cast<double *>(smem_raw) + (size_t)warp * Cfg::PER_WARP;
    double *sIn    = wbase;
    double *sX     = sIn + Cfg::INBUF;
    double *sOut   = sX + Cfg::XBUF;
    uint64_t *bar  = reinterpret_cast<uint64_t *>(sOut + Cfg::OUTBUF);

    const int nElmt = args.nElmt;
    const int nWB   = (nElmt + EPW - 1) / EPW;
    const int GW    = gridDim.x * Cfg::WARPS;
    const int gw    = blockIdx.x * Cfg::WARPS + warp;
    // stage I lane = (e1, s1): s1 < IN_SLABS ; stage II lane = (e2, s2): s2 < OUT rows
    const int e1 = lane / IN_SLABS, s1 = lane - e1 * IN_SLABS;
    const int e2 = lane / Cfg::ROWS, s2 = lane - e2 * Cfg::ROWS;

    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto batch_ne = [&](int wb) { int r = nElmt - wb * EPW; return r < EPW ? r : EPW; };
    auto in_tma   = [&](int wb) { return args.io_aligned && (IN_PAD || (((batch_ne(wb) * IN_EL) & 1) == 0 && ((wb * EPW * IN_EL) & 1) == 0)); };
    auto out_tma  = [&](int wb) { return args.io_aligned && (OUT_PAD || (((batch_ne(wb) * OUT_EL) & 1) == 0 && ((wb * EPW * OUT_EL) & 1) == 0)); };
    // padded blocks (even-length slabs): shared-memory address of double pair `i2` of a batch
    auto pad_in  = [&](int i2) { const int e = (2 * i2) / IN_EL, w = 2 * i2 - e * IN_EL, s = w / IN_SLEN; return e * IN_ES + s * IN_SS + (w - s * IN_SLEN); };
    auto pad_out = [&](int i2) { const int e = (2 * i2) / OUT_EL, w = 2 * i2 - e * OUT_EL, s = w / OUT_SLEN; return e * OUT_ES + s * OUT_SS + (w - s * OUT_SLEN); };
    auto issue    = [&](int wb) { // whole warp; the input buffer is free
        const int ne      = batch_ne(wb);
        if (!in_tma(wb)) return;
        const double *src = args.in + (size_t)wb * EPW * IN_EL;
        if (IN_PAD)
        {
            // warp-wide 16-byte asynchronous copies into the padded slots
            for (int i2 = lane; i2 < ne * IN_EL / 2; i2 += 32) slab_cp_async16(sIn + pad_in(i2), src + 2 * i2);
            return;
        }
        if (lane == 0)
        {
            mbar_expect_tx(bar, (uint32_t)(ne * IN_EL * 8));
            tma_load_1d(sIn, src, (uint32_t)(ne * IN_EL * 8), bar);
        }
    };

    uint32_t phase = 0;
    if (gw < nWB) issue(gw);
    for (int wb = gw; wb < nWB; wb += GW)
    {
        const int ne = batch_ne(wb), wbnext = wb + GW;
        const bool tin = in_tma(wb), tout = out_tma(wb);
        if (!tin)
        {
            const double *src = args.in + (size_t)wb * EPW * IN_EL;
            for (int i = lane; i < ne * IN_EL; i += 32)
            {
                const int e = i / IN_EL, w = i - e * IN_EL, s = w / IN_SLEN;
                sIn[e * IN_ES + s * IN_SS + (w - s * IN_SLEN)] = __ldg(src + i);
            }
        }
        else if (IN_PAD)
            slab_cp_async_wait_all();
        else
        {
            mbar_wait(bar, phase);
            phase ^= 1;
        }
        __syncwarp();

        // ------------------------------------------------------------------ stage I
        if (lane < EPW * IN_SLABS && e1 < ne)
        {
            const double *xs = sIn + e1 * IN_ES + s1 * IN_SS;
            double *X        = sX + e1 * EX;
            if (BWD)
            {
                // lane (e, r): c[q][p] -> w[j][i] = sum_q B[q][j] sum_p B[p][i] c[q][p]
                double c[NM][NM];
#pragma unroll
                for (int q = 0; q < NM; ++q)
#pragma unroll
                    for (int p = 0; p < NM; ++p) c[q][p] = xs[q * NM + p];
#pragma unroll
                for (int i = 0; i < NQ; ++i)
                {
                    double t[NM];
#pragma unroll
                    for (int q = 0; q < NM; ++q)
                    {
                        double s = tab.B[i] * c[q][0];
#pragma unroll
                        for (int p = 1; p < NM; ++p) s = fma(tab.B[p * NQ + i], c[q][p], s);
                        t[q] = s;
                    }
#pragma unroll
                    for (int j = 0; j < NQ; ++j)
                    {
                        double s = tab.B[j] * t[0];
#pragma unroll
                        for (int q = 1; q < NM; ++q) s = fma(tab.B[q * NQ + j], t[q], s);
                        X[j * LS + i * NM + s1] = s;
                    }
                }
            }
            else
            {
                // lane (e, k): g[j][i] = f[j][i] J w_k w_j w_i -> v[q][p] = sum_j B[q][j] sum_i B[p][i] g[j][i]
                double g[NQ][NQ];
                const double jk = __ldg(args.jac + (size_t)wb * EPW + e1) * tab.w[s1];
#pragma unroll
                for (int j = 0; j < NQ; ++j)
#pragma unroll
                    for (int i = 0; i < NQ; ++i)
                    {
                        g[j][i] = xs[j * NQ + i] * ((jk * tab.w[j]) * tab.w[i]);
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
                        X[q * LS + p * NQ + s1] = s;
                    }
                }
            }
        }
        __syncwarp();
        // the input buffer is consumed: request the next batch now, it lands during stage II and the store; the
        // staging buffer must no longer be read by the previous batch's bulk stores (each issuing lane waits for
        // its own groups)
        if (wbnext < nWB) issue(wbnext);
        tma_store_wait_read0();
        __syncwarp();

        // ------------------------------------------------------------------ stage II
        if (lane < EPW * Cfg::ROWS && e2 < ne)
        {
            const double *X = sX + e2 * EX + s2 * LS;
            double *O       = sOut + e2 * OUT_ES;
            if (BWD)
            {
                // lane (e, j): u[k][j][i] = sum_r B[r][k] w[r][j][i]
#pragma unroll
                for (int i = 0; i < NQ; ++i)
                {
                    double v[NM];
#pragma unroll
                    for (int r = 0; r < NM; ++r) v[r] = X[i * NM + r];
#pragma unroll
                    for (int k = 0; k < NQ; ++k)
                    {
                        double s = tab.B[k] * v[0];
#pragma unroll
                        for (int r = 1; r < NM; ++r) s = fma(tab.B[r * NQ + k], v[r], s);
                        O[k * OUT_SS + s2 * NQ + i] = s;
                    }
                }
            }
            else
            {
                // lane (e, q): c[r][q][p] = sum_k B[r][k] v[k][q][p]
#pragma unroll
                for (int p = 0; p < NM; ++p)
                {
                    double v[NQ];
#pragma unroll
                    for (int k = 0; k < NQ; ++k) v[k] = X[p * NQ + k];
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                    {
                        double s = tab.B[r * NQ] * v[0];
#pragma unroll
                        for (int k = 1; k < NQ; ++k) s = fma(tab.B[r * NQ + k], v[k], s);
                        O[r * OUT_SS + s2 * NM + p] = s;
                    }
                }
            }
        }
        double *dst = args.out + (size_t)wb * EPW * OUT_EL;
        if (tout)
        {
            fence_proxy_async();
            __syncwarp();
            if (OUT_PAD && BWD)
            {
                // padded quadrature staging block (nq^3 per element): one bulk store per slab, issued by lane (eo, so)
                // -- measured faster than warp-wide 16-byte stores here (0.94 vs 0.85 of HBM peak at nm = 5)
                const int eo = lane / OUT_SLABS, so = lane - eo * OUT_SLABS;
                if (lane < EPW * OUT_SLABS && eo < ne)
                    tma_store_1d(dst + (size_t)eo * OUT_EL + so * OUT_SLEN, sOut + eo * OUT_ES + so * OUT_SS, (uint32_t)(OUT_SLEN * 8));
                tma_store_commit();
            }
            else if (OUT_PAD)
            {
                // padded coefficient staging block (small): coalesced 16-byte stores by the whole warp
                for (int i2 = lane; i2 < ne * OUT_EL / 2; i2 += 32)
                    *reinterpret_cast<double2 *>(dst + 2 * i2) = *reinterpret_cast<const double2 *>(sOut + pad_out(i2));
                __syncwarp();
            }
            else
            {
                if (lane == 0) tma_store_1d(dst, sOut, (uint32_t)(ne * OUT_EL * 8));
                tma_store_commit();
            }
        }
        else
        {
            __syncwarp();
            for (int i = lane; i < ne * OUT_EL; i += 32)
            {
                const int e = i / OUT_EL, w = i - e * OUT_EL, s = w / OUT_SLEN;
                dst[i] = sOut[e * OUT_ES + s * OUT_SS + (w - s * OUT_SLEN)];
            }
            __syncwarp();
        }
    }
    tma_store_wait0();
}

// ------------------------------------------------------------------------------------------------ PhysDeriv
// PhysDeriv on REGULAR hexahedra (MatrixFreeOps/PhysDerivKernels.hpp:219-372): out_c = sum_d df[3c+d] du/dxi_d.
// Lane (e,k) owns the quadrature slab u[k][.][.] in registers: the xi_0 and xi_1 derivatives are contractions
// inside the slab, the xi_2 derivative reads the other slabs of the element from the shared input block (all
// lanes of an element read the same word: a broadcast, no bank conflict) with the lane's own column D[.][k] held
// in registers.  No exchange, no barrier; the three results are written to the lane's own slabs of three staging
// blocks and leave by TMA bulk stores.  Deformed collections keep the pencil kernel (0.9-1.0 of the HBM peak).
template <int NM> struct PdSlabCfg
{
    static constexpr int NQ = NM + 1, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ;
    static constexpr int EPW = (32 / NQ) >= 2 ? ((32 / NQ) & ~1) : 1;
    static constexpr bool PPAD = (NQ % 2) == 0;
    static constexpr int PS = PPAD ? NQ2 + 2 : NQ2, PE = NQ * PS, PBUF = round_up(EPW * PE, 2);
    static constexpr int PER_WARP = 4 * PBUF + 2;
    static constexpr int W_FIT  = (216 * 1024) / (PER_WARP * 8);
    static constexpr int WARPS  = W_FIT >= 12 ? 12 : (W_FIT >= 8 ? 8 : (W_FIT >= 1 ? W_FIT : 1));
    static constexpr int T      = WARPS * 32;
    static constexpr size_t SMEM = (size_t)WARPS * PER_WARP * 8 + 16;
};

struct PdSlabArgs
{
    const double *in;
    double *out0, *out1, *out2;
    const double *df; // [9][dfStride], regular geometry: one entry per element
    size_t dfStride;
    int nElmt;
    int io_aligned; // in and the three outputs 16-byte aligned
};

template <int NM>
__global__ void __launch_bounds__(PdSlabCfg<NM>::T, 1)
    hex_pd_slab_kernel(const __grid_constant__ HexTab<NM, NM + 1> tab, const __grid_constant__ PdSlabArgs args)
{
    using Cfg = PdSlabCfg<NM>;
    constexpr int NQ = Cfg::NQ, NQ2 = Cfg::NQ2, NQ3 = Cfg::NQ3, EPW = Cfg::EPW, PS = Cfg::PS, PE = Cfg::PE;
    constexpr bool PPAD = Cfg::PPAD;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *wbase  = reinterpret_cast<double *>(smem_raw) + (size_t)warp * Cfg::PER_WARP;
    double *sIn    = wbase;
    double *sO0    = sIn + Cfg::PBUF, *sO1 = sO0 + Cfg::PBUF, *sO2 = sO1 + Cfg::PBUF;
    uint64_t *bar  = reinterpret_cast<uint64_t *>(sO2 + Cfg::PBUF);

    const int nElmt = args.nElmt;
    const int nWB   = (nElmt + EPW - 1) / EPW;
    const int GW    = gridDim.x * Cfg::WARPS;
    const int gw    = blockIdx.x * Cfg::WARPS + warp;
    const int e1 = lane / NQ, k1 = lane - e1 * NQ; // lane = (element, slab)
    const bool lane_on = lane < EPW * NQ;
    // the lane's column of the collocation derivative matrix: D[k'][k] = dh_k'/dz(z_k)
    double dcol[NQ];
#pragma unroll
    for (int m = 0; m < NQ; ++m) dcol[m] = lane_on ? tab.D[m * NQ + k1] : 0.0;

    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto batch_ne = [&](int wb) { int r = nElmt - wb * EPW; return r < EPW ? r : EPW; };
    auto tma_ok   = [&](int wb) { return args.io_aligned && (PPAD || (((batch_ne(wb) * NQ3) & 1) == 0 && ((wb * EPW * NQ3) & 1) == 0)); };
    // padded slots (even-length slabs): address of double pair `i2` of a batch in a padded block
    auto padded = [&](int i2) { const int e = (2 * i2) / NQ3, w = 2 * i2 - e * NQ3, s = w / NQ2; return e * PE + s * PS + (w - s * NQ2); };
    auto issue    = [&](int wb) { // whole warp; the input buffer is free
        const int ne = batch_ne(wb);
        if (!tma_ok(wb)) return;
        const double *src = args.in + (size_t)wb * EPW * NQ3;
        if (PPAD)
        {
            for (int i2 = lane; i2 < ne * NQ3 / 2; i2 += 32) slab_cp_async16(sIn + padded(i2), src + 2 * i2);
            return;
        }
        if (lane == 0)
        {
            mbar_expect_tx(bar, (uint32_t)(ne * NQ3 * 8));
            tma_load_1d(sIn, src, (uint32_t)(ne * NQ3 * 8), bar);
        }
    };

    uint32_t phase = 0;
    if (gw < nWB) issue(gw);
    for (int wb = gw; wb < nWB; wb += GW)
    {
        const int ne = batch_ne(wb), wbnext = wb + GW;
        const bool tma = tma_ok(wb);
        if (!tma)
        {
            const double *src = args.in + (size_t)wb * EPW * NQ3;
            for (int i = lane; i < ne * NQ3; i += 32)
            {
                const int e = i / NQ3, w = i - e * NQ3, s = w / NQ2;
                sIn[e * PE + s * PS + (w - s * NQ2)] = __ldg(src + i);
            }
        }
        else if (PPAD)
            slab_cp_async_wait_all();
        else
        {
            mbar_wait(bar, phase);
            phase ^= 1;
        }
        // the staging blocks must no longer be read by the previous batch's bulk stores
        tma_store_wait_read0();
        __syncwarp();

        if (lane_on && e1 < ne)
        {
            const double *ue = sIn + e1 * PE; // the element's whole block (slab stride PS)
            const double *us = ue + k1 * PS;  // the lane's slab
            double f[9];
#pragma unroll
            for (int n = 0; n < 9; ++n) f[n] = __ldg(args.df + (size_t)n * args.dfStride + (size_t)wb * EPW + e1);
            double u[NQ][NQ];
#pragma unroll
            for (int j = 0; j < NQ; ++j)
#pragma unroll
                for (int i = 0; i < NQ; ++i) u[j][i] = us[j * NQ + i];
            double *o0 = sO0 + e1 * PE + k1 * PS, *o1 = sO1 + e1 * PE + k1 * PS, *o2 = sO2 + e1 * PE + k1 * PS;
#pragma unroll
            for (int j = 0; j < NQ; ++j)
            {
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
                    double d2 = dcol[0] * ue[j * NQ + i];
#pragma unroll
                    for (int m = 1; m < NQ; ++m) d2 = fma(dcol[m], ue[m * PS + j * NQ + i], d2);
                    o0[j * NQ + i] = f[0] * d0 + f[1] * d1 + f[2] * d2;
                    o1[j * NQ + i] = f[3] * d0 + f[4] * d1 + f[5] * d2;
                    o2[j * NQ + i] = f[6] * d0 + f[7] * d1 + f[8] * d2;
                }
            }
        }
        __syncwarp();
        // every lane has read the whole input block: request the next batch, it lands during the stores
        if (wbnext < nWB) issue(wbnext);
        const size_t goff = (size_t)wb * EPW * NQ3;
        if (tma)
        {
            fence_proxy_async();
            __syncwarp();
            if (PPAD)
            {
                // padded staging blocks: coalesced 16-byte stores by the whole warp
                for (int i2 = lane; i2 < ne * NQ3 / 2; i2 += 32)
                {
                    const int a = padded(i2);
                    *reinterpret_cast<double2 *>(args.out0 + goff + 2 * i2) = *reinterpret_cast<const double2 *>(sO0 + a);
                    *reinterpret_cast<double2 *>(args.out1 + goff + 2 * i2) = *reinterpret_cast<const double2 *>(sO1 + a);
                    *reinterpret_cast<double2 *>(args.out2 + goff + 2 * i2) = *reinterpret_cast<const double2 *>(sO2 + a);
                }
                __syncwarp();
            }
            else if (lane == 0)
            {
                tma_store_1d(args.out0 + goff, sO0, (uint32_t)(ne * NQ3 * 8));
                tma_store_1d(args.out1 + goff, sO1, (uint32_t)(ne * NQ3 * 8));
                tma_store_1d(args.out2 + goff, sO2, (uint32_t)(ne * NQ3 * 8));
            }
            tma_store_commit();
        }
        else
        {
            for (int i = lane; i < ne * NQ3; i += 32)
            {
                const int e = i / NQ3, w = i - e * NQ3, s = w / NQ2, a = e * PE + s * PS + (w - s * NQ2);
                args.out0[goff + i] = sO0[a];
                args.out1[goff + i] = sO1[a];
                args.out2[goff + i] = sO2[a];
            }
            __syncwarp();
        }
    }
    tma_store_wait0();
}

} // namespace nekmf

namespace nekmf
{
// ------------------------------------------------------------------------------------- IProductWRTDerivBase
// IProductWRTDerivBase on REGULAR hexahedra (MatrixFreeOps/IProductWRTDerivBase.h:1232-1345):
//   t_d = sum_c df[3c+d] f_c ,   out = IP(dB,B,B)[t_0] + IP(B,dB,B)[t_1] + IP(B,B,dB)[t_2]   (each weighted by J w).
// Lane (e,k) makes three passes over its quadrature slab, one per reference direction d: it forms t_d J w for the
// slab in registers from the three input slabs (shared memory), contracts i->p and j->q with the matrices of that
// direction, and accumulates into two exchange blocks -- X_B (to be contracted with B along k: d = 0, 1) and X_D
// (with dB along k: d = 2); it only ever touches its own entries, so no barrier is needed between the passes.
// Lane (e,q) then contracts k for its row from both blocks.  Same layouts, padding and copy scheme as the
// IProductWRTBase slab kernel.  Deformed collections keep the pencil kernel (0.8-1.0 of the HBM peak).
template <int NM> struct IpwdbSlabCfg
{
    static constexpr int NQ = NM + 1, NM2 = NM * NM, NM3 = NM2 * NM, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ;
    static constexpr int EPW = (32 / NQ) >= 2 ? ((32 / NQ) & ~1) : 1;
    static constexpr bool CPAD = (NM % 2) == 0, PPAD = (NQ % 2) == 0;
    static constexpr int CS = CPAD ? NM2 + 2 : NM2, CE = NM * CS, CBUF = round_up(EPW * CE, 2);
    static constexpr int PS = PPAD ? NQ2 + 2 : NQ2, PE = NQ * PS, PBUF = round_up(EPW * PE, 2);
    static constexpr int LS = (NM * NQ) | 1;             // row q: NM lines (p) of NQ values (k)
    static constexpr int EX = slab_pad16(NM * LS, NQ);
    static constexpr int XBUF = round_up(EPW * EX, 2);
    static constexpr int PER_WARP = 3 * PBUF + 2 * XBUF + CBUF + 2;
    static constexpr int W_FIT  = (216 * 1024) / (PER_WARP * 8);
    static constexpr int WARPS  = W_FIT >= 12 ? 12 : (W_FIT >= 8 ? 8 : (W_FIT >= 1 ? W_FIT : 1));
    static constexpr int T      = WARPS * 32;
    static constexpr size_t SMEM = (size_t)WARPS * PER_WARP * 8 + 16;
};

template <int NM> struct SlabDTab
{
    double dB[NM * (NM + 1)]; // dbdata[m*NQ+i]
};

struct IpwdbSlabArgs
{
    const double *in0, *in1, *in2;
    double *out;
    const double *jac; // [nElmt]
    const double *df;  // [9][dfStride]
    size_t dfStride;
    int nElmt;
    int io_aligned;
};

template <int NM>
__global__ void __launch_bounds__(IpwdbSlabCfg<NM>::T, 1)
    hex_ipwdb_slab_kernel(const __grid_constant__ HexTab<NM, NM + 1> tab, const __grid_constant__ SlabDTab<NM> dtab,
                          const __grid_constant__ IpwdbSlabArgs args)
{
    using Cfg = IpwdbSlabCfg<NM>;
    constexpr int NQ = Cfg::NQ, NM2 = Cfg::NM2, NM3 = Cfg::NM3, NQ2 = Cfg::NQ2, NQ3 = Cfg::NQ3, EPW = Cfg::EPW;
    constexpr int CS = Cfg::CS, CE = Cfg::CE, PS = Cfg::PS, PE = Cfg::PE, LS = Cfg::LS, EX = Cfg::EX;
    constexpr bool CPAD = Cfg::CPAD, PPAD = Cfg::PPAD;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *wbase  = reinterpret_cast<double *>(smem_raw) + (size_t)warp * Cfg::PER_WARP;
    double *sF0 = wbase, *sF1 = sF0 + Cfg::PBUF, *sF2 = sF1 + Cfg::PBUF;
    double *sXB = sF2 + Cfg::PBUF, *sXD = sXB + Cfg::XBUF;
    double *sOut = sXD + Cfg::XBUF;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sOut + Cfg::CBUF);

    const int nElmt = args.nElmt;
    const int nWB   = (nElmt + EPW - 1) / EPW;
    const int GW    = gridDim.x * Cfg::WARPS;
    const int gw    = blockIdx.x * Cfg::WARPS + warp;
    const int e1 = lane / NQ, k1 = lane - e1 * NQ; // stage I lane = (element, slab k)
    const int e2 = lane / NM, q2 = lane - e2 * NM; // stage II lane = (element, row q)
    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto batch_ne = [&](int wb) { int r = nElmt - wb * EPW; return r < EPW ? r : EPW; };
    auto in_tma   = [&](int wb) { return args.io_aligned && (PPAD || (((batch_ne(wb) * NQ3) & 1) == 0 && ((wb * EPW * NQ3) & 1) == 0)); };
    auto out_tma  = [&](int wb) { return args.io_aligned && (CPAD || (((batch_ne(wb) * NM3) & 1) == 0 && ((wb * EPW * NM3) & 1) == 0)); };
    auto pad_in   = [&](int i2) { const int e = (2 * i2) / NQ3, w = 2 * i2 - e * NQ3, s = w / NQ2; return e * PE + s * PS + (w - s * NQ2); };
    auto pad_out  = [&](int i2) { const int e = (2 * i2) / NM3, w = 2 * i2 - e * NM3, s = w / NM2; return e * CE + s * CS + (w - s * NM2); };
    auto issue    = [&](int wb) { // whole warp; the three input buffers are free
        const int ne = batch_ne(wb);
        if (!in_tma(wb)) return;
        const size_t off = (size_t)wb * EPW * NQ3;
        if (PPAD)
        {
            for (int i2 = lane; i2 < ne * NQ3 / 2; i2 += 32)
            {
                const int a = pad_in(i2);
                slab_cp_async16(sF0 + a, args.in0 + off + 2 * i2);
                slab_cp_async16(sF1 + a, args.in1 + off + 2 * i2);
                slab_cp_async16(sF2 + a, args.in2 + off + 2 * i2);
            }
            return;
        }
        if (lane == 0)
        {
            const uint32_t bytes = (uint32_t)(ne * NQ3 * 8);
            mbar_expect_tx(bar, 3 * bytes);
            tma_load_1d(sF0, args.in0 + off, bytes, bar);
            tma_load_1d(sF1, args.in1 + off, bytes, bar);
            tma_load_1d(sF2, args.in2 + off, bytes, bar);
        }
    };

    uint32_t phase = 0;
    if (gw < nWB) issue(gw);
    for (int wb = gw; wb < nWB; wb += GW)
    {
        const int ne = batch_ne(wb), wbnext = wb + GW;
        const bool tin = in_tma(wb), tout = out_tma(wb);
        if (!tin)
        {
            const size_t off = (size_t)wb * EPW * NQ3;
            for (int i = lane; i < ne * NQ3; i += 32)
            {
                const int e = i / NQ3, w = i - e * NQ3, s = w / NQ2, a = e * PE + s * PS + (w - s * NQ2);
                sF0[a] = __ldg(args.in0 + off + i);
                sF1[a] = __ldg(args.in1 + off + i);
                sF2[a] = __ldg(args.in2 + off + i);
            }
        }
        else if (PPAD)
            slab_cp_async_wait_all();
        else
        {
            mbar_wait(bar, phase);
            phase ^= 1;
        }
        __syncwarp();

        // ------------------------------------------------------------------ stage I: lane (e, k), three passes
        if (lane < EPW * NQ && e1 < ne)
        {
            const int so     = e1 * PE + k1 * PS;
            const size_t eg  = (size_t)wb * EPW + e1;
            const double jwk = __ldg(args.jac + eg) * tab.w[k1];
            double *XB = sXB + e1 * EX, *XD = sXD + e1 * EX;
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
                const double c0 = __ldg(args.df + (size_t)(0 + d) * args.dfStride + eg);
                const double c1 = __ldg(args.df + (size_t)(3 + d) * args.dfStride + eg);
                const double c2 = __ldg(args.df + (size_t)(6 + d) * args.dfStride + eg);
                double g[NQ][NQ];
#pragma unroll
                for (int j = 0; j < NQ; ++j)
#pragma unroll
                    for (int i = 0; i < NQ; ++i)
                    {
                        const int a = so + j * NQ + i;
                        g[j][i]     = (c0 * sF0[a] + c1 * sF1[a] + c2 * sF2[a]) * ((jwk * tab.w[j]) * tab.w[i]);
                    }
                // matrices of this pass: direction 0 uses dB along i, direction 1 dB along j
#pragma unroll
                for (int p = 0; p < NM; ++p)
                {
                    double t[NQ];
#pragma unroll
                    for (int j = 0; j < NQ; ++j)
                    {
                        double s = (d == 0 ? dtab.dB[p * NQ] : tab.B[p * NQ]) * g[j][0];
#pragma unroll
                        for (int i = 1; i < NQ; ++i) s = fma(d == 0 ? dtab.dB[p * NQ + i] : tab.B[p * NQ + i], g[j][i], s);
                        t[j] = s;
                    }
#pragma unroll
                    for (int q = 0; q < NM; ++q)
                    {
                        double s = (d == 1 ? dtab.dB[q * NQ] : tab.B[q * NQ]) * t[0];
#pragma unroll
                        for (int j = 1; j < NQ; ++j) s = fma(d == 1 ? dtab.dB[q * NQ + j] : tab.B[q * NQ + j], t[j], s);
                        const int x = q * LS + p * NQ + k1;
                        if (d == 0) XB[x] = s;
                        else if (d == 1) XB[x] += s;
                        else XD[x] = s;
                    }
                }
                asm volatile("" ::: "memory"); // keep the three passes apart: one slab of registers at a time
            }
        }
        __syncwarp();
        // the input buffers are consumed: request the next batch; the staging block must be free of the previous
        // batch's bulk store
        if (wbnext < nWB) issue(wbnext);
        tma_store_wait_read0();
        __syncwarp();

        // ------------------------------------------------------------------ stage II: lane (e, q)
        if (lane < EPW * NM && e2 < ne)
        {
            const double *XB = sXB + e2 * EX + q2 * LS, *XD = sXD + e2 * EX + q2 * LS;
            double *O = sOut + e2 * CE;
#pragma unroll
            for (int p = 0; p < NM; ++p)
            {
                double vb[NQ], vd[NQ];
#pragma unroll
                for (int k = 0; k < NQ; ++k)
                {
                    vb[k] = XB[p * NQ + k];
                    vd[k] = XD[p * NQ + k];
                }
#pragma unroll
                for (int r = 0; r < NM; ++r)
                {
                    double s = tab.B[r * NQ] * vb[0];
#pragma unroll
                    for (int k = 1; k < NQ; ++k) s = fma(tab.B[r * NQ + k], vb[k], s);
#pragma unroll
                    for (int k = 0; k < NQ; ++k) s = fma(dtab.dB[r * NQ + k], vd[k], s);
                    O[r * CS + q2 * NM + p] = s;
                }
            }
        }
        double *dst = args.out + (size_t)wb * EPW * NM3;
        if (tout)
        {
            if (CPAD)
            {
                __syncwarp();
                for (int i2 = lane; i2 < ne * NM3 / 2; i2 += 32)
                    *reinterpret_cast<double2 *>(dst + 2 * i2) = *reinterpret_cast<const double2 *>(sOut + pad_out(i2));
                __syncwarp();
            }
            else
            {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) tma_store_1d(dst, sOut, (uint32_t)(ne * NM3 * 8));
                tma_store_commit();
            }
        }
        else
        {
            __syncwarp();
            for (int i = lane; i < ne * NM3; i += 32)
            {
                const int e = i / NM3, w = i - e * NM3, s = w / NM2;
                dst[i] = sOut[e * CE + s * CS + (w - s * NM2)];
            }
            __syncwarp();
        }
    }
    tma_store_wait0();
}

} // namespace nekmf
