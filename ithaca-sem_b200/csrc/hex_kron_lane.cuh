// hex_kron_lane.cuh -- coefficient-space Helmholtz for regular hexahedra with a diagonal metric at nm = 2, 3, 4:
// ONE LANE PER ELEMENT.  Included by hex_kron.cu.
//
// Same operator as hex_helm_kron_kernel (out = J [ lam MMM + G00 MMK + G11 MKM + G22 KMM ] in).  At these orders an
// element (8, 27, 64 coefficients) fits the registers of a single lane, so the whole triple contraction happens
// there, one output column p' at a time, with no exchange between lanes and no barrier -- the scheme of
// quad_kron.cu.  A warp is an independent worker with two 32-element buffers (one at nm = 4, where eight warps with a single
// buffer beat four with two); results are written back into the
// lane's own slot and leave from there.  Even-sized elements (nm = 2, 4) sit in slots padded by two doubles that
// are filled by warp-wide 16-byte cp.async copies and drained by warp-wide 16-byte stores; nm = 3 (27 doubles, odd
// stride, conflict free) travels as one bulk TMA copy per batch each way.
#pragma once

namespace nekmf
{

template <int NM> struct KronLaneCfg
{
    static constexpr int NM2 = NM * NM, NM3 = NM2 * NM;
    static constexpr bool PADDED  = (NM3 % 2) == 0;
    static constexpr int ES       = PADDED ? NM3 + 2 : NM3;
    static constexpr int BUF      = round_up(32 * ES, 2);
    static constexpr int GEO      = 32 * 4;
    // two buffers (the next batch lands while this one is computed) as long as >= 8 warps still fit; at nm = 4 a
    // single buffer and three times the warps hide the latency instead
    static constexpr bool DBUF    = (2 * BUF + 2 * GEO + 2) * 8 * 8 <= 200 * 1024;
    static constexpr int PER_WARP = (DBUF ? 2 : 1) * (BUF + GEO) + 2;
    static constexpr int W_FIT    = (200 * 1024) / (PER_WARP * 8);
    static constexpr int WARPS    = W_FIT >= 16 ? 16 : (W_FIT >= 12 ? 12 : (W_FIT >= 8 ? 8 : (W_FIT >= 4 ? 4 : 1)));
    static constexpr int T        = WARPS * 32;
    static constexpr size_t SMEM  = (size_t)WARPS * PER_WARP * 8 + 16;
};

__device__ __forceinline__ void lane_cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void lane_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void lane_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NM, bool SPARSEK>
__global__ void __launch_bounds__(KronLaneCfg<NM>::T, 1)
    hex_helm_kronlane_kernel(const __grid_constant__ KronTab<NM> tab, const __grid_constant__ KronArgs args)
{
    using Cfg = KronLaneCfg<NM>;
    constexpr int NM2 = Cfg::NM2, NM3 = Cfg::NM3, ES = Cfg::ES, BUF = Cfg::BUF;
    constexpr bool PADDED = Cfg::PADDED;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *wbase  = reinterpret_cast<double *>(smem_raw) + (size_t)warp * Cfg::PER_WARP;
    constexpr bool DBUF = Cfg::DBUF;
    constexpr int NB    = DBUF ? 2 : 1;
    double *sBuf   = wbase;                // [NB][BUF]
    double *sGeo   = wbase + NB * BUF;     // [NB][GEO]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sGeo + NB * Cfg::GEO); // [2]
#define LM(a, b) tab.Ms[tri(a, b, NM)]
#define LK(a, b) tab.Ks[tri(a, b, NM)]
#define LNZ(a, b) (!SPARSEK || (a) == (b) || ((a) < 2 && (b) < 2))

    const int nElmt = args.nElmt;
    const int nB    = (nElmt + 31) / 32;
    const int GW    = gridDim.x * Cfg::WARPS;
    const int gw    = blockIdx.x * Cfg::WARPS + warp;
    if (lane == 0)
    {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto batch_ne = [&](int b) { int r = nElmt - b * 32; return r < 32 ? r : 32; };
    auto fast_ok  = [&](int b) { return args.io_aligned && (PADDED || ((batch_ne(b) * NM3) & 1) == 0); };
    // padded slot address of double pair i2 of a batch
    auto padded = [&](int i2) { const int e = (2 * i2) / NM3; return e * ES + (2 * i2 - e * NM3); };
    auto issue  = [&](int b, int s) { // whole warp: geometry by TMA (mbarrier s), coefficients by cp.async or TMA
        const int ne    = batch_ne(b);
        const bool fast = fast_ok(b);
        double *dst     = sBuf + s * BUF;
        const double *g = args.in + (size_t)b * 32 * NM3;
        if (lane == 0)
        {
            mbar_expect_tx(&bars[s], (uint32_t)(ne * 32 + ((fast && !PADDED) ? ne * NM3 * 8 : 0)));
            tma_load_1d(sGeo + s * Cfg::GEO, args.geo4 + (size_t)b * 32 * 4, (uint32_t)(ne * 32), &bars[s]);
            if (fast && !PADDED) tma_load_1d(dst, g, (uint32_t)(ne * NM3 * 8), &bars[s]);
        }
        if (fast && PADDED)
            for (int i2 = lane; i2 < ne * NM3 / 2; i2 += 32) lane_cp_async16(dst + padded(i2), g + 2 * i2);
        if (PADDED) lane_cp_async_commit(); // one group per issue, even when empty: keeps the wait counts uniform
    };

    uint32_t phase[2] = {0u, 0u};
    if (DBUF && gw < nB) issue(gw, 0);
    int it = 0;
    for (int b = gw; b < nB; b += GW, ++it)
    {
        const int s = DBUF ? (it & 1) : 0, ne = batch_ne(b), bnext = b + GW;
        const bool fast = fast_ok(b);
        double *buf     = sBuf + s * BUF;
        if (!DBUF)
        {
            // single buffer: it was drained at the end of the previous iteration
            tma_store_wait_read0();
            __syncwarp();
            issue(b, 0);
            if (PADDED) lane_cp_async_wait<0>();
        }
        else if (bnext < nB)
        {
            // the other buffer was drained by the previous iteration (bulk store: wait until it has been read)
            tma_store_wait_read0();
            __syncwarp();
            issue(bnext, s ^ 1);
            if (PADDED) lane_cp_async_wait<1>(); // everything but the group just issued has landed
        }
        else if (PADDED)
            lane_cp_async_wait<0>();
        if (!fast)
        {
            const double *g = args.in + (size_t)b * 32 * NM3;
            for (int i = lane; i < ne * NM3; i += 32) buf[(i / NM3) * ES + (i % NM3)] = __ldg(g + i);
        }
        mbar_wait(&bars[s], phase[s]);
        phase[s] ^= 1;
        __syncwarp();

        if (lane < ne)
        {
            double *xe        = buf + lane * ES;
            const double *geo = sGeo + s * Cfg::GEO + lane * 4;
            const double lamJ = args.lambda * geo[0], jg00 = geo[1], jg11 = geo[2], jg22 = geo[3];
            double x[NM][NM][NM]; // x[r][q][p]
#pragma unroll
            for (int r = 0; r < NM; ++r)
#pragma unroll
                for (int q = 0; q < NM; ++q)
#pragma unroll
                    for (int p = 0; p < NM; ++p) x[r][q][p] = xe[(r * NM + q) * NM + p];
#pragma unroll
            for (int pp = 0; pp < NM; ++pp)
            {
                // p-contraction for this output column:  u = lamJ a_M + g00 a_K,  v = a_M   (both [r][q])
                double u[NM][NM], v[NM][NM];
#pragma unroll
                for (int r = 0; r < NM; ++r)
#pragma unroll
                    for (int q = 0; q < NM; ++q)
                    {
                        double m = LM(pp, 0) * x[r][q][0], k = 0.0;
                        bool kset = false;
#pragma unroll
                        for (int p = 1; p < NM; ++p) m = fma(LM(pp, p), x[r][q][p], m);
#pragma unroll
                        for (int p = 0; p < NM; ++p)
                            if (LNZ(pp, p))
                            {
                                k    = kset ? fma(LK(pp, p), x[r][q][p], k) : LK(pp, p) * x[r][q][p];
                                kset = true;
                            }
                        u[r][q] = fma(lamJ, m, jg00 * k);
                        v[r][q] = m;
                    }
#pragma unroll
                for (int qq = 0; qq < NM; ++qq)
                {
                    // q-contraction:  b1[r] goes through M_r, b2[r] through K_r
                    double b1[NM], b2[NM];
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                    {
                        double s1 = LM(qq, 0) * u[r][0], s2 = LM(qq, 0) * v[r][0];
#pragma unroll
                        for (int q = 1; q < NM; ++q)
                        {
                            s1 = fma(LM(qq, q), u[r][q], s1);
                            s2 = fma(LM(qq, q), v[r][q], s2);
                        }
                        double s3 = 0.0;
                        bool set3 = false;
#pragma unroll
                        for (int q = 0; q < NM; ++q)
                            if (LNZ(qq, q))
                            {
                                s3   = set3 ? fma(LK(qq, q), v[r][q], s3) : LK(qq, q) * v[r][q];
                                set3 = true;
                            }
                        b1[r] = fma(jg11, s3, s1);
                        b2[r] = jg22 * s2;
                    }
#pragma unroll
                    for (int rr = 0; rr < NM; ++rr)
                    {
                        double o = LM(rr, 0) * b1[0];
#pragma unroll
                        for (int r = 1; r < NM; ++r) o = fma(LM(rr, r), b1[r], o);
#pragma unroll
                        for (int r = 0; r < NM; ++r)
                            if (LNZ(rr, r)) o = fma(LK(rr, r), b2[r], o);
                        xe[(rr * NM + qq) * NM + pp] = o; // the lane's own slot; x is in registers
                    }
                }
            }
        }
        double *dstg = args.out + (size_t)b * 32 * NM3;
        if (fast && PADDED)
        {
            __syncwarp();
            for (int i2 = lane; i2 < ne * NM3 / 2; i2 += 32)
                *reinterpret_cast<double2 *>(dstg + 2 * i2) = *reinterpret_cast<const double2 *>(buf + padded(i2));
            __syncwarp();
        }
        else if (fast)
        {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) tma_store_1d(dstg, buf, (uint32_t)(ne * NM3 * 8));
            tma_store_commit();
        }
        else
        {
            __syncwarp();
            for (int i = lane; i < ne * NM3; i += 32) dstg[i] = buf[(i / NM3) * ES + (i % NM3)];
            __syncwarp();
        }
    }
    tma_store_wait0();
#undef LM
#undef LK
#undef LNZ
}

} // namespace nekmf
