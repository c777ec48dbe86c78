// hex_kron_rows.cuh -- coefficient-space Helmholtz for regular hexahedra with a DIAGONAL metric at nm = 7..10,
// where the accumulator blocks of hex_helm_kron_kernel (3 nm^2 doubles per lane) no longer fit the register file.
// Included by hex_kron.cu.
//
// Same operator (hex_kron.cu header):  out = J [ lam MMM + G00 MMK + G11 MKM + G22 KMM ] in.
// Row streaming instead of accumulation: lane (e,r) produces ONE output row q' of its slab at a time,
//     U_M[q'][.] = sum_q M[q'][q] (lamJ a_M + g00 a_K)_q + K[q'][q] (g11 a_M)_q ,   U_K[q'][.] = g22 sum_q M[q'][q] (a_M)_q
// with the 1-D products a_M = M x_q, a_K = K x_q recomputed for the few rows q that M[q'][.] and K[q'][.] touch
// (the modified C0 basis makes M penta-diagonal-like and K diagonal + vertex block: <= 4 rows), and writes it to
// two exchange blocks at once.  Lane (e,p') then contracts r for one q' at a time from both blocks and writes the
// result row to a separate staging block.  Registers: a few rows, no nm^2 arrays.
#pragma once

namespace nekmf
{

template <int NM> struct KronRowsCfg
{
    static constexpr int NM2 = NM * NM, NM3 = NM2 * NM;
    static constexpr int EPW = 32 / NM;
    static constexpr int INB = round_up(EPW * NM3, 2);
    static constexpr int PS  = kron_pad(NM2, 1);
    static constexpr int ES  = kron_pad(NM * PS, NM);
    static constexpr int XB  = round_up(EPW * ES, 2);
    static constexpr int GEO = EPW * 4;
    static constexpr int PER_WARP = 2 * INB + 2 * XB + GEO + 2; // input, staging, two exchange blocks
    static constexpr int W_FIT = (224 * 1024) / (PER_WARP * 8);
    static constexpr int WARPS = W_FIT >= 8 ? 8 : (W_FIT >= 4 ? 4 : (W_FIT >= 1 ? W_FIT : 1));
    static constexpr int T     = WARPS * 32;
    static constexpr size_t SMEM = (size_t)WARPS * PER_WARP * 8 + 16;
};

// sparsity predicates of the modified C0 basis (verified numerically at creation: KronState::sparse_full)
__host__ __device__ constexpr bool rows_mnz(int a, int b)
{
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return hi < 2 || (lo < 2 && hi <= 3) || (lo >= 2 && (hi - lo) % 2 == 0 && hi - lo <= 2);
}
__host__ __device__ constexpr bool rows_knz(int a, int b) { return a == b || (a < 2 && b < 2); }

template <int NM>
__global__ void __launch_bounds__(KronRowsCfg<NM>::T, 1)
    hex_helm_kronrows_kernel(const __grid_constant__ KronTab<NM> tab, const __grid_constant__ KronArgs args)
{
    using Cfg = KronRowsCfg<NM>;
    constexpr int NM2 = Cfg::NM2, NM3 = Cfg::NM3, EPW = Cfg::EPW, INB = Cfg::INB, PS = Cfg::PS, ES = Cfg::ES, XB = Cfg::XB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *wbase = reinterpret_cast<double *>(smem_raw) + (size_t)warp * Cfg::PER_WARP;
    double *sIn   = wbase;            // [INB] input block
    double *sOut  = sIn + INB;        // [INB] output staging
    double *sX1   = sOut + INB;       // [XB]  U_M, goes through M_r
    double *sX2   = sX1 + XB;         // [XB]  U_K, goes through K_r
    double *sGeo  = sX2 + XB;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sGeo + Cfg::GEO);

    const int nElmt = args.nElmt;
    const int nWB   = (nElmt + EPW - 1) / EPW;
    const int GW    = gridDim.x * Cfg::WARPS;
    const int gw    = blockIdx.x * Cfg::WARPS + warp;
    const int e     = lane / NM;
    const int s1    = lane - e * NM;
    const bool active = lane < EPW * NM;
#define RM(a, b) tab.Ms[tri(a, b, NM)]
#define RK(a, b) tab.Ks[tri(a, b, NM)]

    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto batch_ne = [&](int wb) { int r = nElmt - wb * EPW; return r < EPW ? r : EPW; };
    auto tma_ok   = [&](int wb) { return args.io_aligned && ((batch_ne(wb) * NM3) & 1) == 0 && (((wb * EPW * NM3) & 1) == 0); };
    auto issue    = [&](int wb) { // lane 0; sIn and sGeo are free
        const int ne   = batch_ne(wb);
        uint32_t bytes = (uint32_t)(ne * 32);
        if (tma_ok(wb)) bytes += (uint32_t)(ne * NM3 * 8);
        mbar_expect_tx(bar, bytes);
        tma_load_1d(sGeo, args.geo4 + (size_t)wb * EPW * 4, (uint32_t)(ne * 32), bar);
        if (tma_ok(wb)) tma_load_1d(sIn, args.in + (size_t)wb * EPW * NM3, (uint32_t)(ne * NM3 * 8), bar);
    };

    uint32_t phase = 0;
    if (lane == 0 && gw < nWB) issue(gw);

    for (int wb = gw; wb < nWB; wb += GW)
    {
        const int ne      = batch_ne(wb);
        const int wbnext  = wb + GW;
        const bool tma_in = tma_ok(wb);
        if (!tma_in)
        {
            const double *src = args.in + (size_t)wb * EPW * NM3;
            for (int i = lane; i < ne * NM3; i += 32) sIn[i] = __ldg(src + i);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        __syncwarp();

        const double *g   = sGeo + (e < ne ? e : 0) * 4;
        const double lamJ = args.lambda * g[0], jg00 = g[1], jg11 = g[2], jg22 = g[3];

        // ---- stage I: lane (e, r) streams the output rows q' of its slab into the two exchange blocks
        if (active)
        {
            const double *xin = sIn + e * NM3 + s1 * NM2;
#pragma unroll
            for (int qq = 0; qq < NM; ++qq)
            {
                double um[NM], uk[NM];
#pragma unroll
                for (int pp = 0; pp < NM; ++pp) um[pp] = uk[pp] = 0.0;
#pragma unroll
                for (int q = 0; q < NM; ++q)
                {
                    if (!rows_mnz(qq, q) && !rows_knz(qq, q)) continue;
                    double xr[NM];
#pragma unroll
                    for (int p = 0; p < NM; ++p) xr[p] = xin[q * NM + p];
#pragma unroll
                    for (int pp = 0; pp < NM; ++pp)
                    {
                        double m = 0.0, k = 0.0;
                        bool mset = false, kset = false;
#pragma unroll
                        for (int p = 0; p < NM; ++p)
                        {
                            if (rows_mnz(pp, p)) { m = mset ? fma(RM(pp, p), xr[p], m) : RM(pp, p) * xr[p]; mset = true; }
                            if (rows_mnz(qq, q) && rows_knz(pp, p)) { k = kset ? fma(RK(pp, p), xr[p], k) : RK(pp, p) * xr[p]; kset = true; }
                        }
                        if (rows_mnz(qq, q))
                        {
                            um[pp] = fma(RM(qq, q), fma(lamJ, m, jg00 * k), um[pp]);
                            uk[pp] = fma(RM(qq, q), jg22 * m, uk[pp]);
                        }
                        if (rows_knz(qq, q)) um[pp] = fma(RK(qq, q), jg11 * m, um[pp]);
                    }
                }
#pragma unroll
                for (int pp = 0; pp < NM; ++pp)
                {
                    sX1[e * ES + pp * PS + qq * NM + s1] = um[pp];
                    sX2[e * ES + pp * PS + qq * NM + s1] = uk[pp];
                }
            }
        }
        __syncwarp();
        // sIn / sGeo are consumed: request the next batch; the staging block must be free of the previous bulk store
        if (lane == 0)
        {
            tma_store_wait_read0();
            if (wbnext < nWB) issue(wbnext);
        }
        __syncwarp();
        // ---- stage II: lane (e, p') contracts r for one q' at a time
        if (active)
        {
            const double *v1 = sX1 + e * ES + s1 * PS, *v2 = sX2 + e * ES + s1 * PS;
#pragma unroll
            for (int qq = 0; qq < NM; ++qq)
            {
                double c1[NM], c2[NM];
#pragma unroll
                for (int r = 0; r < NM; ++r)
                {
                    c1[r] = v1[qq * NM + r];
                    c2[r] = v2[qq * NM + r];
                }
#pragma unroll
                for (int rr = 0; rr < NM; ++rr)
                {
                    double o = 0.0;
                    bool oset = false;
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                    {
                        if (rows_mnz(rr, r)) { o = oset ? fma(RM(rr, r), c1[r], o) : RM(rr, r) * c1[r]; oset = true; }
                        if (rows_knz(rr, r)) { o = oset ? fma(RK(rr, r), c2[r], o) : RK(rr, r) * c2[r]; oset = true; }
                    }
                    sOut[e * NM3 + rr * NM2 + qq * NM + s1] = o;
                }
            }
        }
        if (tma_in)
        {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0)
            {
                tma_store_1d(args.out + (size_t)wb * EPW * NM3, sOut, (uint32_t)(ne * NM3 * 8));
                tma_store_commit();
            }
        }
        else
        {
            __syncwarp();
            double *dst = args.out + (size_t)wb * EPW * NM3;
            for (int i = lane; i < ne * NM3; i += 32) dst[i] = sOut[i];
        }
        __syncwarp();
    }
    if (lane == 0) tma_store_wait0();
#undef RM
#undef RK
}

} // namespace nekmf
