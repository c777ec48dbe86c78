// hex_kron_fullrows.cuh -- coefficient-space Helmholtz for REGULAR (affine) hexahedra with a FULL constant Laplacian
// metric (sheared / rotated parallelepipeds) at nm = 7..10, where the two nm^2 accumulator blocks plus the nm^2
// stage-II block of hex_helm_kronfull_kernel no longer fit the register file.  Included by hex_kron.cu.
//
// Reference semantics: MatrixFreeOps/Helmholtz.h:764-993 (DEFORMED=false).  Same operator as hex_kron_full.cuh:
//   out = J [ lam MMM + G00 MMK + G11 MKM + G22 KMM
//             + G01 M (x) (S_pp' S_q'q + S_p'p S_qq') + G02 (S_pp' S_r'r + S_p'p S_rr') (x) M_q
//             + G12 (S_qq' S_r'r + S_q'q S_rr') (x) M_p ] in
// (factors listed r (x) q (x) p), M = B W B^T, K = (DB) W (DB)^T, S[a][b] = sum_i w_i (DB)_a(i) B_b(i).
//
// The four slab-level intermediates are produced ONE AT A TIME,
//   U_M (goes through M_r), U_K (through K_r), U_T (through sum_r S[r'][r]), U_S (through sum_r S[r][r']),
// each in chunks of QC output rows q' so that a lane holds QC x nm doubles of it, handed to lane (e,p') through the
// one exchange block, contracted along r there and accumulated in the shared-memory staging block that finally
// leaves by a bulk TMA store (lane-private addresses: read-modify-write without any synchronisation).  The 1-D row
// products a_M, a_K, a_S, a_T are recomputed per intermediate for the rows its sparsity pattern touches: about
// 3300 FMA per (element, slab) at nm = 7 instead of 2400 with everything in registers, against 6200 per slab for
// the quadrature-space kernel -- and the only HBM traffic is the coefficient block in and out.
// The modified-basis sparsity of M, K and S is compiled in (verified at creation, else the quadrature-space
// kernel stays).
#pragma once

namespace nekmf
{

template <int NM> struct KronFullRowsCfg
{
    static constexpr int NM2 = NM * NM, NM3 = NM2 * NM;
    static constexpr int EPW = 32 / NM;
    static constexpr int INB = round_up(EPW * NM3, 2);
    static constexpr int PS  = kron_pad(NM2, 1);
    static constexpr int ES  = kron_pad(NM * PS, NM);
    static constexpr int XB  = round_up(EPW * ES, 2);
    static constexpr int GEO = EPW * 8;
    static constexpr int PER_WARP = 2 * INB + XB + GEO + 2; // input, accumulator / output staging, exchange block
    static constexpr int W_FIT = (224 * 1024) / (PER_WARP * 8);
    static constexpr int WARPS = W_FIT >= 8 ? 8 : (W_FIT >= 6 ? 6 : (W_FIT >= 4 ? 4 : (W_FIT >= 1 ? W_FIT : 1)));
    static constexpr int T     = WARPS * 32;
    static constexpr size_t SMEM = (size_t)WARPS * PER_WARP * 8 + 16;
    // output rows q' per stage-I chunk: QC * NM doubles of the intermediate live in registers
    static constexpr int QC  = NM <= 7 ? NM : (NM + 1) / 2;
    static constexpr int NCH = (NM + QC - 1) / QC;
};

// sparsity predicates of the modified C0 basis (verified numerically at creation)
__host__ __device__ constexpr bool fr_mnz(int a, int b) { return rows_mnz(a, b); }
__host__ __device__ constexpr bool fr_knz(int a, int b) { return rows_knz(a, b); }
__host__ __device__ constexpr bool fr_snz(int a, int b)
{
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return hi < 2 || (lo < 2 && hi == 2) || (lo >= 2 && hi - lo == 1);
}
// does input row q contribute to output row qq of intermediate TYPE (0: U_M, 1: U_K, 2: U_T, 3: U_S)?
// diag: diagonal metric (G01 = G02 = G12 = 0), no S terms and only the intermediates 0 and 1
__host__ __device__ constexpr bool fr_touch(int type, int qq, int q, bool diag)
{
    return type == 0 ? (fr_mnz(qq, q) || fr_knz(qq, q) || (!diag && fr_snz(qq, q))) : (type == 1 ? fr_mnz(qq, q) : (fr_mnz(qq, q) || fr_snz(qq, q)));
}
// is q the first input row that touches output row qq?
__host__ __device__ constexpr bool fr_first_touch(int type, int qq, int q, bool diag)
{
    bool earlier = false;
    for (int k = 0; k < q; ++k) earlier = earlier || fr_touch(type, qq, k, diag);
    return !earlier;
}
__host__ __device__ constexpr bool fr_touch_chunk(int type, int q0, int q1, int q, bool diag)
{
    bool t = false;
    for (int qq = q0; qq < q1; ++qq) t = t || fr_touch(type, qq, q, diag);
    return t;
}

// compile-time loops: the sparsity predicates must fold, whatever the unroller's size heuristics say
template <int I> struct fr_int { static constexpr int value = I; };
template <int I, int N, class F> __device__ __forceinline__ void fr_static_for(F &&f)
{
    if constexpr (I < N)
    {
        f(fr_int<I>{});
        fr_static_for<I + 1, N>(f);
    }
}

// DIAG: the same kernel for a diagonal metric (axis-aligned elements) -- the S terms and the intermediates U_T, U_S
// are compiled out
template <int NM, bool DIAG>
__global__ void __launch_bounds__(KronFullRowsCfg<NM>::T, 1)
    hex_helm_kronfullrows_kernel(const __grid_constant__ KronFullTab<NM> tab, const __grid_constant__ KronFullArgs args)
{
    using Cfg = KronFullRowsCfg<NM>;
    constexpr int NM2 = Cfg::NM2, NM3 = Cfg::NM3, EPW = Cfg::EPW, INB = Cfg::INB, PS = Cfg::PS, ES = Cfg::ES;
    constexpr int QC = Cfg::QC, NCH = Cfg::NCH;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *wbase = reinterpret_cast<double *>(smem_raw) + (size_t)warp * Cfg::PER_WARP;
    double *sIn   = wbase;            // [INB] input block
    double *sAcc  = sIn + INB;        // [INB] accumulator of stage II, then the source of the bulk store
    double *sX    = sAcc + INB;       // [XB]  exchange block of the current intermediate
    double *sGeo  = sX + Cfg::XB;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sGeo + Cfg::GEO);

    const int nElmt = args.nElmt;
    const int nWB   = (nElmt + EPW - 1) / EPW;
    const int GW    = gridDim.x * Cfg::WARPS;
    const int gw    = blockIdx.x * Cfg::WARPS + warp;
    const int e     = lane / NM;
    const int s1    = lane - e * NM; // r in stage I, p' in stage II
    const bool active = lane < EPW * NM;
#define QM(a, b) tab.Ms[tri(a, b, NM)]
#define QK(a, b) tab.Ks[tri(a, b, NM)]
#define QS(a, b) tab.S[(a) * NM + (b)]
    // acc (+)= coef * val, the first contribution initialises (flags fold at compile time after unrolling)
#define QACC(flag, accv, coef, val) { accv = flag ? fma(coef, val, accv) : (coef) * (val); flag = true; }

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
        uint32_t bytes = (uint32_t)(ne * 64);
        if (tma_ok(wb)) bytes += (uint32_t)(ne * NM3 * 8);
        mbar_expect_tx(bar, bytes);
        tma_load_1d(sGeo, args.geo8 + (size_t)wb * EPW * 8, (uint32_t)(ne * 64), bar);
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

        const double *g   = sGeo + (e < ne ? e : 0) * 8;
        const double lamJ = args.lambda * g[0], g00 = g[1], g11 = g[2], g22 = g[3], g01 = g[4], g02 = g[5], g12 = g[6];
        const double *xin = sIn + e * NM3 + s1 * NM2;

        fr_static_for<0, (DIAG ? 2 : 4)>([&](auto type_c) {
            constexpr int type = decltype(type_c)::value;
            // ---- stage I: lane (e, r) builds the intermediate of this type, QC rows q' at a time
            if (active)
            {
                fr_static_for<0, NCH>([&](auto ch_c) {
                    constexpr int q0 = decltype(ch_c)::value * QC, q1 = (q0 + QC < NM) ? q0 + QC : NM;
                    double U[QC][NM];
                    fr_static_for<0, NM>([&](auto q_c) {
                        constexpr int q = decltype(q_c)::value;
                        if constexpr (fr_touch_chunk(type, q0, q1, q, DIAG))
                        {
                        double xr[NM];
#pragma unroll
                        for (int p = 0; p < NM; ++p) xr[p] = xin[q * NM + p];
                        // 1-D products along p: aM = M x, aK = K x, aS[p'] = sum_p S[p][p'] x[p], aT[p'] = sum_p S[p'][p] x[p]
                        double v1[NM], v2[NM], v3[NM], v4[NM];
#pragma unroll
                        for (int pp = 0; pp < NM; ++pp)
                        {
                            double m = 0.0, s = 0.0, t = 0.0, k = 0.0;
                            bool mset = false, sset = false, tset = false, kset = false;
#pragma unroll
                            for (int p = 0; p < NM; ++p)
                            {
                                if (fr_mnz(pp, p)) QACC(mset, m, QM(pp, p), xr[p])
                                if (!DIAG && (type == 0 || type == 2) && fr_snz(p, pp)) QACC(sset, s, QS(p, pp), xr[p])
                                if (!DIAG && (type == 0 || type == 3) && fr_snz(pp, p)) QACC(tset, t, QS(pp, p), xr[p])
                                if (type == 0 && fr_knz(pp, p)) QACC(kset, k, QK(pp, p), xr[p])
                            }
                            if (type == 0)
                            {
                                v1[pp] = fma(lamJ, m, g00 * k); // with M_q
                                v2[pp] = g11 * m;               // with K_q
                                v3[pp] = g01 * s;               // with sum_q S[q'][q]
                                v4[pp] = g01 * t;               // with sum_q S[q][q']
                            }
                            else if (type == 1) v1[pp] = g22 * m; // with M_q
                            else if (type == 2)
                            {
                                v1[pp] = g02 * s; // with M_q
                                v4[pp] = g12 * m; // with sum_q S[q][q']
                            }
                            else
                            {
                                v1[pp] = g02 * t; // with M_q
                                v3[pp] = g12 * m; // with sum_q S[q'][q]
                            }
                        }
                        fr_static_for<0, QC>([&](auto a_c) {
                            constexpr int a = decltype(a_c)::value, qq = q0 + a;
                            if constexpr (qq < q1 && fr_touch(type, qq < NM ? qq : 0, q, DIAG))
                            {
                                // the first input row that touches output row qq initialises it
                                constexpr bool first = fr_first_touch(type, qq, q, DIAG);
#pragma unroll
                                for (int pp = 0; pp < NM; ++pp)
                                {
                                    double u  = first ? 0.0 : U[a][pp];
                                    bool have = !first;
                                    if (fr_mnz(qq, q)) QACC(have, u, QM(qq, q), v1[pp])
                                    if (type == 0 && fr_knz(qq, q)) QACC(have, u, QK(qq, q), v2[pp])
                                    if (!DIAG && (type == 0 || type == 3) && fr_snz(qq, q)) QACC(have, u, QS(qq, q), v3[pp])
                                    if (!DIAG && (type == 0 || type == 2) && fr_snz(q, qq)) QACC(have, u, QS(q, qq), v4[pp])
                                    U[a][pp] = u;
                                }
                            }
                        });
                        }
                    });
                    fr_static_for<0, QC>([&](auto a_c) {
                        constexpr int a = decltype(a_c)::value, qq = q0 + a;
                        if constexpr (qq < q1)
                        {
#pragma unroll
                            for (int pp = 0; pp < NM; ++pp) sX[e * ES + pp * PS + qq * NM + s1] = U[a][pp];
                        }
                    });
                });
            }
            __syncwarp();
            if (type == 0)
            {
                // the staging block may still be the source of the previous batch's bulk store
                if (lane == 0) tma_store_wait_read0();
                __syncwarp();
            }
            if (type == (DIAG ? 1 : 3))
            {
                // sIn / sGeo are consumed: request the next warp batch now
                if (lane == 0 && wbnext < nWB) issue(wbnext);
                __syncwarp();
            }
            // ---- stage II: lane (e, p') contracts r with the matrix of this type, one q' at a time
            if (active)
            {
                const double *v = sX + e * ES + s1 * PS;
                double *acc     = sAcc + e * NM3 + s1;
#pragma unroll
                for (int qq = 0; qq < NM; ++qq)
                {
                    double col[NM];
#pragma unroll
                    for (int r = 0; r < NM; ++r) col[r] = v[qq * NM + r];
#pragma unroll
                    for (int rr = 0; rr < NM; ++rr)
                    {
                        double o  = 0.0;
                        bool oset = false;
                        if (type != 0)
                        {
                            o    = acc[rr * NM2 + qq * NM];
                            oset = true;
                        }
#pragma unroll
                        for (int r = 0; r < NM; ++r)
                        {
                            if (type == 0 && fr_mnz(rr, r)) QACC(oset, o, QM(rr, r), col[r])
                            if (type == 1 && fr_knz(rr, r)) QACC(oset, o, QK(rr, r), col[r])
                            if (type == 2 && fr_snz(rr, r)) QACC(oset, o, QS(rr, r), col[r])
                            if (type == 3 && fr_snz(r, rr)) QACC(oset, o, QS(r, rr), col[r])
                        }
                        acc[rr * NM2 + qq * NM] = o;
                    }
                }
            }
            __syncwarp(); // the exchange block is free for the next intermediate
        });
        if (tma_in)
        {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0)
            {
                tma_store_1d(args.out + (size_t)wb * EPW * NM3, sAcc, (uint32_t)(ne * NM3 * 8));
                tma_store_commit();
            }
        }
        else
        {
            double *dst = args.out + (size_t)wb * EPW * NM3;
            for (int i = lane; i < ne * NM3; i += 32) dst[i] = sAcc[i];
        }
        __syncwarp();
    }
    if (lane == 0) tma_store_wait0();
#undef QM
#undef QK
#undef QS
#undef QACC
}

} // namespace nekmf
