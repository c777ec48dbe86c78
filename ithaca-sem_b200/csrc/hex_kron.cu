// hex_kron.cu -- Helmholtz on REGULAR (affine) hexahedra whose Laplacian metric is diagonal
// (axis-aligned boxes: every structured mesh), evaluated entirely in coefficient space.
//
// Reference semantics: MatrixFreeOps/Helmholtz.h:764-993 (HelmholtzHexImpl, DEFORMED=false).
// For constant geometric factors the reference's chain
//     BwdTrans -> lambda*IProduct -> PhysDerivTensor -> G (3x3, constant) -> 3x IProduct(dbdata)
// is algebraically a sum of Kronecker products of 1-D matrices, because quadrature, basis and
// derivative are all tensor products:
//     out = J [ lambda M(x)M(x)M + G00 M(x)M(x)K + G11 M(x)K(x)M + G22 K(x)M(x)M ] in
// with the nm x nm matrices  M = B W B^T  (1-D mass)  and  K = (DB) W (DB)^T  (1-D stiffness)
// built once per operator from exactly the reference's tables (bdata, D, quadrature weights).
// Same numbers up to rounding (checked to 1e-12 in tests/), about 3.4x fewer flops, no
// quadrature-space intermediates: 7 small matrix applications on a 5x5x5 block.
//
// B200 mapping: thread = one (element, r) slab of the coefficient block.  Its 25 values are
// contracted in the p and q directions entirely in registers (matrices are constant-bank
// operands); one transposing exchange through shared memory gives every thread the r-lines of a
// fixed p for the third contraction.  Coefficient blocks stream in with TMA bulk loads
// (double-buffered, mbarrier) and out with TMA bulk stores; per-element geometry is 4 doubles.
#include "hex_kernels.cuh"
#include "op_internal.h"
#include <cmath>
#include <stdlib.h>
#include <string.h>

namespace nekmf
{

// 1-D mass / stiffness matrices, upper triangles only (both are symmetric): Ms[tri(a,b)], a <= b.
// 2 * NM(NM+1)/2 doubles -- 60 uniform registers at NM = 5, so every DFMA takes its matrix entry
// from the uniform register file without reloads.
template <int NM> struct KronTab
{
    double Ms[NM * (NM + 1) / 2];
    double Ks[NM * (NM + 1) / 2];
};
__host__ __device__ constexpr int tri(int a, int b, int n)
{
    return a <= b ? a * n - a * (a - 1) / 2 + (b - a) : b * n - b * (b - 1) / 2 + (a - b);
}

struct KronArgs
{
    const double *in;
    double *out;
    const double *geo4; // [nElmt][4] = J, J*G00, J*G11, J*G22
    int nElmt;
    int io_aligned; // in and out 16-byte aligned
    double lambda;
    // GATHER variant (CG mat-vec): `in` is the GLOBAL vector, the coefficient block of local DOF i is
    // sign[i] * in[map[i]] (AssemblyMapCG::v_GlobalToLocal fused into the operator's load)
    const int *map;
    const double *sign;
    const int *skip; // GATHER: non-zero flag = nothing to do (the solver converged after this launch was enqueued)
    int gather_rows; // gather order: 1 = one (q, r) row of ALL the batch's elements per instruction, 0 = local index order
};

constexpr int kron_pad(int minimum, int residue) // smallest v >= minimum with v % 16 == residue
{
    int v = minimum;
    while (v % 16 != residue % 16) ++v;
    return v;
}

// Every warp is an independent worker: it owns EPW elements per step, its own TMA-fed input
// buffer, its own exchange/staging buffer and its own mbarrier -- no CTA-wide barrier in the loop.
template <int NM, bool GATHER = false, int WSEL = 0> struct KronCfg
{
    static constexpr int NM2 = NM * NM, NM3 = NM2 * NM;
    static constexpr int EPW   = 32 / NM;                // elements per warp step (NM lanes per element)
    static constexpr int INB   = round_up(EPW * NM3, 2); // doubles in the input buffer
    // exchange layout X[e*ES + p*PS + q*NM + r]: writers (lanes = (e,r)) and readers (lanes = (e,p))
    // both hit 16 distinct 8-byte banks per half-warp
    static constexpr int PS = kron_pad(NM2, 1);
    static constexpr int ES = kron_pad(NM * PS, NM);
    static constexpr int XB = round_up(EPW * ES > EPW * NM3 ? EPW * ES : EPW * NM3, 2);
    static constexpr int GEO = EPW * 4;
    static constexpr int MAPB = GATHER ? round_up(EPW * NM3, 4) / 2 : 0; // doubles holding EPW*NM3 ints
    static constexpr int PER_WARP = INB + GEO + XB + 2 + MAPB; // doubles (+2: mbarrier, 16-byte slot)
    // one CTA per SM; warps in multiples of 4 (one FP64 pipe per SM sub-partition)
    static constexpr int W_FIT = (224 * 1024) / (PER_WARP * 8);
    // WSEL != 0: explicit warp count (NM = 5 is instantiated with 8 and 12 for the A/B in tools/prof_helm.py)
    static constexpr int WARPS = WSEL ? WSEL : (W_FIT >= 12 ? 12 : (W_FIT >= 8 ? 8 : 4));
    static constexpr int T     = WARPS * 32;
    static constexpr size_t SMEM = (size_t)WARPS * PER_WARP * 8 + 16;
};

// SPARSEK: the stiffness matrix has the structure of the modified C0 basis -- a 2x2 vertex block
// plus a diagonal (interior modes have orthogonal derivatives) -- verified numerically at creation.
// 4/8-byte asynchronous global -> shared copies (LDGSTS): the gather lands in shared memory without
// passing through registers, so a whole batch of indirect loads is in flight per warp
__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// SPARSEK == 2: the mass matrix is sparse as well (vertex block, vertex x modes {2,3}, interior |a-b| in {0,2}:
// 17 of 25 entries at NM = 5), also verified at creation -- about a quarter of the DFMAs of the kernel go away.
template <int NM, int SPARSEK, bool GATHER = false, int WSEL = 0>
__global__ void __launch_bounds__(KronCfg<NM, GATHER, WSEL>::T, 1)
    hex_helm_kron_kernel(const __grid_constant__ KronTab<NM> tab, const __grid_constant__ KronArgs args)
{
    using Cfg = KronCfg<NM, GATHER, WSEL>;
    constexpr int NM2 = Cfg::NM2, NM3 = Cfg::NM3, EPW = Cfg::EPW, INB = Cfg::INB, PS = Cfg::PS, ES = Cfg::ES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (GATHER && args.skip && *args.skip) return;
    double *wbase = reinterpret_cast<double *>(smem_raw) + (size_t)warp * Cfg::PER_WARP;
    double *sIn   = wbase;                  // [INB]  input block of the current step
    double *sGeo  = wbase + INB;            // [GEO]  per-element scalars
    double *sX    = sGeo + Cfg::GEO;        // [XB]   exchange, then output staging
    uint64_t *bar = reinterpret_cast<uint64_t *>(sX + Cfg::XB);
    int *sMap     = reinterpret_cast<int *>(sX + Cfg::XB + 2); // [EPW*NM3] (GATHER only)

    const int nElmt = args.nElmt;
    const int nWB   = (nElmt + EPW - 1) / EPW;          // warp batches
    const int GW    = gridDim.x * Cfg::WARPS;           // warps in the grid
    const int gw    = blockIdx.x * Cfg::WARPS + warp;
    const int e     = lane / NM;                        // element within the warp batch
    const int s1    = lane - e * NM;                    // r in stage I, p' in stage II
    const bool active = lane < EPW * NM;
#define KM(a, b) tab.Ms[tri(a, b, NM)]
#define KK(a, b) tab.Ks[tri(a, b, NM)]
#define KNZ(a, b) (!SPARSEK || (a) == (b) || ((a) < 2 && (b) < 2))
#define KLO(a, b) ((a) < (b) ? (a) : (b))
#define KHI(a, b) ((a) < (b) ? (b) : (a))
#define MNZ(a, b) (SPARSEK < 2 || KHI(a, b) < 2 || (KLO(a, b) < 2 && KHI(a, b) <= 3) || (KLO(a, b) >= 2 && (KHI(a, b) - KLO(a, b)) % 2 == 0 && KHI(a, b) - KLO(a, b) <= 2))

    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto batch_ne = [&](int wb) { int r = nElmt - wb * EPW; return r < EPW ? r : EPW; };
    auto tma_ok   = [&](int wb) { return !GATHER && args.io_aligned && ((batch_ne(wb) * NM3) & 1) == 0 && (((wb * EPW * NM3) & 1) == 0); };
    auto out_tma_ok = [&](int wb) { return args.io_aligned && ((batch_ne(wb) * NM3) & 1) == 0 && (((wb * EPW * NM3) & 1) == 0); };
    // GATHER: whole warp.  fetch_map(wb): local-to-global indices of batch wb -> sMap (asynchronous, completes at
    // the next cp_async_wait_all)
    auto fetch_map = [&](int wb) {
        const int n    = batch_ne(wb) * NM3;
        const int *src = args.map + (size_t)wb * EPW * NM3; // EPW * NM3 even: 8-byte aligned pairs
        if constexpr ((EPW * NM3) % 2 == 0)
        {
            for (int i = lane; i < n / 2; i += 32) cp_async8(sMap + 2 * i, src + 2 * i);
            if ((n & 1) && lane == 0) cp_async4(sMap + n - 1, src + n - 1);
        }
        else
            for (int i = lane; i < n; i += 32) cp_async4(sMap + i, src + i);
    };
    // gather(wb): sIn[i] <- in[sMap[i]], 8-byte asynchronous copies that land in shared memory without passing through
    // registers (completes at the next cp_async_wait_all).  Measured alternative: the values through registers
    // (__ldg early, st.shared after the exchanges) halves the shared-memory wavefronts but exposes the load latency
    // with 8 warps per SM: 0.75 -> 1.47 ms on 2^20 elements.
    // Order: with consecutive elements of a batch adjacent in the mesh (structured numbering, or any element ordering
    // with locality) the p-lines of one (q, r) row of ALL EPW elements are neighbours in the global vector, so one
    // instruction whose lanes are (element, p) touches 2-3 cache lines instead of the 8-9 that 32 consecutive local
    // indices of one element span -- the L1 tag stage handles one line per cycle and was the limiter of this kernel.
    auto gather = [&](int wb) {
        const int ne = batch_ne(wb);
        if (args.gather_rows)
        {
            if (active && e < ne)
            {
                const int base = e * NM3 + s1;
#pragma unroll 5
                for (int qr = 0; qr < NM2; ++qr) cp_async8(sIn + base + qr * NM, args.in + sMap[base + qr * NM]);
            }
        }
        else
        {
            const int n = ne * NM3;
#pragma unroll 4
            for (int i = lane; i < n; i += 32) cp_async8(sIn + i, args.in + sMap[i]);
        }
    };
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
    if (GATHER && gw < nWB)
    {
        fetch_map(gw);
        cp_async_wait_all();
        __syncwarp();
        gather(gw);
        __syncwarp(); // every lane has read its sMap entries
        if (gw + GW < nWB) fetch_map(gw + GW);
    }

    for (int wb = gw; wb < nWB; wb += GW)
    {
        const int ne      = batch_ne(wb);
        const int wbnext  = wb + GW;
        const bool tma_in = tma_ok(wb);
        if (GATHER)
        {
            cp_async_wait_all(); // this batch's gathered block (and the next batch's indices) have landed
            if (args.sign)
            {
                __syncwarp();
                const double *sg = args.sign + (size_t)wb * EPW * NM3;
                for (int i = lane; i < ne * NM3; i += 32) sIn[i] *= __ldg(sg + i);
            }
        }
        else if (!tma_in)
        {
            // 8-byte aligned caller arrays or an odd-sized tail: plain loads by the warp
            const double *src = args.in + (size_t)wb * EPW * NM3;
            for (int i = lane; i < ne * NM3; i += 32) sIn[i] = __ldg(src + i);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        __syncwarp();

        const double *g   = sGeo + (e < ne ? e : 0) * 4;
        const double lamJ = args.lambda * g[0], jg00 = g[1], jg11 = g[2], jg22 = g[3];

        // ---- stage I: lane (e, r): contract p then q in registers
        //      A2 = (M (x) M) x,  R = jg00 (M (x) K) x + jg11 (K (x) M) x   for the lane's r-slab
        double A2[NM][NM], R[NM][NM];
#pragma unroll
        for (int a = 0; a < NM; ++a)
#pragma unroll
            for (int c = 0; c < NM; ++c) A2[a][c] = R[a][c] = 0.0;
        if (active)
        {
            const double *xin = sIn + e * NM3 + s1 * NM2;
#pragma unroll
            for (int q = 0; q < NM; ++q)
            {
                double xr[NM], am[NM], bk[NM], a11[NM];
#pragma unroll
                for (int p = 0; p < NM; ++p) xr[p] = xin[q * NM + p];
#pragma unroll
                for (int pp = 0; pp < NM; ++pp)
                {
                    double m = 0.0, k = 0.0;
                    bool kset = false, mset = false;
#pragma unroll
                    for (int p = 0; p < NM; ++p)
                        if (MNZ(pp, p))
                        {
                            m    = mset ? fma(KM(pp, p), xr[p], m) : KM(pp, p) * xr[p];
                            mset = true;
                        }
#pragma unroll
                    for (int p = 0; p < NM; ++p)
                        if (KNZ(pp, p))
                        {
                            k    = kset ? fma(KK(pp, p), xr[p], k) : KK(pp, p) * xr[p];
                            kset = true;
                        }
                    am[pp]  = m;
                    bk[pp]  = jg00 * k;
                    a11[pp] = jg11 * m;
                }
#pragma unroll
                for (int qq = 0; qq < NM; ++qq)
#pragma unroll
                    for (int pp = 0; pp < NM; ++pp)
                    {
                        if (MNZ(qq, q))
                        {
                            A2[qq][pp] = fma(KM(qq, q), am[pp], A2[qq][pp]);
                            R[qq][pp]  = fma(KM(qq, q), bk[pp], R[qq][pp]);
                        }
                        if (KNZ(qq, q)) R[qq][pp] = fma(KK(qq, q), a11[pp], R[qq][pp]);
                    }
            }
        }
        __syncwarp();
        // sIn / sGeo are consumed: request the next warp batch now, it lands during the exchanges
        if (lane == 0)
        {
            tma_store_wait_read0(); // the previous step's bulk store has finished reading sX
            if (wbnext < nWB) issue(wbnext);
        }
        if (GATHER && wbnext < nWB)
        {
            gather(wbnext); // sMap holds the indices of the next batch; sIn is consumed
            __syncwarp();
            if (wbnext + GW < nWB) fetch_map(wbnext + GW);
        }
        __syncwarp();
        // ---- exchange 1: U_M = lamJ A2 + R.  lane (e,r) scatters, lane (e,p') gathers its [q'][r] block
        double acc[NM][NM];
        if (active)
        {
#pragma unroll
            for (int qq = 0; qq < NM; ++qq)
#pragma unroll
                for (int pp = 0; pp < NM; ++pp) sX[e * ES + pp * PS + qq * NM + s1] = fma(lamJ, A2[qq][pp], R[qq][pp]);
        }
        __syncwarp();
        if (active)
        {
            const double *v = sX + e * ES + s1 * PS;
#pragma unroll
            for (int qq = 0; qq < NM; ++qq)
            {
                double col[NM];
#pragma unroll
                for (int r = 0; r < NM; ++r) col[r] = v[qq * NM + r];
#pragma unroll
                for (int rr = 0; rr < NM; ++rr)
                {
                    double sacc = 0.0;
                    bool sset   = false;
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                        if (MNZ(rr, r))
                        {
                            sacc = sset ? fma(KM(rr, r), col[r], sacc) : KM(rr, r) * col[r];
                            sset = true;
                        }
                    acc[rr][qq] = sacc;
                }
            }
        }
        __syncwarp();
        // ---- exchange 2: U_K = jg22 A2
        if (active)
        {
#pragma unroll
            for (int qq = 0; qq < NM; ++qq)
#pragma unroll
                for (int pp = 0; pp < NM; ++pp) sX[e * ES + pp * PS + qq * NM + s1] = jg22 * A2[qq][pp];
        }
        __syncwarp();
        if (active)
        {
            const double *v = sX + e * ES + s1 * PS;
            double col[NM][NM];
#pragma unroll
            for (int qq = 0; qq < NM; ++qq)
#pragma unroll
                for (int r = 0; r < NM; ++r) col[qq][r] = v[qq * NM + r];
#pragma unroll
            for (int qq = 0; qq < NM; ++qq)
#pragma unroll
                for (int rr = 0; rr < NM; ++rr)
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                        if (KNZ(rr, r)) acc[rr][qq] = fma(KK(rr, r), col[qq][r], acc[rr][qq]);
        }
        __syncwarp(); // every lane has read its exchange block: sX becomes the output staging buffer
        if (active)
        {
#pragma unroll
            for (int rr = 0; rr < NM; ++rr)
#pragma unroll
                for (int qq = 0; qq < NM; ++qq) sX[e * NM3 + rr * NM2 + qq * NM + s1] = acc[rr][qq];
        }
        if (GATHER ? out_tma_ok(wb) : tma_in)
        {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0)
            {
                tma_store_1d(args.out + (size_t)wb * EPW * NM3, sX, (uint32_t)(ne * NM3 * 8));
                tma_store_commit();
            }
        }
        else
        {
            __syncwarp();
            double *dst = args.out + (size_t)wb * EPW * NM3;
            for (int i = lane; i < ne * NM3; i += 32) dst[i] = sX[i];
        }
        __syncwarp();
    }
    if (lane == 0) tma_store_wait0();
#undef KM
#undef KK
#undef KNZ
#undef MNZ
#undef KLO
#undef KHI
}

} // namespace nekmf
#include "hex_kron_full.cuh"
#include "hex_kron_rows.cuh"
#include "hex_kron_fullrows.cuh"
#include "hex_kron_lane.cuh"
namespace nekmf
{

// full constant metric of every element for hex_helm_kronfull_kernel
__global__ void kron_prepare_full_kernel(const double *__restrict__ jac, const double *__restrict__ df, int nElmt,
                                         double *__restrict__ geo8)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nElmt) return;
    double f[9];
#pragma unroll
    for (int n = 0; n < 9; ++n) f[n] = df[(size_t)n * nElmt + e];
    const double j = jac[e];
    double *g      = geo8 + (size_t)e * 8;
    g[0] = j;
    g[1] = j * (f[0] * f[0] + f[3] * f[3] + f[6] * f[6]);
    g[2] = j * (f[1] * f[1] + f[4] * f[4] + f[7] * f[7]);
    g[3] = j * (f[2] * f[2] + f[5] * f[5] + f[8] * f[8]);
    g[4] = j * (f[0] * f[1] + f[3] * f[4] + f[6] * f[7]);
    g[5] = j * (f[0] * f[2] + f[3] * f[5] + f[6] * f[8]);
    g[6] = j * (f[1] * f[2] + f[4] * f[5] + f[7] * f[8]);
    g[7] = 0.0;
}

// G off-diagonal == 0 for every element?  (computed exactly as the quadrature-space kernel would)
__global__ void kron_prepare_kernel(const double *__restrict__ jac, const double *__restrict__ df, int nElmt,
                                    double *__restrict__ geo4, int *__restrict__ nondiag)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nElmt) return;
    double f[9];
#pragma unroll
    for (int n = 0; n < 9; ++n) f[n] = df[(size_t)n * nElmt + e];
    const double m00 = f[0] * f[0] + f[3] * f[3] + f[6] * f[6];
    const double m01 = f[0] * f[1] + f[3] * f[4] + f[6] * f[7];
    const double m02 = f[0] * f[2] + f[3] * f[5] + f[6] * f[8];
    const double m11 = f[1] * f[1] + f[4] * f[4] + f[7] * f[7];
    const double m12 = f[1] * f[2] + f[4] * f[5] + f[7] * f[8];
    const double m22 = f[2] * f[2] + f[5] * f[5] + f[8] * f[8];
    if (m01 != 0.0 || m02 != 0.0 || m12 != 0.0) atomicOr(nondiag, 1);
    const double j = jac[e];
    geo4[(size_t)e * 4 + 0] = j;
    geo4[(size_t)e * 4 + 1] = j * m00;
    geo4[(size_t)e * 4 + 2] = j * m11;
    geo4[(size_t)e * 4 + 3] = j * m22;
}

struct KronState
{
    void *tab      = nullptr;
    void *tab_full = nullptr; // KronFullTab<NM>
    bool sparse_k  = false;
    double *d_geo4 = nullptr;
    double *d_geo8 = nullptr; // full constant metric (non-diagonal collections)
    bool use_full  = false;
    bool sparse_full = false; // K, M and S all have the modified-basis sparsity patterns
    bool rows_kind   = false; // nm = 7..10: row-streaming kernels (diagonal metric: hex_kron_rows.cuh, full metric: hex_kron_fullrows.cuh; no fused gather)
    int blocks_per_sm_full = 0;
    int blocks_per_sm = 0, blocks_per_sm_lane = 0;
    int bps_slot[2][2] = {{0, 0}, {0, 0}}; // [gather][8-warp variant]
    // the quadrature-space launcher this operator falls back to for non-diagonal metrics
    int (*fallback)(nekmf_op_s *, const double *const in[3], double *const out[3]) = nullptr;
    void *fallback_state                                                           = nullptr;
    void (*fallback_free)(void *)                                                  = nullptr;
    std::string fallback_name;
    bool use_kron = false;
};

template <int NM, bool GATHER, int WSEL = 0> static int kron_launch_t(nekmf_op_s *op, KronState *st, const double *in, double *out)
{
    using Cfg = KronCfg<NM, GATHER, WSEL>;
    auto kern = st->sparse_full ? hex_helm_kron_kernel<NM, 2, GATHER, WSEL>
                                : (st->sparse_k ? hex_helm_kron_kernel<NM, 1, GATHER, WSEL> : hex_helm_kron_kernel<NM, 0, GATHER, WSEL>);
    int &bps  = st->bps_slot[GATHER ? 1 : 0][WSEL == 8 ? 1 : 0];
    if (bps == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("kron kernel does not fit on an SM"); return NEKMF_ERR_CUDA; }
        bps = nb;
    }
    KronArgs a;
    a.in = in; a.out = out; a.geo4 = st->d_geo4 + (size_t)op->run_e0 * 4; a.nElmt = op->run_ne; a.lambda = op->lambda;
    a.map  = GATHER ? op->gather_map + (size_t)op->run_e0 * Cfg::NM3 : nullptr;
    a.sign = GATHER && op->gather_sign ? op->gather_sign + (size_t)op->run_e0 * Cfg::NM3 : nullptr;
    static const int gather_rows = [] { const char *v = getenv("NEKMF_GATHER_ROWS"); return (v && v[0] == '0') ? 0 : 1; }(); // A/B knob
    a.gather_rows = gather_rows;
    a.skip        = GATHER ? op->gather_skip : nullptr;
    a.io_aligned = GATHER ? ((((uintptr_t)out) & 15) == 0) : ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0;
    const int nBatches = (op->run_ne + Cfg::EPW * Cfg::WARPS - 1) / (Cfg::EPW * Cfg::WARPS);
    int grid           = bps * NUM_SMS;
    if (grid > nBatches) grid = nBatches;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(*static_cast<const KronTab<NM> *>(st->tab), a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <int NM> static int kron_full_launch(nekmf_op_s *op, KronState *st, const double *in, double *out)
{
    using Cfg = KronFullCfg<NM>;
    auto kern = st->sparse_full ? hex_helm_kronfull_kernel<NM, true> : hex_helm_kronfull_kernel<NM, false>;
    if (st->blocks_per_sm_full == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("full-metric kron kernel does not fit on an SM"); return NEKMF_ERR_CUDA; }
        st->blocks_per_sm_full = nb;
    }
    KronFullArgs a;
    a.in = in; a.out = out; a.geo8 = st->d_geo8 + (size_t)op->run_e0 * 8; a.nElmt = op->run_ne; a.lambda = op->lambda;
    a.io_aligned = ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0;
    const int nBatches = (op->run_ne + Cfg::EPW * Cfg::WARPS - 1) / (Cfg::EPW * Cfg::WARPS);
    int grid           = st->blocks_per_sm_full * NUM_SMS;
    if (grid > nBatches) grid = nBatches;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(*static_cast<const KronFullTab<NM> *>(st->tab_full), a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <int NM, bool DIAG> static int kron_fullrows_launch(nekmf_op_s *op, KronState *st, const double *in, double *out)
{
    using Cfg = KronFullRowsCfg<NM>;
    auto kern = hex_helm_kronfullrows_kernel<NM, DIAG>;
    if (st->blocks_per_sm_full == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("full-metric row-streaming kron kernel does not fit on an SM"); return NEKMF_ERR_CUDA; }
        st->blocks_per_sm_full = nb;
    }
    KronFullArgs a;
    a.in = in; a.out = out; a.geo8 = st->d_geo8 + (size_t)op->run_e0 * 8; a.nElmt = op->run_ne; a.lambda = op->lambda;
    a.io_aligned = ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0;
    const int nBatches = (op->run_ne + Cfg::EPW * Cfg::WARPS - 1) / (Cfg::EPW * Cfg::WARPS);
    int grid           = st->blocks_per_sm_full * NUM_SMS;
    if (grid > nBatches) grid = nBatches;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(*static_cast<const KronFullTab<NM> *>(st->tab_full), a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <int NM> static int kron_rows_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    KronState *st = static_cast<KronState *>(op->kstate);
    if (st->use_full && !op->gather_map)
        return st->use_kron ? kron_fullrows_launch<NM, true>(op, st, in[0], out[0]) : kron_fullrows_launch<NM, false>(op, st, in[0], out[0]);
    if (!st->use_kron || op->gather_map)
    {
        if (op->gather_map) { set_error("fused gather requested from a kernel that does not provide it"); return NEKMF_ERR_ARG; }
        void *saved  = op->kstate;
        op->kstate   = st->fallback_state;
        const int rc = st->fallback(op, in, out);
        op->kstate   = saved;
        return rc;
    }
    using Cfg = KronRowsCfg<NM>;
    auto kern = hex_helm_kronrows_kernel<NM>;
    if (st->blocks_per_sm == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("row-streaming kron kernel does not fit on an SM"); return NEKMF_ERR_CUDA; }
        st->blocks_per_sm = nb;
    }
    KronArgs a;
    a.in = in[0]; a.out = out[0]; a.geo4 = st->d_geo4 + (size_t)op->run_e0 * 4; a.nElmt = op->run_ne; a.lambda = op->lambda;
    a.map = nullptr; a.sign = nullptr; a.gather_rows = 0; a.skip = nullptr;
    a.io_aligned = ((((uintptr_t)in[0]) | ((uintptr_t)out[0])) & 15) == 0;
    const int nBatches = (op->run_ne + Cfg::EPW * Cfg::WARPS - 1) / (Cfg::EPW * Cfg::WARPS);
    int grid           = st->blocks_per_sm * NUM_SMS;
    if (grid > nBatches) grid = nBatches;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(*static_cast<const KronTab<NM> *>(st->tab), a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

// nm <= 4: one lane per element (hex_kron_lane.cuh; 1.06 -> 1.18, 0.87 -> 0.99, 0.61 -> 0.79 of the HBM copy peak at
// nm = 2, 3, 4); NEKMF_HEX_KRON_LANE=0 keeps the slab-per-lane kernel
template <int NM> static int kron_lane_launch(nekmf_op_s *op, KronState *st, const double *in, double *out)
{
    using Cfg = KronLaneCfg<NM>;
    auto kern = st->sparse_k ? hex_helm_kronlane_kernel<NM, true> : hex_helm_kronlane_kernel<NM, false>;
    if (st->blocks_per_sm_lane == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("lane-per-element kron kernel does not fit on an SM"); return NEKMF_ERR_CUDA; }
        st->blocks_per_sm_lane = nb;
    }
    KronArgs a;
    a.in = in; a.out = out; a.geo4 = st->d_geo4 + (size_t)op->run_e0 * 4; a.nElmt = op->run_ne; a.lambda = op->lambda;
    a.map = nullptr; a.sign = nullptr; a.gather_rows = 0; a.skip = nullptr;
    a.io_aligned = ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0;
    const int nBatches = (op->run_ne + 32 * Cfg::WARPS - 1) / (32 * Cfg::WARPS);
    int grid           = st->blocks_per_sm_lane * NUM_SMS;
    if (grid > nBatches) grid = nBatches;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(*static_cast<const KronTab<NM> *>(st->tab), a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <int NM> static int kron_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    KronState *st = static_cast<KronState *>(op->kstate);
    if (st->use_full && !op->gather_map) return kron_full_launch<NM>(op, st, in[0], out[0]);
    if constexpr (NM <= 4)
    {
        static const bool lane_on = [] { const char *v = getenv("NEKMF_HEX_KRON_LANE"); return !(v && v[0] == '0'); }();
        if (lane_on && st->use_kron && !op->gather_map) return kron_lane_launch<NM>(op, st, in[0], out[0]);
    }
    if (!st->use_kron)
    {
        if (op->gather_map) { set_error("fused gather requested from a kernel that does not provide it"); return NEKMF_ERR_ARG; }
        void *saved = op->kstate;
        op->kstate  = st->fallback_state;
        const int rc = st->fallback(op, in, out);
        op->kstate  = saved;
        return rc;
    }
    if constexpr (NM == 5)
    {
        const char *v = getenv("NEKMF_KRON_WARPS"); // A/B knob, read per launch
        if (v && atoi(v) == 8)
            return op->gather_map ? kron_launch_t<NM, true, 8>(op, st, in[0], out[0]) : kron_launch_t<NM, false, 8>(op, st, in[0], out[0]);
    }
    return op->gather_map ? kron_launch_t<NM, true>(op, st, in[0], out[0]) : kron_launch_t<NM, false>(op, st, in[0], out[0]);
}

template <int NM> static void kron_wrap(nekmf_op_s *op)
{
    // 1-D mass and stiffness matrices from the operator's own tables
    const int nq = op->nq[0];
    auto *tab    = new KronTab<NM>;
    const double *B = op->b[0].data(), *dB = op->db[0].data(), *w = op->ws[0].data();
    double kmax = 0.0, koff = 0.0;
    for (int a = 0; a < NM; ++a)
        for (int c = a; c < NM; ++c)
        {
            double m = 0.0, k = 0.0;
            for (int i = 0; i < nq; ++i)
            {
                m += B[a * nq + i] * w[i] * B[c * nq + i];
                k += dB[a * nq + i] * w[i] * dB[c * nq + i];
            }
            tab->Ms[tri(a, c, NM)] = m;
            tab->Ks[tri(a, c, NM)] = k;
            const bool pattern = a == c || (a < 2 && c < 2);
            if (pattern) kmax = std::fmax(kmax, std::fabs(k));
            else koff = std::fmax(koff, std::fabs(k));
        }
    auto *tabf = new KronFullTab<NM>;
    memcpy(tabf->Ms, tab->Ms, sizeof(tab->Ms));
    memcpy(tabf->Ks, tab->Ks, sizeof(tab->Ks));
    // sparsity patterns the full-metric kernel's sparse variant relies on (hex_kron_full.cuh): entries outside them
    // must be at round-off level
    double mmax = 0.0, moff = 0.0, smax = 0.0, soff = 0.0;
    for (int a = 0; a < NM; ++a)
        for (int c = 0; c < NM; ++c)
        {
            double sv = 0.0;
            for (int i = 0; i < nq; ++i) sv += dB[a * nq + i] * w[i] * B[c * nq + i];
            tabf->S[a * NM + c] = sv;
            const int lo = a < c ? a : c, hi = a < c ? c : a;
            const bool spat = hi < 2 || (lo < 2 && hi == 2) || (lo >= 2 && hi - lo == 1);
            const bool mpat = hi < 2 || (lo < 2 && hi <= 3) || (lo >= 2 && (hi - lo) % 2 == 0 && hi - lo <= 2);
            const double mv = std::fabs(tab->Ms[tri(a, c, NM)]);
            if (spat) smax = std::fmax(smax, std::fabs(sv)); else soff = std::fmax(soff, std::fabs(sv));
            if (mpat) mmax = std::fmax(mmax, mv); else moff = std::fmax(moff, mv);
        }
    const bool sparse_ms = moff <= 1e-14 * mmax && soff <= 1e-14 * smax;
    // entries outside the (vertex block + diagonal) pattern are quadrature round-off for the modified
    // basis; drop them only when they are at round-off level relative to the matrix
    const bool sparse_k = koff <= 1e-14 * kmax;
    KronState *st      = new KronState;
    st->tab            = tab;
    st->tab_full       = tabf;
    st->sparse_k       = sparse_k;
    st->sparse_full    = sparse_k && sparse_ms;
    st->fallback       = op->launch;
    st->fallback_state = op->kstate;
    st->fallback_free  = op->kstate_free;
    st->fallback_name  = op->kname;
    op->kstate         = st;
    op->kstate_free    = [](void *p) {
        KronState *s = static_cast<KronState *>(p);
        if (s->fallback_state && s->fallback_free) s->fallback_free(s->fallback_state);
        delete static_cast<KronTab<NM> *>(s->tab);
        delete static_cast<KronFullTab<NM> *>(s->tab_full);
        cudaFree(s->d_geo4);
        cudaFree(s->d_geo8);
        delete s;
    };
    op->launch = kron_launch<NM>;
}

// nm = 7..10: hex_helm_kronrows_kernel.  The kernel has the modified-basis sparsity of M and K compiled in, so the
// wrap is skipped (quadrature-space kernel stays) when the tables do not show it.
template <int NM> static void kron_rows_wrap(nekmf_op_s *op)
{
    const int nq = op->nq[0];
    auto *tab    = new KronTab<NM>;
    const double *B = op->b[0].data(), *dB = op->db[0].data(), *w = op->ws[0].data();
    double kmax = 0.0, koff = 0.0, mmax = 0.0, moff = 0.0;
    for (int a = 0; a < NM; ++a)
        for (int c = a; c < NM; ++c)
        {
            double m = 0.0, k = 0.0;
            for (int i = 0; i < nq; ++i)
            {
                m += B[a * nq + i] * w[i] * B[c * nq + i];
                k += dB[a * nq + i] * w[i] * dB[c * nq + i];
            }
            tab->Ms[tri(a, c, NM)] = m;
            tab->Ks[tri(a, c, NM)] = k;
            if (rows_knz(a, c)) kmax = std::fmax(kmax, std::fabs(k)); else koff = std::fmax(koff, std::fabs(k));
            if (rows_mnz(a, c)) mmax = std::fmax(mmax, std::fabs(m)); else moff = std::fmax(moff, std::fabs(m));
        }
    if (!(koff <= 1e-14 * kmax && moff <= 1e-14 * mmax))
    {
        delete tab;
        return;
    }
    // full-metric kernel (hex_kron_fullrows.cuh): the mixed matrix S and its sparsity pattern
    auto *tabf = new KronFullTab<NM>;
    memcpy(tabf->Ms, tab->Ms, sizeof(tab->Ms));
    memcpy(tabf->Ks, tab->Ks, sizeof(tab->Ks));
    double smax = 0.0, soff = 0.0;
    for (int a = 0; a < NM; ++a)
        for (int c = 0; c < NM; ++c)
        {
            double sv = 0.0;
            for (int i = 0; i < nq; ++i) sv += dB[a * nq + i] * w[i] * B[c * nq + i];
            tabf->S[a * NM + c] = sv;
            if (fr_snz(a, c)) smax = std::fmax(smax, std::fabs(sv)); else soff = std::fmax(soff, std::fabs(sv));
        }
    const char *vfr = getenv("NEKMF_HEX_KRON_FULLROWS"); // =0: sheared elements stay on the quadrature-space kernel (A/B)
    if (!(soff <= 1e-14 * smax) || (vfr && vfr[0] == '0'))
    {
        delete tabf;
        tabf = nullptr;
    }
    KronState *st      = new KronState;
    st->tab            = tab;
    st->tab_full       = tabf;
    st->sparse_full    = tabf != nullptr;
    st->sparse_k       = true;
    st->rows_kind      = true;
    st->fallback       = op->launch;
    st->fallback_state = op->kstate;
    st->fallback_free  = op->kstate_free;
    st->fallback_name  = op->kname;
    op->kstate         = st;
    op->kstate_free    = [](void *p) {
        KronState *s = static_cast<KronState *>(p);
        if (s->fallback_state && s->fallback_free) s->fallback_free(s->fallback_state);
        delete static_cast<KronTab<NM> *>(s->tab);
        delete static_cast<KronFullTab<NM> *>(s->tab_full);
        cudaFree(s->d_geo4);
        cudaFree(s->d_geo8);
        delete s;
    };
    op->launch = kron_rows_launch<NM>;
    op->kron   = 1;
}

// called from select_hex_fast after the quadrature-space launcher is installed
void kron_maybe_wrap(nekmf_op_s *op)
{
    if (op->shape != NEKMF_HEX || op->optype != NEKMF_HELMHOLTZ || op->deformed) return;
    const char *v = getenv("NEKMF_HEX_KRON"); // NEKMF_HEX_KRON=0: quadrature-space kernel only (cross-kernel tests)
    if (v && v[0] == '0') return;
    switch (op->nm[0])
    {
        case 2: kron_wrap<2>(op); break;
        case 3: kron_wrap<3>(op); break;
        case 4: kron_wrap<4>(op); break;
        case 5: kron_wrap<5>(op); break;
        case 6: kron_wrap<6>(op); break;
        case 7: kron_rows_wrap<7>(op); return;
        case 8: kron_rows_wrap<8>(op); return;
        case 9: kron_rows_wrap<9>(op); return;
        case 10: kron_rows_wrap<10>(op); return;
        case 11: kron_rows_wrap<11>(op); return; // axis-aligned elements only, through hex_kron_fullrows.cuh<DIAG> (kron_geom_changed)
        default: return;
    }
    op->kron = 1;
}

// called after set_geom: decide between the coefficient-space and the quadrature-space kernel
int kron_geom_changed(nekmf_op_s *op)
{
    if (op->kron != 1) return NEKMF_OK;
    KronState *st = static_cast<KronState *>(op->kstate);
    st->use_kron  = false;
    op->gather_ok = false;
    op->kname     = st->fallback_name;
    if (!op->has_jac || !op->has_df || op->nElmt == 0) return NEKMF_OK;
    if (!st->d_geo4) NEKMF_CUDA(cudaMalloc(&st->d_geo4, (size_t)op->nElmt * 4 * 8));
    int *d_flag = nullptr;
    NEKMF_CUDA(cudaMalloc(&d_flag, 4));
    NEKMF_CUDA(cudaMemset(d_flag, 0, 4));
    kron_prepare_kernel<<<(op->nElmt + 255) / 256, 256>>>(op->d_jac, op->d_df, op->nElmt, st->d_geo4, d_flag);
    ++g_launches;
    int flag = 1;
    NEKMF_CUDA(cudaMemcpy(&flag, d_flag, 4, cudaMemcpyDeviceToHost));
    cudaFree(d_flag);
    st->use_full = false;
    st->blocks_per_sm_full = 0; // the full-metric kernel variant may change with the geometry
    if (st->rows_kind)
    {
        char name[96];
        if (flag == 0)
        {
            st->use_kron = true;
            snprintf(name, sizeof(name), "hex_helm_kronrows_kernel<nm=%d>(regular,diagonal metric)", op->nm[0]);
            op->kname = name;
        }
        // the one-intermediate-at-a-time kernel (hex_kron_fullrows.cuh): sheared elements, and -- with the S terms
        // compiled out -- axis-aligned ones where it measured faster than hex_kron_rows.cuh
        const char *vd     = getenv("NEKMF_HEX_KRON_DIAGROWS"); // =1 / =0: force / forbid it for diagonal metrics (A/B)
        // diagonal metric, measured against hex_kron_rows.cuh (profiles/r02_sweep_hex_diagrows_*.jsonl, fraction of the
        // HBM peak): nm = 7: 0.72 / 0.46, 8: 0.30 / 0.30, 9: 0.25 / 0.19, 10: 0.41 / 0.29
        // nm = 11: only the DIAG variant of hex_kron_fullrows.cuh is used (hex_kron_rows.cuh fits two warps of 22 lanes per
        // SM there and loses to the quadrature-space kernel)
        const bool diag_fr = vd ? vd[0] == '1' : op->nm[0] != 8;
        if (op->nm[0] == 11 && !(st->tab_full && diag_fr))
        {
            st->use_kron = false;
            op->kname    = st->fallback_name;
        }
        // full metric, measured against the quadrature-space kernel (profiles/r02_sweep_hex_fullrows_*.jsonl):
        // 3.1x / 1.3x / 1.2x faster at nm = 7 / 8 / 9, 0.87x at nm = 10 (kept on the quadrature-space kernel)
        const char *vf     = getenv("NEKMF_HEX_KRON_FULLROWS"); // =all: at every instantiated order (A/B)
        const bool full_fr = flag != 0 && (op->nm[0] <= 9 || (vf && vf[0] == 'a'));
        if (st->tab_full && (full_fr || (flag == 0 && diag_fr)))
        {
            if (!st->d_geo8) NEKMF_CUDA(cudaMalloc(&st->d_geo8, (size_t)op->nElmt * 8 * 8));
            kron_prepare_full_kernel<<<(op->nElmt + 255) / 256, 256>>>(op->d_jac, op->d_df, op->nElmt, st->d_geo8);
            ++g_launches;
            NEKMF_CUDA(cudaGetLastError());
            st->use_full = true;
            snprintf(name, sizeof(name), "hex_helm_kronfullrows_kernel<nm=%d,%s>(regular,%s metric)", op->nm[0],
                     flag == 0 ? "diag" : "full", flag == 0 ? "diagonal" : "full");
            op->kname = name;
        }
        return NEKMF_OK;
    }
    if (flag != 0)
    {
        // constant but non-diagonal metric (sheared / rotated affine elements): full-metric coefficient-space kernel
        if (!st->d_geo8) NEKMF_CUDA(cudaMalloc(&st->d_geo8, (size_t)op->nElmt * 8 * 8));
        kron_prepare_full_kernel<<<(op->nElmt + 255) / 256, 256>>>(op->d_jac, op->d_df, op->nElmt, st->d_geo8);
        ++g_launches;
        NEKMF_CUDA(cudaGetLastError());
        st->use_full = true;
        char name[96];
        snprintf(name, sizeof(name), "hex_helm_kronfull_kernel<nm=%d,%s>(regular,full metric)", op->nm[0],
                 st->sparse_full ? "sparse" : "dense");
        op->kname = name;
    }
    if (flag == 0)
    {
        st->use_kron  = true;
        op->gather_ok = true;
        char name[96];
        snprintf(name, sizeof(name), "hex_helm_kron_kernel<nm=%d,%s>(regular,diagonal metric)", op->nm[0],
                 st->sparse_full ? "sparseKM" : (st->sparse_k ? "sparseK" : "denseK"));
        op->kname = name;
    }
    return NEKMF_OK;
}

} // namespace nekmf
