// prism_helm_dmma.cu -- Helmholtz on REGULAR prisms of any orientation (general affine: non-extruded, sheared) at
// nm = 5..7, the whole reference chain fused in quadrature space on FP64 tensor-core tiles, one warp per element.
//
// Reference semantics: MatrixFreeOps/Helmholtz.h:1291-1458 (HelmholtzPrismImpl, DEFORMED = false):
//   bwd = BwdTrans(in);  out = lambda IProduct(bwd);  (d0,d1,d2) = PhysDerivTensor(bwd);  g = G(xi) d  (Laplacian metric
//   with the collapsed-coordinate factors);  out += IProduct_{dB0}(g0) + IProduct_{B,dB1}(g1) + IProduct_{dB2}(g2).
// Because dbdata = D bdata, the four IProducts are ONE: out = B^T [ lambda wJ bwd + D0^T(wJ g0) + D1^T(wJ g1) + D2^T(wJ g2) ].
//
// Layout: after the chained BwdTrans tiles of prism_dmma.cu lane (g, t) holds u[k][j = g][i = 2t, 2t+1] for all planes k
// (the m8n8k4 C fragment).  From there
//   xi_0 derivative and its transpose  the lane's two values are the A operand of two k-steps whose contracted index is
//                                      ordered i = (2t | 2t+1); B operand = rows / columns of D (fixed registers)
//   xi_1 derivative and its transpose  A = D^T / D (fixed), B = the plane in the B-fragment layout: one 16-byte store
//                                      and two 8-byte loads through a per-warp 8 x 8 scratch
//   xi_2 derivative and its transpose  the k-line of the lane's two points is in its registers: plain DFMA
//   metric, weights, mass term         pointwise in registers (constant factors; per element only h0(i) terms vary)
//   IProduct                           the accumulated plane G[k] is the B operand of the first tile AS IT IS (A = the
//                                      matching columns of the basis), its C fragment the B operand of the second; the
//                                      collapsed xi_2 contraction on the fly in the lane that owns (p, q)
// 112 DMMA per element instead of the 224 of the eight-term coefficient-space kernel (dense_helm.cu) and about a third
// of the shared-memory traffic of the pencil kernel (shape_kernels.cuh).
#include "op_internal.h"
#include <stdlib.h>
#include <string.h>

namespace nekmf
{

template <int NM> struct PrismHelmTab
{
    static constexpr int NQ0 = NM + 1, NQ2 = NM, NPAIR = NM * (NM + 1) / 2;
    double b0[NM * NQ0];    // bdata of direction 0 (and 1), [p][i]
    double b2[NPAIR * NQ2]; // bdata of direction 2, rows (p, r)
    double D0[NQ0 * NQ0];   // D[a*nq+b] = dh_a/dz(z_b), directions 0 and 1
    double D2[NQ2 * NQ2];
    double w0[NQ0], w2[NQ2]; // weights (collapsed-coordinate factor folded into w2)
    double h0[NQ0], h1[NQ2]; // (1 + z0_i) / 2,  2 / (1 - z2_k)
};

struct PrismHelmArgs
{
    const double *in;
    double *out;
    const double *jac, *df; // regular geometry: jac[nElmt], df[9][dfStride]
    size_t dfStride;
    int nElmt;
    int in_aligned, out_aligned;
    double lambda;
};

__device__ __forceinline__ void ph_mma(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NM, int W> struct PrismHelmCfg
{
    static constexpr int NQ0 = NM + 1, NQ1 = NM + 1, NQ2 = NM, NPAIR = NM * (NM + 1) / 2;
    static constexpr int NMT = NM * NPAIR, NQT = NQ0 * NQ1 * NQ2;
    static constexpr int EPB = (NMT % 2) ? 2 : 1;
    static constexpr int BUF = EPB * NMT;
    static constexpr int SU  = NQ2 * 64;            // u planes, 8 x 8 each
    static constexpr int SX  = NQ2 * 64;            // the planes of the xi_1 flux (one scratch plane each: no barrier per plane)
    static constexpr int STG = (NMT + 1) & ~1;      // output staging
    static constexpr int PER_WARP = 2 * BUF + SU + SX + STG + 2;
    static constexpr int B2S = ((NPAIR + NM) * NQ2 + 1) & ~1; // + nm zero rows: lines of absent modes read zeros, no index clamps
    static constexpr int WARPS = W, T = WARPS * 32; // 8: 254 registers, 12: 168, 16: 128 (about 100 bytes of spills)
    static constexpr size_t SMEM = (size_t)(B2S + WARPS * PER_WARP) * 8 + 16;
};

template <int NM, int W>
__global__ void __launch_bounds__(PrismHelmCfg<NM, W>::T, 1)
    prism_helm_dmma_kernel(const __grid_constant__ PrismHelmTab<NM> tab, const __grid_constant__ PrismHelmArgs args)
{
    using Cfg = PrismHelmCfg<NM, W>;
    constexpr int NQ0 = Cfg::NQ0, NQ1 = Cfg::NQ1, NQ2 = Cfg::NQ2, NPAIR = Cfg::NPAIR, NMT = Cfg::NMT;
    constexpr int BUF = Cfg::BUF, EPB = Cfg::EPB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = lane & 3, g = lane >> 2;
    double *sB2   = reinterpret_cast<double *>(smem_raw);
    double *wbase = sB2 + Cfg::B2S + (size_t)warp * Cfg::PER_WARP;
    double *sU    = wbase + 2 * BUF;
    double *sX    = sU + Cfg::SU;
    double *sStg  = sX + Cfg::SX;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sStg + Cfg::STG);

    for (int i = threadIdx.x; i < Cfg::B2S; i += Cfg::T) sB2[i] = i < NPAIR * NQ2 ? tab.b2[i] : 0.0;
    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int nBlk = (args.nElmt + EPB - 1) / EPB;
    const int GW = gridDim.x * Cfg::WARPS, gw = blockIdx.x * Cfg::WARPS + warp;
    auto blk_ne = [&](int b) { return args.nElmt - EPB * b >= EPB ? EPB : 1; };
    auto tma_ok = [&](int b) { return args.in_aligned && blk_ne(b) == EPB; };
    auto issue  = [&](int b, int slot) { // lane 0
        if (!tma_ok(b)) return;
        mbar_expect_tx(bar + slot, (uint32_t)(BUF * 8));
        tma_load_1d(wbase + slot * BUF, args.in + (size_t)b * BUF, (uint32_t)(BUF * 8), bar + slot);
    };
    auto mpr = [](int p) { return p * NM - p * (p - 1) / 2; };

    // ---- fixed fragments (lane (g, t))
    const int i0 = 2 * t, i1 = 2 * t + 1;
    const bool vi0 = i0 < NQ0, vi1 = i1 < NQ0, vg0 = g < NQ0, vgm = g < NM;
    double Abw[2], Ad1[2], At1[2], Bd0[2], Bt0[2], Aip[2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
    {
        const int c4 = 4 * s + t;  // contracted index of a plain k-step
        const int c2 = 2 * t + s;  // contracted index of a k-step fed from a C fragment
        Abw[s] = (vg0 && c4 < NM) ? tab.b0[c4 * NQ0 + g] : 0.0;            // BwdTrans: A[i | j = g][p | q = c4]
        Ad1[s] = (vg0 && c4 < NQ0) ? tab.D0[c4 * NQ0 + g] : 0.0;           // d/dxi_1:  A[j' = g][j = c4] = D[j][j']
        At1[s] = (vg0 && c4 < NQ0) ? tab.D0[g * NQ0 + c4] : 0.0;           // its transpose: A[a = g][j = c4] = D[a][j]
        Bd0[s] = (vg0 && c2 < NQ0) ? tab.D0[c2 * NQ0 + g] : 0.0;           // d/dxi_0:  B[i = c2][i' = g] = D[i][i']
        Bt0[s] = (vg0 && c2 < NQ0) ? tab.D0[g * NQ0 + c2] : 0.0;           // its transpose: B[i = c2][a = g] = D[a][i]
        Aip[s] = (vgm && c2 < NQ0) ? tab.b0[g * NQ0 + c2] : 0.0;           // IProduct: A[p | q = g][i | j = c2]
    }
    const int col0 = (g & 1) ? 4 + (g >> 1) : (g >> 1); // BwdTrans pass 1: tile column g <-> q
    const double h00 = vi0 ? tab.h0[i0] : 0.0, h01 = vi1 ? tab.h0[vi1 ? i1 : 0] : 0.0;
    const double wi0 = vi0 ? tab.w0[i0] : 0.0, wi1 = vi1 ? tab.w0[vi1 ? i1 : 0] : 0.0;
    const double wj  = g < NQ1 ? tab.w0[g] : 0.0;

    uint32_t phase[2] = {0, 0};
    int slot          = 0;
    if (lane == 0 && gw < nBlk) issue(gw, 0);

    for (int b = gw; b < nBlk; b += GW, slot ^= 1)
    {
        const int ne = blk_ne(b);
        double *sIn  = wbase + slot * BUF;
        if (lane == 0 && b + GW < nBlk) issue(b + GW, slot ^ 1);
        if (tma_ok(b))
        {
            mbar_wait(bar + slot, phase[slot]);
            phase[slot] ^= 1;
        }
        else
        {
            const double *src = args.in + (size_t)b * BUF;
            for (int i = lane; i < ne * NMT; i += 32) sIn[i] = __ldg(src + i);
        }
        __syncwarp();

#pragma unroll 1
        for (int e = 0; e < ne; ++e)
        {
            const size_t el = (size_t)b * EPB + e;
            const double *U = sIn + e * NMT;
            // ---- constant geometry of the element (Helmholtz.h:1382-1430 with constant factors)
            double f[9];
#pragma unroll
            for (int n = 0; n < 9; ++n) f[n] = __ldg(args.df + (size_t)n * args.dfStride + el);
            const double J   = __ldg(args.jac + el);
            const double m11 = fma(f[7], f[7], fma(f[4], f[4], f[1] * f[1]));
            const double m22 = fma(f[8], f[8], fma(f[5], f[5], f[2] * f[2]));
            const double m12 = fma(f[7], f[8], fma(f[4], f[5], f[1] * f[2]));
            double A00[2], A01[2], A02[2]; // m00 = h1^2 A00, m01 = h1 A01, m02 = h1 A02 at the lane's two i
#pragma unroll
            for (int s = 0; s < 2; ++s)
            {
                const double h0 = s ? h01 : h00;
                const double a1 = fma(h0, f[2], f[0]), a2 = fma(h0, f[5], f[3]), a3 = fma(h0, f[8], f[6]);
                A00[s] = fma(a3, a3, fma(a2, a2, a1 * a1));
                A01[s] = fma(f[7], a3, fma(f[4], a2, f[1] * a1));
                A02[s] = fma(f[8], a3, fma(f[5], a2, f[2] * a1));
            }
            const double wjJ = wj * J;

            // ---- phase 1: BwdTrans (prism_dmma.cu), u[k][j = g][i = 2t, 2t+1]
            double2 u[NQ2];
            {
                const int q  = col0;
                const int p0 = t, p1 = 4 + t;
                const bool v0 = q < NM && p0 < NM, v1 = q < NM && p1 < NM;
                double c0[NM], c1[NM > 4 ? NM - 4 : 1];
                const int base0 = v0 ? NM * mpr(p0) + q * (NM - p0) : 0;
                const int base1 = v1 ? NM * mpr(p1) + q * (NM - p1) : 0;
#pragma unroll
                for (int r = 0; r < NM; ++r)
                {
                    const bool ok  = v0 && r < NM - p0;
                    const double x = U[ok ? base0 + r : 0];
                    c0[r]          = ok ? x : 0.0;
                }
#pragma unroll
                for (int r = 0; r < NM - 4; ++r)
                {
                    const bool ok  = v1 && r < NM - p1;
                    const double x = U[ok ? base1 + r : 0];
                    c1[r]          = ok ? x : 0.0;
                }
                const bool cor  = p0 == 1 && v0; // singular edge: mode (0, q, 1) also feeds f[1][q] through row (0, 1)
                const double xc = U[cor ? q * NM + 1 : 0];
                const double cc = cor ? xc : 0.0;
                const int row0 = v0 ? mpr(p0) : 0, row1 = v1 ? mpr(p1) : 0;
#pragma unroll
                for (int k = 0; k < NQ2; ++k)
                {
                    double f0 = cc * sB2[NQ2 + k], f1 = 0.0;
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                    {
                        f0 = fma(c0[r], sB2[(row0 + r) * NQ2 + k], f0);
                    }
#pragma unroll
                    for (int r = 0; r < NM - 4; ++r)
                    {
                        f1 = fma(c1[r], sB2[(row1 + r) * NQ2 + k], f1);
                    }
                    double x0 = 0.0, x1 = 0.0, d0 = 0.0, d1 = 0.0;
                    ph_mma(x0, x1, Abw[0], f0);
                    ph_mma(x0, x1, Abw[1], f1);
                    ph_mma(d0, d1, Abw[0], x0);
                    ph_mma(d0, d1, Abw[1], x1);
                    u[k] = make_double2(d0, d1);
                    *reinterpret_cast<double2 *>(sU + k * 64 + g * 8 + 2 * t) = u[k];
                }
            }
            __syncwarp();

            // ---- phase 2: derivatives, metric, weights, transposed derivatives; G[k] accumulates the quadrature-space sum
            double2 G[NQ2];
#pragma unroll
            for (int k = 0; k < NQ2; ++k) G[k] = make_double2(0.0, 0.0);
#pragma unroll
            for (int kk = 0; kk < NQ2; ++kk)
            {
                double d0x = 0.0, d0y = 0.0, d1x = 0.0, d1y = 0.0;
                ph_mma(d0x, d0y, u[kk].x, Bd0[0]);
                ph_mma(d0x, d0y, u[kk].y, Bd0[1]);
                const double q0 = sU[kk * 64 + t * 8 + g], q1 = sU[kk * 64 + (4 + t) * 8 + g];
                ph_mma(d1x, d1y, Ad1[0], q0);
                ph_mma(d1x, d1y, Ad1[1], q1);
                double d2x = tab.D2[kk] * u[0].x, d2y = tab.D2[kk] * u[0].y;
#pragma unroll
                for (int k = 1; k < NQ2; ++k)
                {
                    d2x = fma(tab.D2[k * NQ2 + kk], u[k].x, d2x);
                    d2y = fma(tab.D2[k * NQ2 + kk], u[k].y, d2y);
                }
                const double h1k = tab.h1[kk], h1s = h1k * h1k;
                const double m00x = h1s * A00[0], m01x = h1k * A01[0], m02x = h1k * A02[0];
                const double m00y = h1s * A00[1], m01y = h1k * A01[1], m02y = h1k * A02[1];
                const double wx = wi0 * (wjJ * tab.w2[kk]), wy = wi1 * (wjJ * tab.w2[kk]);
                const double g0x = wx * fma(m02x, d2x, fma(m01x, d1x, m00x * d0x)), g0y = wy * fma(m02y, d2y, fma(m01y, d1y, m00y * d0y));
                const double g1x = wx * fma(m12, d2x, fma(m11, d1x, m01x * d0x)), g1y = wy * fma(m12, d2y, fma(m11, d1y, m01y * d0y));
                const double g2x = wx * fma(m22, d2x, fma(m12, d1x, m02x * d0x)), g2y = wy * fma(m22, d2y, fma(m12, d1y, m02y * d0y));
                // mass term, D0^T (A = the lane's own flux values)
                G[kk].x = fma(args.lambda * wx, u[kk].x, G[kk].x);
                G[kk].y = fma(args.lambda * wy, u[kk].y, G[kk].y);
                ph_mma(G[kk].x, G[kk].y, g0x, Bt0[0]);
                ph_mma(G[kk].x, G[kk].y, g0y, Bt0[1]);
                // D1^T goes through the scratch plane of kk (read back after the plane loop)
                *reinterpret_cast<double2 *>(sX + kk * 64 + g * 8 + 2 * t) = make_double2(g1x, g1y);
                // D2^T in the owning lane
#pragma unroll
                for (int a = 0; a < NQ2; ++a)
                {
                    G[a].x = fma(tab.D2[a * NQ2 + kk], g2x, G[a].x);
                    G[a].y = fma(tab.D2[a * NQ2 + kk], g2y, G[a].y);
                }
            }

            __syncwarp();
#pragma unroll
            for (int kk = 0; kk < NQ2; ++kk)
            {
                const double r0 = sX[kk * 64 + t * 8 + g], r1 = sX[kk * 64 + (4 + t) * 8 + g];
                ph_mma(G[kk].x, G[kk].y, At1[0], r0);
                ph_mma(G[kk].x, G[kk].y, At1[1], r1);
            }

            // ---- phase 3: IProduct without weights; the lane owns (p, q) = (2t, g), (2t + 1, g)
            {
                const int p0 = 2 * t, p1 = 2 * t + 1;
                const int row0 = p0 < NM ? mpr(p0) : 0, row1 = p1 < NM ? mpr(p1) : 0;
                double acc0[NM], acc1[NM > 1 ? NM - 1 : 1];
#pragma unroll
                for (int r = 0; r < NM; ++r) acc0[r] = 0.0;
#pragma unroll
                for (int r = 0; r < NM - 1; ++r) acc1[r] = 0.0;
#pragma unroll
                for (int k = 0; k < NQ2; ++k)
                {
                    double x0 = 0.0, x1 = 0.0, d0 = 0.0, d1 = 0.0;
                    ph_mma(x0, x1, Aip[0], G[k].x); // C1[p = g][j = 2t, 2t + 1]
                    ph_mma(x0, x1, Aip[1], G[k].y);
                    ph_mma(d0, d1, Aip[0], x0);     // C2[q = g][p = 2t, 2t + 1]
                    ph_mma(d0, d1, Aip[1], x1);
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                    {
                        acc0[r] = fma(sB2[(row0 + r) * NQ2 + k], d0, acc0[r]);
                    }
#pragma unroll
                    for (int r = 0; r < NM - 1; ++r)
                    {
                        acc1[r] = fma(sB2[(row1 + r) * NQ2 + k], d1, acc1[r]);
                    }
                    if (NM > 1) acc0[1] = fma(t == 0 ? sB2[NQ2 + k] : 0.0, d1, acc0[1]); // singular edge
                }
                if (g < NM)
                {
                    if (p0 < NM)
                    {
                        double *o = sStg + NM * mpr(p0) + g * (NM - p0);
#pragma unroll
                        for (int r = 0; r < NM; ++r)
                            if (r < NM - p0) o[r] = acc0[r];
                    }
                    if (p1 < NM)
                    {
                        double *o = sStg + NM * mpr(p1) + g * (NM - p1);
#pragma unroll
                        for (int r = 0; r < NM - 1; ++r)
                            if (r < NM - p1) o[r] = acc1[r];
                    }
                }
                __syncwarp();
                double *o = args.out + el * NMT;
                if ((NMT % 2 == 0) && args.out_aligned)
                    for (int i2 = lane; i2 < NMT / 2; i2 += 32)
                        *reinterpret_cast<double2 *>(o + 2 * i2) = *reinterpret_cast<const double2 *>(sStg + 2 * i2);
                else
                    for (int i = lane; i < NMT; i += 32) o[i] = sStg[i];
                __syncwarp(); // staging and the u planes free for the next element
            }
        }
        __syncwarp(); // every lane is done with this buffer before lane 0 refills it
    }
}

template <int NM> struct PrismHelmState
{
    PrismHelmTab<NM> tab;
    int bps = 0, warps = 12;
};

template <int NM, int W> static int prism_helm_fused_launch_w(PrismHelmState<NM> *st, nekmf_op_s *op, const double *in, double *out)
{
    using Cfg = PrismHelmCfg<NM, W>;
    auto kern = prism_helm_dmma_kernel<NM, W>;
    if (st->bps == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::T, Cfg::SMEM));
        if (nb < 1) { set_error("fused prism Helmholtz kernel <%d> does not fit on an SM", NM); return NEKMF_ERR_CUDA; }
        st->bps = nb;
    }
    PrismHelmArgs a;
    a.in = in; a.out = out;
    a.jac = op->d_jac + op->run_e0; a.df = op->d_df + op->run_e0; a.dfStride = (size_t)op->nElmt;
    a.nElmt = op->run_ne; a.lambda = op->lambda;
    a.in_aligned  = (((uintptr_t)in) & 15) == 0;
    a.out_aligned = (((uintptr_t)out) & 15) == 0;
    const int nBlk = (op->run_ne + Cfg::EPB - 1) / Cfg::EPB;
    int grid       = st->bps * NUM_SMS;
    const int need = (nBlk + Cfg::WARPS - 1) / Cfg::WARPS;
    if (grid > need) grid = need;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Cfg::T, Cfg::SMEM, op->run_stream>>>(st->tab, a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}
template <int NM> static int prism_helm_fused_launch_nm(void *state, nekmf_op_s *op, const double *in, double *out)
{
    auto *st = static_cast<PrismHelmState<NM> *>(state);
    if (st->warps == 8) return prism_helm_fused_launch_w<NM, 8>(st, op, in, out);
    if (st->warps == 16) return prism_helm_fused_launch_w<NM, 16>(st, op, in, out);
    return prism_helm_fused_launch_w<NM, 12>(st, op, in, out);
}

template <int NM> static void *prism_helm_fused_create_nm(nekmf_op_s *op)
{
    using Tab = PrismHelmTab<NM>;
    constexpr int NQ0 = NM + 1, NQ2 = NM;
    if ((int)op->b[0].size() != NM * NQ0 || op->b[2].size() != sizeof(Tab::b2) / 8 || (int)op->D[0].size() != NQ0 * NQ0 ||
        (int)op->D[2].size() != NQ2 * NQ2)
        return nullptr;
    if (op->b[1] != op->b[0] || op->ws[1] != op->ws[0] || op->D[1] != op->D[0]) return nullptr; // directions 0 and 1 share tables
    auto *st = new PrismHelmState<NM>;
    memcpy(st->tab.b0, op->b[0].data(), sizeof(st->tab.b0));
    memcpy(st->tab.b2, op->b[2].data(), sizeof(st->tab.b2));
    memcpy(st->tab.D0, op->D[0].data(), sizeof(st->tab.D0));
    memcpy(st->tab.D2, op->D[2].data(), sizeof(st->tab.D2));
    memcpy(st->tab.w0, op->ws[0].data(), sizeof(st->tab.w0));
    memcpy(st->tab.w2, op->ws[2].data(), sizeof(st->tab.w2));
    for (int i = 0; i < NQ0; ++i) st->tab.h0[i] = 0.5 * (1.0 + op->Z[0][i]);
    for (int k = 0; k < NQ2; ++k) st->tab.h1[k] = 2.0 / (1.0 - op->Z[2][k]);
    st->warps = NM == 6 ? 16 : 12; // measured (profiles/r02_sweep_prism_fused_w*.jsonl): nm = 7: 0.98 / 0.85 / 0.90 ms at 8 / 12 / 16 warps, nm = 6: 1.17 / 1.05 / 0.97
    if (const char *vw = getenv("NEKMF_PRISM_FUSED_WARPS")) // A/B knob: warps per SM (register budget)
    {
        const int w = atoi(vw);
        if (w == 8 || w == 12 || w == 16) st->warps = w;
    }
    return st;
}

// fused quadrature-space Helmholtz for general regular prisms (called from dense_helm.cu: prism_geom_changed / prism_launch)
void *prism_helm_fused_create(nekmf_op_s *op)
{
    if (op->shape != NEKMF_PRISM || op->optype != NEKMF_HELMHOLTZ || op->deformed) return nullptr;
    const int nm = op->nm[0];
    if (op->nm[1] != nm || op->nm[2] != nm || op->nq[0] != nm + 1 || op->nq[1] != nm + 1 || op->nq[2] != nm) return nullptr;
    switch (nm)
    {
        case 5: return prism_helm_fused_create_nm<5>(op);
        case 6: return prism_helm_fused_create_nm<6>(op);
        case 7: return prism_helm_fused_create_nm<7>(op);
        default: return nullptr;
    }
}
int prism_helm_fused_launch(void *state, nekmf_op_s *op, const double *in, double *out)
{
    switch (op->nm[0])
    {
        case 5: return prism_helm_fused_launch_nm<5>(state, op, in, out);
        case 6: return prism_helm_fused_launch_nm<6>(state, op, in, out);
        case 7: return prism_helm_fused_launch_nm<7>(state, op, in, out);
        default: set_error("fused prism Helmholtz: no instantiation for nm = %d", op->nm[0]); return NEKMF_ERR_UNSUPPORTED;
    }
}
void prism_helm_fused_free(void *state, int nm)
{
    switch (nm)
    {
        case 5: delete static_cast<PrismHelmState<5> *>(state); break;
        case 6: delete static_cast<PrismHelmState<6> *>(state); break;
        case 7: delete static_cast<PrismHelmState<7> *>(state); break;
        default: break;
    }
}

} // namespace nekmf
