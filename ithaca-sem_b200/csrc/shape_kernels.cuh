// shape_kernels.cuh -- compile-time sized sum-factorised operators for Quad, Tri, Prism, Pyr and Tet
// (default quadrature nq0 = nm+1; Gauss-Radau directions nq = nm) on sm_100a.
//
// Reference semantics (what is computed, not how):
//   BwdTrans         MatrixFreeOps/BwdTransKernels.hpp:35-76 (Quad) 78-126 (Tri) 128-224 (Pyr) 226-300 (Prism) 374-484 (Tet)
//   IProductWRTBase  MatrixFreeOps/IProductKernels.hpp:76-133 (Quad) 135-234 (Tri) 316-450 (Prism) 452-598 (Pyr) 600-761 (Tet)
//   PhysDeriv        MatrixFreeOps/PhysDerivKernels.hpp:39-90,186 (2-D) 374-430 (Prism) 505-527 (Pyr) 555-696 (Tet)
//   Helmholtz        MatrixFreeOps/Helmholtz.h:138-275 (Quad) 506-635 (Tri) 1291-1458 (Prism) 1773-1950 (Pyr) 2266-2448 (Tet)
//
// Design: a persistent CTA works on batches of E elements held in shared memory.  Every 1-D contraction
// is a pencil pass: one thread owns one line of the (collapsed) tensor along the contracted direction,
// keeps it in registers and produces the whole output line.  Tensor-product directions take their matrix
// entries as constant-bank operands (kernel parameter); collapsed directions (eModified_B / eModified_C
// rows depend on the outer mode indices) read their rows from a shared-memory copy.  The singular-vertex /
// edge terms of the collapsed bases (`CORRECT` in the reference) are folded into the intermediate
// sum-factorisation arrays, so they cost a handful of FMAs per line instead of an extra pass over the
// quadrature points.  Helmholtz is fused as in hex_kernels.cuh: the last-direction derivative, the
// Laplacian metric (with the collapsed-coordinate factors h0..h3), weights and the transposed derivative
// stay in the registers of the column owner; mass and stiffness contributions are summed in quadrature
// space so ONE transposed-basis pass replaces the reference's four IProduct calls.
// Quadrature arrays use an odd i-pitch in shared memory so that the i-direction passes are bank-conflict free.
#pragma once
#include "../../include/nekmf_b200.h"
#include "common.cuh"

namespace nekmf
{

constexpr int shp_round2(int a) { return (a + 1) & ~1; }

template <int SHAPE, int NM> struct ShpDims
{
    static constexpr bool IS_QUAD = SHAPE == NEKMF_QUAD, IS_TRI = SHAPE == NEKMF_TRI, IS_PRISM = SHAPE == NEKMF_PRISM,
                          IS_TET = SHAPE == NEKMF_TET, IS_PYR = SHAPE == NEKMF_PYR;
    static constexpr int DIM = (IS_QUAD || IS_TRI) ? 2 : 3;
    static constexpr int NQ0 = NM + 1;
    static constexpr int NQ1 = (IS_TRI || IS_TET) ? NM : NM + 1;
    static constexpr int NQ2 = DIM == 2 ? 1 : NM; // Prism / Tet: Gauss-Radau in the collapsed direction
    static constexpr int NQT = NQ0 * NQ1 * NQ2;
    static constexpr int P1  = NQ0 | 1; // odd pitch of an i-line in shared memory
    static constexpr int NQP = P1 * NQ1 * NQ2;
    static constexpr int NPAIR = NM * (NM + 1) / 2;
    static constexpr int NMT   = IS_QUAD ? NM * NM : (IS_TRI ? NPAIR : (IS_PRISM ? NM * NPAIR : (IS_PYR ? NM * (NM + 1) * (2 * NM + 1) / 6 : NM * (NM + 1) * (NM + 2) / 6)));
    static constexpr int B1C_ROWS = (IS_TRI || IS_TET) ? NPAIR : 0;         // collapsed rows of direction 1
    static constexpr int B2C_ROWS = IS_PRISM ? NPAIR : ((IS_TET || IS_PYR) ? NMT : 0); // collapsed rows of direction 2
    // aux table (global -> shared): [b1c | b2c | w0 w1 w2 | h0 h1 h2 h3]
    static constexpr int NQM     = NQ0; // every per-direction helper array is padded to NQ0 entries
    static constexpr int OFF_B1C = 0;
    static constexpr int OFF_B2C = OFF_B1C + B1C_ROWS * NQ1;
    static constexpr int OFF_W   = OFF_B2C + B2C_ROWS * NQ2;
    static constexpr int OFF_H   = OFF_W + 3 * NQM;
    static constexpr int AUX_LEN = shp_round2(OFF_H + 4 * NQM);
    // lines per element in the widest pass
    static constexpr int L_JK = NQ1 * NQ2, L_IK = NQ0 * NQ2, L_IJ = DIM == 3 ? NQ0 * NQ1 : NQ0;
    static constexpr int L_MAX0 = L_JK > L_IK ? L_JK : L_IK;
    static constexpr int L_MAX  = L_MAX0 > L_IJ ? L_MAX0 : L_IJ;
    static constexpr int T      = 256;
    static constexpr int PER_ELMT = 3 * NQP + NMT + 3 * NQ2; // doubles of shared memory per element
    static constexpr int E_THR  = T / L_MAX < 1 ? 1 : T / L_MAX;
    static constexpr int E_MEM  = (96 * 1024 / 8 - AUX_LEN) / PER_ELMT < 1 ? 1 : (96 * 1024 / 8 - AUX_LEN) / PER_ELMT;
    static constexpr int E_RAW  = E_THR < E_MEM ? E_THR : E_MEM;
    static constexpr int E      = E_RAW > 32 ? 32 : E_RAW;
    static constexpr int CINSZ  = shp_round2(E * NMT);
    static constexpr int BUF    = shp_round2(E * NQP);
    static constexpr int GSZ    = shp_round2(E * 3 * NQ2);
    static constexpr size_t SMEM = (size_t)(AUX_LEN + CINSZ + 3 * BUF + GSZ) * 8 + (size_t)(4 * NPAIR + 4) * 4;
};

// constant-bank tables: only compile-time indexed entries live here
template <int SHAPE, int NM> struct ShpTab
{
    using Dm = ShpDims<SHAPE, NM>;
    double b0[NM * Dm::NQ0];          // bdata of direction 0, [p][i]
    double b1t[NM * Dm::NQ1];         // bdata of direction 1 when it is a tensor direction (Quad, Prism), [q][j]
    double D0[Dm::NQ0 * Dm::NQ0];     // D[a*nq+b] = dh_a/dz(z_b)
    double D1[Dm::NQ1 * Dm::NQ1];
    double D2[Dm::NQ2 * Dm::NQ2];
};

struct ShpArgs
{
    const double *in0;
    const double *in1 = nullptr, *in2 = nullptr; // IProductWRTDerivBase: the other components of the input field
    double *out0, *out1, *out2;
    const double *jac, *df; // already offset to the first element of this launch
    const double *aux;      // packed collapsed tables, weights, collapsed-coordinate factors (ShpDims::OFF_*)
    size_t dfStride;
    int nElmt;
    double lambda;
};

// y[b] = sum_a M[a*NOUT+b] x[a]
template <int NIN, int NOUT> __device__ __forceinline__ void shp_fwd(const double *M, const double (&x)[NIN], double (&y)[NOUT])
{
#pragma unroll
    for (int b = 0; b < NOUT; ++b)
    {
        double s = M[b] * x[0];
#pragma unroll
        for (int a = 1; a < NIN; ++a) s = fma(M[a * NOUT + b], x[a], s);
        y[b] = s;
    }
}
// y[a] = sum_b M[a*NIN+b] x[b]
template <int NIN, int NOUT> __device__ __forceinline__ void shp_tr(const double *M, const double (&x)[NIN], double (&y)[NOUT])
{
#pragma unroll
    for (int a = 0; a < NOUT; ++a)
    {
        double s = M[a * NIN] * x[0];
#pragma unroll
        for (int b = 1; b < NIN; ++b) s = fma(M[a * NIN + b], x[b], s);
        y[a] = s;
    }
}

// Laplacian metric G = (collapsed chain rule)^T df^T df (collapsed chain rule), per quadrature point.
// Formulas: Helmholtz.h:213-243 (Quad), 581-603 (Tri), 885-935 (Hex), 1382-1430 (Prism), 2360-2420 (Tet).
template <int SHAPE> struct ShpMetric;
template <> struct ShpMetric<NEKMF_QUAD>
{
    __device__ __forceinline__ static void eval(const double *f, double, double, double &m00, double &m01, double &m11)
    {
        m00 = f[0] * f[0]; m00 = fma(f[2], f[2], m00);
        m01 = f[0] * f[1]; m01 = fma(f[2], f[3], m01);
        m11 = f[1] * f[1]; m11 = fma(f[3], f[3], m11);
    }
};
template <> struct ShpMetric<NEKMF_TRI>
{
    __device__ __forceinline__ static void eval(const double *f, double h0i, double h1j, double &m00, double &m01, double &m11)
    {
        m00 = h1j * (f[0] + h0i * f[1]);
        m01 = m00 * f[1];
        m00 = m00 * m00;
        const double t = h1j * (f[2] + h0i * f[3]);
        m01 = fma(t, f[3], m01);
        m00 = fma(t, t, m00);
        m11 = f[1] * f[1]; m11 = fma(f[3], f[3], m11);
    }
};

// Helmholtz at nm >= 7 compiles to 146-254 registers, i.e. ONE resident CTA of eight warps where shared memory admits two
// (cuobjdump -res-usage): bounded to 128 registers there.  A/B (profiles/r02_sweep_shp_minb_B.jsonl against
// r02_final_sweep_*.jsonl): deformed Quad 1.57 -> 1.43, 1.92 -> 1.57, 1.93 -> 1.55 ms at nm = 7, 8, 9; deformed Prism 3.74 -> 2.59,
// 4.87 -> 2.81 ms and Pyr 4.22 -> 3.05, 5.27 -> 3.27 ms at nm = 8, 9; deformed Tet 6.84 -> 4.86 ms at nm = 9; nothing slower.
// (0 = no bound; a minimum of ONE block makes ptxas spend up to 255 registers and costs up to 1.6x)
template <int SHAPE, int OP, int NM, bool DEF>
__global__ void __launch_bounds__(256, (OP == NEKMF_HELMHOLTZ && NM >= 7) ? 2 : 0)
    shape_op_kernel(const __grid_constant__ ShpTab<SHAPE, NM> tab, const __grid_constant__ ShpArgs args)
{
    using Dm = ShpDims<SHAPE, NM>;
    constexpr int DIM = Dm::DIM, NQ0 = Dm::NQ0, NQ1 = Dm::NQ1, NQ2 = Dm::NQ2, NQT = Dm::NQT, P1 = Dm::P1, NQP = Dm::NQP;
    constexpr int NMT = Dm::NMT, NPAIR = Dm::NPAIR, E = Dm::E, T = Dm::T, NQM = Dm::NQM;
    constexpr bool IS_QUAD = Dm::IS_QUAD, IS_TRI = Dm::IS_TRI, IS_PRISM = Dm::IS_PRISM, IS_TET = Dm::IS_TET, IS_PYR = Dm::IS_PYR;
    constexpr bool COEFF_IN = OP == NEKMF_BWDTRANS || OP == NEKMF_HELMHOLTZ;
    // IProductWRTDerivBase (IProductWRTDerivBase.h:542-640 Quad, 891-1040 Tri, 1630-1740 Prism, 2484-2610 Tet):
    // dbdata = D bdata (Foundations/Basis.cpp:418-420, 506, 561), so sum_d (dB_d)^T W t_d = B^T sum_d D_d^T (W t_d):
    // a pointwise chain-rule stage feeds the transposed-derivative + IProduct half of the fused Helmholtz kernel
    constexpr bool IPWDB     = OP == NEKMF_IPRODUCTWRTDERIVBASE;
    constexpr bool HELM_BACK = OP == NEKMF_HELMHOLTZ || IPWDB;
    constexpr bool COEFF_OUT = OP == NEKMF_HELMHOLTZ || OP == NEKMF_IPRODUCTWRTBASE || IPWDB;
    constexpr int NDF = DIM * DIM;
    constexpr int LN  = NQ1 * NQ2; // (j,k) lines, ln = k*NQ1 + j
    // shared-memory pitches of the two mode-indexed intermediates, odd so that lanes owning consecutive lines hit
    // different banks: FP[p][k][j] is written / read by lanes (p, k) NQ1 doubles apart (NQ1 = 8 at nm = 7 for Quad / Prism /
    // Pyr, at nm = 8 for Tri / Tet: a whole half-warp on two banks), FPQ[pair][k] by lanes (pair) NQ2 doubles apart.  Both
    // padded layouts still fit the quadrature-sized buffers (asserted below).
    // (only where the pitch is a multiple of four doubles, i.e. 4-way conflicts or worse: the 2-way cases -- pitches 6 and
    // 10 -- measured equal or slower with the padded index arithmetic, profiles/r02_final_sweep_collapsed.jsonl history)
    constexpr int J1  = (NQ1 % 4 == 0) ? NQ1 + 1 : NQ1, LNP = NQ2 * J1;
    constexpr int K2  = (NQ2 % 4 == 0) ? NQ2 + 1 : NQ2;
    static_assert(NM * LNP <= NQP, "padded FP layout does not fit the work buffer");
    static_assert(DIM == 2 || (IS_TET ? NPAIR : NM * NM) * K2 <= NQP, "padded FPQ layout does not fit the work buffer");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sAux = reinterpret_cast<double *>(smem_raw);
    double *sCin = sAux + Dm::AUX_LEN;
    double *sU   = sCin + Dm::CINSZ;
    double *sA   = sU + Dm::BUF;
    double *sB   = sA + Dm::BUF;
    double *sG   = sB + Dm::BUF;                        // Tet: [e][3][NQ2] transposed singular-vertex sums
    int *sPQp    = reinterpret_cast<int *>(sG + Dm::GSZ); // pair index -> p
    int *sPQq    = sPQp + NPAIR;                        // pair index -> q
    int *sPQm    = sPQq + NPAIR;                        // Tet: first mode of pair (p,q)
    int *sM0     = sPQp;                                // Pyr: first mode of (p,q), NM*NM <= 3*NPAIR entries over the three arrays
    // coefficient-input operators: second coefficient buffer, filled by cp.async with the NEXT batch while this
    // one is computed (ncu: the exposed load + barrier at the top of a batch was 20 % of the prism Helmholtz kernel)
    double *sCinAlt = reinterpret_cast<double *>(smem_raw + Dm::SMEM);
    const double *b1c = sAux + Dm::OFF_B1C, *b2c = sAux + Dm::OFF_B2C;
    const double *sW0 = sAux + Dm::OFF_W, *sW1 = sW0 + NQM, *sW2 = sW1 + NQM;
    const double *sH0 = sAux + Dm::OFF_H, *sH1 = sH0 + NQM, *sH2 = sH1 + NQM, *sH3 = sH2 + NQM;

    const int tid = threadIdx.x;
    for (int i = tid; i < Dm::AUX_LEN; i += T) sAux[i] = __ldg(args.aux + i);
    if (tid == 0)
    {
        if (IS_PYR)
        {
            // mode order of the pyramid (BwdTransKernels.hpp:146-177): p outer, q, then r < NM - max(p,q)
            int m = 0;
            for (int p = 0; p < NM; ++p)
                for (int q = 0; q < NM; ++q)
                {
                    sM0[p * NM + q] = m;
                    m += NM - (p > q ? p : q);
                }
        }
        else
        {
            int c = 0, m = 0;
            for (int p = 0; p < NM; ++p)
                for (int q = 0; q < NM - p; ++q, ++c)
                {
                    sPQp[c] = p;
                    sPQq[c] = q;
                    sPQm[c] = m;
                    m += NM - p - q;
                }
        }
    }
    __syncthreads();

    const int nElmt    = args.nElmt;
    const int nBatches = (nElmt + E - 1) / E;
    // first pair / (p,r) row of outer index p: p*NM - p(p-1)/2
    auto tri0 = [](int p) { return p * NM - (p * (p - 1)) / 2; };

    auto prefetch = [&](int b, double *dst) {
        const int e0 = b * E, ne = nElmt - e0 < E ? nElmt - e0 : E;
        const double *src = args.in0 + (size_t)e0 * NMT;
        for (int i = tid; i < ne * NMT; i += T)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst + i)), "l"(src + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (COEFF_IN && (int)blockIdx.x < nBatches) prefetch(blockIdx.x, sCin);
    // regular IProductWRTBase: the quadrature values of the NEXT batch are copied (raw) into sU as soon as the first
    // transposed pass has consumed it; Jacobian and weights are applied when that pass reads them
    constexpr bool IP_PREF = OP == NEKMF_IPRODUCTWRTBASE && !DEF;
    // PhysDeriv: a second quadrature buffer behind the common layout (same place as sCinAlt) takes the next batch
    constexpr bool PD_PREF = OP == NEKMF_PHYSDERIV;
    double *sUAlt          = reinterpret_cast<double *>(smem_raw + Dm::SMEM);
    auto prefetch_phys = [&](int b, double *dstU) {
        const int e0 = b * E, ne = nElmt - e0 < E ? nElmt - e0 : E;
        const double *src = args.in0 + (size_t)e0 * NQT;
        for (int g = tid; g < ne * NQT; g += T)
        {
            const int e = g / NQT, r = g - e * NQT;
            const int line = r / NQ0, i = r - line * NQ0;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dstU + e * NQP + line * P1 + i)), "l"(src + g)
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if ((IP_PREF || PD_PREF) && (int)blockIdx.x < nBatches) prefetch_phys(blockIdx.x, sU);

    for (int b = blockIdx.x; b < nBatches; b += gridDim.x)
    {
        const int e0 = b * E;
        const int ne = nElmt - e0 < E ? nElmt - e0 : E;

        // ------------------------------------------------------------------ load
        if (COEFF_IN)
        {
            asm volatile("cp.async.wait_group 0;" ::: "memory"); // this batch (requested one batch ago) has landed
            __syncthreads();                                      // for every thread; the other buffer is free
            if (b + (int)gridDim.x < nBatches) prefetch(b + gridDim.x, sCinAlt);
        }
        else if (IP_PREF)
        {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        else if (PD_PREF)
        {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            if (b + (int)gridDim.x < nBatches) prefetch_phys(b + gridDim.x, sUAlt);
        }
        else if (IPWDB)
        {
            // in0 -> sA, in1 -> sB (3-D) | sU (2-D), in2 -> sU
            const size_t goff = (size_t)e0 * NQT;
            for (int g = tid; g < ne * NQT; g += T)
            {
                const int e = g / NQT, r = g - e * NQT;
                const int line = r / NQ0, i = r - line * NQ0;
                const int s = e * NQP + line * P1 + i;
                sA[s] = __ldg(args.in0 + goff + g);
                if (DIM == 3)
                {
                    sB[s] = __ldg(args.in1 + goff + g);
                    sU[s] = __ldg(args.in2 + goff + g);
                }
                else
                    sU[s] = __ldg(args.in1 + goff + g);
            }
        }
        else
        {
            const double *src = args.in0 + (size_t)e0 * NQT;
            for (int g = tid; g < ne * NQT; g += T)
            {
                const int e = g / NQT, r = g - e * NQT;
                const int line = r / NQ0, i = r - line * NQ0;
                double v = __ldg(src + g);
                if (OP == NEKMF_IPRODUCTWRTBASE)
                {
                    const int k = line / NQ1, j = line - k * NQ1;
                    const double jc = DEF ? __ldg(args.jac + (size_t)e0 * NQT + g) : __ldg(args.jac + e0 + e);
                    double w = sW0[i] * sW1[j];
                    if (DIM == 3) w *= sW2[k];
                    v *= jc * w;
                }
                sU[e * NQP + line * P1 + i] = v;
            }
        }
        __syncthreads();

        if (COEFF_IN)
        {
            // -------------------------------------------------------------- S1: r -> k (3-D)
            if (IS_TET)
            {
                // lines (e, pair c): FPQ[c][k] = sum_r in[m0+r] b2[m0+r][k]     -> sA
                for (int l = tid; l < E * NPAIR; l += T)
                {
                    const int e = l / NPAIR, c = l - e * NPAIR;
                    const int m0 = sPQm[c], len = NM - sPQp[c] - sPQq[c];
                    const double *cin = sCin + e * NMT + m0;
                    const double *row = b2c + m0 * NQ2;
                    double y[NQ2];
#pragma unroll
                    for (int k = 0; k < NQ2; ++k) y[k] = 0.0;
                    for (int r = 0; r < len; ++r)
                    {
                        const double x = cin[r];
#pragma unroll
                        for (int k = 0; k < NQ2; ++k) y[k] = fma(row[r * NQ2 + k], x, y[k]);
                    }
#pragma unroll
                    for (int k = 0; k < NQ2; ++k) sA[e * NQP + c * K2 + k] = y[k];
                }
                __syncthreads();
            }
            if (IS_PRISM)
            {
                // lines (e, p, q): FPQ[p][q][k] = sum_r in[m0+r] b2[(p,r)][k]  (+ singular edge folded into p = 1)
                for (int l = tid; l < E * NM * NM; l += T)
                {
                    const int e = l / (NM * NM), pq = l - e * (NM * NM);
                    const int p = pq / NM, q = pq - p * NM;
                    const int r0 = tri0(p), len = NM - p;
                    const double *cin = sCin + e * NMT + NM * r0 + q * len;
                    const double *row = b2c + r0 * NQ2;
                    double y[NQ2];
#pragma unroll
                    for (int k = 0; k < NQ2; ++k) y[k] = 0.0;
                    for (int r = 0; r < len; ++r)
                    {
                        const double x = cin[r];
#pragma unroll
                        for (int k = 0; k < NQ2; ++k) y[k] = fma(row[r * NQ2 + k], x, y[k]);
                    }
                    if (p == 1)
                    {
                        const double x = sCin[e * NMT + q * NM + 1];
#pragma unroll
                        for (int k = 0; k < NQ2; ++k) y[k] = fma(b2c[NQ2 + k], x, y[k]);
                    }
#pragma unroll
                    for (int k = 0; k < NQ2; ++k) sA[e * NQP + pq * K2 + k] = y[k];
                }
                __syncthreads();
            }
            if (IS_PYR)
            {
                // lines (e, p, q): FPQ[p][q][k] = sum_{r < NM - max(p,q)} in[m0+r] b2[m0+r][k]; the top-vertex term of
                // BwdTransKernels.hpp:204-221 is in[1] b2[1][k] (b0[0] b1[1] + b0[1] b1[0] + b0[1] b1[1]): added to the
                // entries (0,1), (1,0), (1,1) it passes through the two tensor contractions below
                for (int l = tid; l < E * NM * NM; l += T)
                {
                    const int e = l / (NM * NM), pq = l - e * (NM * NM);
                    const int p = pq / NM, q = pq - p * NM;
                    const int m0 = sM0[pq], len = NM - (p > q ? p : q);
                    const double *cin = sCin + e * NMT + m0;
                    const double *row = b2c + m0 * NQ2;
                    double y[NQ2];
#pragma unroll
                    for (int k = 0; k < NQ2; ++k) y[k] = 0.0;
                    for (int r = 0; r < len; ++r)
                    {
                        const double x = cin[r];
#pragma unroll
                        for (int k = 0; k < NQ2; ++k) y[k] = fma(row[r * NQ2 + k], x, y[k]);
                    }
                    if (p < 2 && q < 2 && p + q > 0)
                    {
                        const double x = sCin[e * NMT + 1];
#pragma unroll
                        for (int k = 0; k < NQ2; ++k) y[k] = fma(b2c[NQ2 + k], x, y[k]);
                    }
#pragma unroll
                    for (int k = 0; k < NQ2; ++k) sA[e * NQP + pq * K2 + k] = y[k];
                }
                __syncthreads();
            }
            // -------------------------------------------------------------- S2: q -> j      FP[p][k][j] -> sB
            for (int l = tid; l < E * NM * NQ2; l += T)
            {
                const int e = l / (NM * NQ2), pk = l - e * (NM * NQ2);
                const int p = pk / NQ2, k = pk - p * NQ2;
                double y[NQ1];
                if (IS_QUAD)
                {
                    double x[NM];
#pragma unroll
                    for (int q = 0; q < NM; ++q) x[q] = sCin[e * NMT + q * NM + p];
                    shp_fwd<NM, NQ1>(tab.b1t, x, y);
                }
                else if (IS_PRISM || IS_PYR)
                {
                    double x[NM];
#pragma unroll
                    for (int q = 0; q < NM; ++q) x[q] = sA[e * NQP + (p * NM + q) * K2 + k];
                    shp_fwd<NM, NQ1>(tab.b1t, x, y);
                }
                else
                {
                    // Tri / Tet: rows c0 .. c0+len of the eModified_B table
                    const int c0 = tri0(p), len = NM - p;
                    const double *row = b1c + c0 * NQ1;
#pragma unroll
                    for (int j = 0; j < NQ1; ++j) y[j] = 0.0;
                    for (int q = 0; q < len; ++q)
                    {
                        const double x = IS_TRI ? sCin[e * NMT + c0 + q] : sA[e * NQP + (c0 + q) * K2 + k];
#pragma unroll
                        for (int j = 0; j < NQ1; ++j) y[j] = fma(row[q * NQ1 + j], x, y[j]);
                    }
                    if (IS_TRI && p == 1)
                    {
                        const double x = sCin[e * NMT + 1];
#pragma unroll
                        for (int j = 0; j < NQ1; ++j) y[j] = fma(b1c[NQ1 + j], x, y[j]);
                    }
                    if (IS_TET && p < 2)
                    {
                        const double *cin = sCin + e * NMT;
                        const double c1   = b2c[NQ2 + k] * cin[1];
                        if (p == 0)
                        {
#pragma unroll
                            for (int j = 0; j < NQ1; ++j) y[j] = fma(b1c[NQ1 + j], c1, y[j]);
                        }
                        else
                        {
                            double c2 = b2c[k] * cin[NM];
                            for (int r = 1; r < NM - 1; ++r) c2 = fma(b2c[(r + 1) * NQ2 + k], cin[NM + r], c2);
                            const double c12 = c1 + c2;
#pragma unroll
                            for (int j = 0; j < NQ1; ++j) y[j] = fma(b1c[NQ1 + j], c12, fma(b1c[j], c1, y[j]));
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < NQ1; ++j) sB[e * NQP + p * LNP + k * J1 + j] = y[j];
            }
            __syncthreads();
            // -------------------------------------------------------------- S3: p -> i      u -> sU (Helmholtz: du/dxi0 -> sA)
            for (int l = tid; l < E * LN; l += T)
            {
                const int e = l / LN, ln = l - e * LN;
                const int lnp = ln + (ln / NQ1) * (J1 - NQ1); // (k, j) in the padded [p][k][J1] layout
                double x[NM], y[NQ0];
#pragma unroll
                for (int p = 0; p < NM; ++p) x[p] = sB[e * NQP + p * LNP + lnp];
                shp_fwd<NM, NQ0>(tab.b0, x, y);
#pragma unroll
                for (int i = 0; i < NQ0; ++i) sU[e * NQP + ln * P1 + i] = y[i];
                if (OP == NEKMF_HELMHOLTZ)
                {
                    double d[NQ0];
                    shp_fwd<NQ0, NQ0>(tab.D0, y, d);
#pragma unroll
                    for (int i = 0; i < NQ0; ++i) sA[e * NQP + ln * P1 + i] = d[i];
                }
            }
            __syncthreads();
        }

        if (OP == NEKMF_BWDTRANS)
        {
            double *dst = args.out0 + (size_t)e0 * NQT;
            for (int g = tid; g < ne * NQT; g += T)
            {
                const int e = g / NQT, r = g - e * NQT;
                const int line = r / NQ0, i = r - line * NQ0;
                dst[g] = sU[e * NQP + line * P1 + i];
            }
            __syncthreads();
            {
                double *t = sCin; sCin = sCinAlt; sCinAlt = t;
            }
            continue;
        }

        if (OP == NEKMF_PHYSDERIV)
        {
            // du/dxi0: lines (j,k) along i -> sA
            for (int l = tid; l < E * LN; l += T)
            {
                const int e = l / LN, ln = l - e * LN;
                double x[NQ0], d[NQ0];
#pragma unroll
                for (int i = 0; i < NQ0; ++i) x[i] = sU[e * NQP + ln * P1 + i];
                shp_fwd<NQ0, NQ0>(tab.D0, x, d);
#pragma unroll
                for (int i = 0; i < NQ0; ++i) sA[e * NQP + ln * P1 + i] = d[i];
            }
        }
        if (DIM == 3 && (OP == NEKMF_HELMHOLTZ || OP == NEKMF_PHYSDERIV))
        {
            // du/dxi1: lines (i,k) along j -> sB
            for (int l = tid; l < E * NQ0 * NQ2; l += T)
            {
                const int e = l / (NQ0 * NQ2), ik = l - e * (NQ0 * NQ2);
                const int k = ik / NQ0, i = ik - k * NQ0;
                const int base = e * NQP + k * NQ1 * P1 + i;
                double x[NQ1], d[NQ1];
#pragma unroll
                for (int j = 0; j < NQ1; ++j) x[j] = sU[base + j * P1];
                shp_fwd<NQ1, NQ1>(tab.D1, x, d);
#pragma unroll
                for (int j = 0; j < NQ1; ++j) sB[base + j * P1] = d[j];
            }
        }
        if (OP == NEKMF_HELMHOLTZ || OP == NEKMF_PHYSDERIV) __syncthreads();

        // ------------------------------------------------------------------ column pass along the last direction
        if (OP == NEKMF_HELMHOLTZ || OP == NEKMF_PHYSDERIV || IPWDB)
        {
            constexpr int NL = DIM == 3 ? NQ2 : NQ1;          // points along the column
            constexpr int LS = DIM == 3 ? NQ1 * P1 : P1;      // shared-memory stride along the column
            constexpr int GS = DIM == 3 ? NQ1 * NQ0 : NQ0;    // dense (global) stride along the column
            constexpr int NC = DIM == 3 ? NQ0 * NQ1 : NQ0;    // columns per element
            for (int l = tid; l < E * NC; l += T)
            {
                const int e = l / NC, c = l - e * NC;
                const int j = DIM == 3 ? c / NQ0 : 0, i = DIM == 3 ? c - j * NQ0 : c;
                const int col  = e * NQP + (DIM == 3 ? j * P1 + i : i);
                const int gcol = DIM == 3 ? j * NQ0 + i : i;
                const bool ev  = e < ne;
                double rdf[NDF], rjac = 0.0;
#pragma unroll
                for (int n = 0; n < NDF; ++n) rdf[n] = 0.0;
                const size_t gbase = DEF ? (size_t)(e0 + e) * NQT + gcol : (size_t)(e0 + e);
                if (!DEF && ev)
                {
#pragma unroll
                    for (int n = 0; n < NDF; ++n) rdf[n] = __ldg(args.df + (size_t)n * args.dfStride + gbase);
                    if (HELM_BACK) rjac = __ldg(args.jac + gbase);
                }
                const double h0i = sH0[i];
                double u[NL], dl[NL];
#pragma unroll
                for (int s = 0; s < NL; ++s) u[s] = sU[col + s * LS];
                if constexpr (!IPWDB)
                {
                    if (DIM == 3) shp_fwd<NL, NL>(tab.D2, u, dl);
                    else shp_fwd<NL, NL>(tab.D1, u, dl);
                }

                if constexpr (IPWDB)
                {
                    // t_d = sum_c df[c*dim+d] in_c, collapsed-coordinate factors, Jacobian and weights; the last
                    // direction's transposed derivative is taken here, the other two by the passes below
                    const double wij = DIM == 3 ? sW0[i] * sW1[j] : sW0[i];
                    double vl[NL];
#pragma unroll
                    for (int s = 0; s < NL; ++s)
                    {
                        const int pt = col + s * LS;
                        double f[NDF], jc;
#pragma unroll
                        for (int n = 0; n < NDF; ++n)
                            f[n] = DEF ? (ev ? __ldg(args.df + (size_t)n * args.dfStride + gbase + s * GS) : 0.0) : rdf[n];
                        jc = DEF ? (ev ? __ldg(args.jac + gbase + s * GS) : 0.0) : rjac;
                        const double jw = jc * (wij * (DIM == 3 ? sW2[s] : sW1[s]));
                        if constexpr (DIM == 2)
                        {
                            const double x = sA[pt], y = u[s];
                            double t0 = fma(f[2], y, f[0] * x);
                            const double t1 = fma(f[3], y, f[1] * x);
                            if (IS_TRI) t0 = fma(h0i, t1, t0) * sH1[s]; // IProductWRTDerivBase.h:1006-1036
                            sA[pt] = jw * t0;
                            vl[s]  = jw * t1;
                        }
                        else
                        {
                            const double x = sA[pt], y = sB[pt], z = u[s];
                            double t0 = fma(f[6], z, fma(f[3], y, f[0] * x));
                            double t1 = fma(f[7], z, fma(f[4], y, f[1] * x));
                            const double t2 = fma(f[8], z, fma(f[5], y, f[2] * x));
                            if (IS_PRISM) t0 = fma(h0i, t2, t0) * sH1[s]; // IProductWRTDerivBase.h:1697-1733
                            if (IS_PYR)
                            {
                                // IProductWRTDerivBase.h:2121-2164: both base directions collapse towards the apex
                                t0 = fma(h0i, t2, t0) * sH1[s];
                                t1 = fma(sH2[j], t2, t1) * sH1[s];
                            }
                            if (IS_TET)
                            {
                                // IProductWRTDerivBase.h:2551-2603
                                const double f2 = sH3[s];
                                t0 = fma(t1 + t2, h0i, t0) * (f2 * sH2[j]);
                                t1 = fma(t2, sH1[j], t1) * f2;
                            }
                            sA[pt] = jw * t0;
                            sB[pt] = jw * t1;
                            vl[s]  = jw * t2;
                        }
                    }
                    double t[NL];
                    if (DIM == 3) shp_tr<NL, NL>(tab.D2, vl, t);
                    else shp_tr<NL, NL>(tab.D1, vl, t);
#pragma unroll
                    for (int s = 0; s < NL; ++s) sU[col + s * LS] = t[s];
                }
                else if (OP == NEKMF_PHYSDERIV)
                {
#pragma unroll
                    for (int s = 0; s < NL; ++s)
                    {
                        const int pt = col + s * LS;
                        double f[NDF];
#pragma unroll
                        for (int n = 0; n < NDF; ++n)
                            f[n] = DEF ? (ev ? __ldg(args.df + (size_t)n * args.dfStride + gbase + s * GS) : 0.0) : rdf[n];
                        if constexpr (DIM == 2)
                        {
                            double d0 = sA[pt], d1 = dl[s];
                            if (IS_TRI)
                            {
                                d0 = sH1[s] * d0;      // 2/(1-z1_j), s == j
                                d1 = fma(d0, h0i, d1); // + d0 * (1+z0_i)/2
                            }
                            sA[pt] = fma(d1, f[1], d0 * f[0]);
                            sU[pt] = fma(d1, f[3], d0 * f[2]);
                        }
                        else
                        {
                            double d0 = sA[pt], d1 = sB[pt], d2 = dl[s];
                            if (IS_PRISM)
                            {
                                d0 = d0 * sH1[s];      // 2/(1-z2_k), s == k
                                d2 = fma(h0i, d0, d2);
                            }
                            if (Dm::IS_PYR)
                            {
                                // PhysDerivKernels.hpp:505-527: both base directions collapse towards the apex
                                d0 = d0 * sH1[s];      // 2/(1-z2_k)
                                d1 = d1 * sH1[s];
                                d2 = fma(h0i, d0, d2);   // + d0 (1+z0_i)/2
                                d2 = fma(sH2[j], d1, d2); // + d1 (1+z1_j)/2
                            }
                            if (IS_TET)
                            {
                                // PhysDerivKernels.hpp:595-694
                                const double x2 = sH3[s], x1 = sH2[j];
                                d0              = (x1 * x2) * d0;
                                const double a  = h0i * d0;
                                const double d1s = x2 * d1;
                                d1              = a + d1s;
                                d2              = fma(d1s, sH1[j], a) + d2;
                            }
                            sA[pt] = fma(d2, f[2], fma(d1, f[1], d0 * f[0]));
                            sB[pt] = fma(d2, f[5], fma(d1, f[4], d0 * f[3]));
                            sU[pt] = fma(d2, f[8], fma(d1, f[7], d0 * f[6]));
                        }
                    }
                }
                else
                {
                    const double wij = DIM == 3 ? sW0[i] * sW1[j] : sW0[i];
                    double vl[NL], acc[NL];
#pragma unroll
                    for (int s = 0; s < NL; ++s)
                    {
                        const int pt = col + s * LS;
                        double f[NDF], jc;
#pragma unroll
                        for (int n = 0; n < NDF; ++n)
                            f[n] = DEF ? (ev ? __ldg(args.df + (size_t)n * args.dfStride + gbase + s * GS) : 0.0) : rdf[n];
                        jc = DEF ? (ev ? __ldg(args.jac + gbase + s * GS) : 0.0) : rjac;
                        const double jw = jc * (wij * (DIM == 3 ? sW2[s] : sW1[s]));
                        if constexpr (DIM == 2)
                        {
                            double m00, m01, m11;
                            ShpMetric<IS_QUAD ? NEKMF_QUAD : NEKMF_TRI>::eval(f, h0i, sH1[s], m00, m01, m11);
                            const double d0 = sA[pt], d1 = dl[s];
                            sA[pt] = jw * fma(m01, d1, m00 * d0);
                            vl[s]  = jw * fma(m11, d1, m01 * d0);
                        }
                        else
                        {
                            double m00, m01, m02, m11, m12, m22;
                            if (IS_PRISM)
                            {
                                const double h1 = sH1[s];
                                const double t1 = h1 * fma(h0i, f[2], f[0]), t2 = h1 * fma(h0i, f[5], f[3]),
                                             t3 = h1 * fma(h0i, f[8], f[6]);
                                m00 = fma(t3, t3, fma(t2, t2, t1 * t1));
                                m01 = fma(f[7], t3, fma(f[4], t2, f[1] * t1));
                                m02 = fma(f[8], t3, fma(f[5], t2, f[2] * t1));
                                m11 = fma(f[7], f[7], fma(f[4], f[4], f[1] * f[1]));
                                m22 = fma(f[8], f[8], fma(f[5], f[5], f[2] * f[2]));
                                m12 = fma(f[7], f[8], fma(f[4], f[5], f[1] * f[2]));
                            }
                            else if (IS_PYR)
                            {
                                // Helmholtz.h:1842-1902
                                const double a = sH1[s], h1j = sH2[j];
                                const double t0 = a * fma(h0i, f[2], f[0]), t1 = a * fma(h0i, f[5], f[3]),
                                             t2 = a * fma(h0i, f[8], f[6]);
                                const double t3 = a * fma(h1j, f[2], f[1]), t4 = a * fma(h1j, f[5], f[4]),
                                             t5 = a * fma(h1j, f[8], f[7]);
                                m00 = fma(t2, t2, fma(t1, t1, t0 * t0));
                                m11 = fma(t5, t5, fma(t4, t4, t3 * t3));
                                m22 = fma(f[8], f[8], fma(f[5], f[5], f[2] * f[2]));
                                m01 = fma(t2, t5, fma(t1, t4, t0 * t3));
                                m02 = fma(f[8], t2, fma(f[5], t1, f[2] * t0));
                                m12 = fma(f[8], t5, fma(f[5], t4, f[2] * t3));
                            }
                            else
                            {
                                const double h3 = sH3[s], h1 = sH1[j], h2 = sH2[j];
                                const double h2h3 = h2 * h3, h1h3 = h1 * h3, h0h2h3 = h0i * h2h3;
                                const double t1 = fma(f[0], h2h3, h0h2h3 * (f[1] + f[2]));
                                const double t2 = fma(f[3], h2h3, h0h2h3 * (f[4] + f[5]));
                                const double t3 = fma(f[6], h2h3, h0h2h3 * (f[7] + f[8]));
                                m00 = fma(t3, t3, fma(t2, t2, t1 * t1));
                                m02 = fma(f[8], t3, fma(f[5], t2, f[2] * t1));
                                const double t4 = fma(f[2], h1h3, f[1] * h3);
                                const double t5 = fma(f[5], h1h3, f[4] * h3);
                                const double t6 = fma(f[8], h1h3, f[7] * h3);
                                m01 = fma(t3, t6, fma(t2, t5, t1 * t4));
                                m11 = fma(t6, t6, fma(t5, t5, t4 * t4));
                                m12 = fma(f[8], t6, fma(f[5], t5, f[2] * t4));
                                m22 = fma(f[8], f[8], fma(f[5], f[5], f[2] * f[2]));
                            }
                            const double d0 = sA[pt], d1 = sB[pt], d2 = dl[s];
                            sA[pt] = jw * fma(m02, d2, fma(m01, d1, m00 * d0));
                            sB[pt] = jw * fma(m12, d2, fma(m11, d1, m01 * d0));
                            vl[s]  = jw * fma(m22, d2, fma(m12, d1, m02 * d0));
                        }
                        acc[s] = (args.lambda * jw) * u[s];
                    }
                    double t[NL];
                    if (DIM == 3) shp_tr<NL, NL>(tab.D2, vl, t);
                    else shp_tr<NL, NL>(tab.D1, vl, t);
#pragma unroll
                    for (int s = 0; s < NL; ++s) sU[col + s * LS] = acc[s] + t[s];
                }
            }
            __syncthreads();
        }

        if (OP == NEKMF_PHYSDERIV)
        {
            // out0 <- sA, out1 <- sB (3-D) | sU (2-D), out2 <- sU
            const size_t goff = (size_t)e0 * NQT;
            for (int g = tid; g < ne * NQT; g += T)
            {
                const int e = g / NQT, r = g - e * NQT;
                const int line = r / NQ0, i = r - line * NQ0;
                const int s = e * NQP + line * P1 + i;
                args.out0[goff + g] = sA[s];
                if (DIM == 3)
                {
                    args.out1[goff + g] = sB[s];
                    args.out2[goff + g] = sU[s];
                }
                else
                    args.out1[goff + g] = sU[s];
            }
            __syncthreads();
            {
                double *t = sU; sU = sUAlt; sUAlt = t;
            }
            continue;
        }

        // ------------------------------------------------------------------ transposed passes (Helmholtz, IProduct)
        if (HELM_BACK && DIM == 3)
        {
            // sU += D1^T sB: lines (i,k) along j
            for (int l = tid; l < E * NQ0 * NQ2; l += T)
            {
                const int e = l / (NQ0 * NQ2), ik = l - e * (NQ0 * NQ2);
                const int k = ik / NQ0, i = ik - k * NQ0;
                const int base = e * NQP + k * NQ1 * P1 + i;
                double x[NQ1], t[NQ1];
#pragma unroll
                for (int j = 0; j < NQ1; ++j) x[j] = sB[base + j * P1];
                shp_tr<NQ1, NQ1>(tab.D1, x, t);
#pragma unroll
                for (int j = 0; j < NQ1; ++j) sU[base + j * P1] += t[j];
            }
            __syncthreads();
        }
        // T3: i -> p, lines (j,k): f[p][ln] -> sB
        for (int l = tid; l < E * LN; l += T)
        {
            const int e = l / LN, ln = l - e * LN;
            double v[NQ0], f[NM];
#pragma unroll
            for (int i = 0; i < NQ0; ++i) v[i] = sU[e * NQP + ln * P1 + i];
            if (IP_PREF)
            {
                const int k = ln / NQ1, j = ln - k * NQ1;
                double jwl = (e < ne ? __ldg(args.jac + e0 + e) : 0.0) * sW1[j];
                if (DIM == 3) jwl *= sW2[k];
#pragma unroll
                for (int i = 0; i < NQ0; ++i) v[i] *= jwl * sW0[i];
            }
            if (HELM_BACK)
            {
                double a[NQ0], t[NQ0];
#pragma unroll
                for (int i = 0; i < NQ0; ++i) a[i] = sA[e * NQP + ln * P1 + i];
                shp_tr<NQ0, NQ0>(tab.D0, a, t);
#pragma unroll
                for (int i = 0; i < NQ0; ++i) v[i] += t[i];
            }
            shp_tr<NQ0, NM>(tab.b0, v, f);
#pragma unroll
            for (int p = 0; p < NM; ++p) sB[e * NQP + p * LNP + ln + (ln / NQ1) * (J1 - NQ1)] = f[p];
        }
        __syncthreads();
        if (IP_PREF && b + (int)gridDim.x < nBatches) prefetch_phys(b + gridDim.x, sU);
        // T2: j -> q, lines (p,k)
        for (int l = tid; l < E * NM * NQ2; l += T)
        {
            const int e = l / (NM * NQ2), pk = l - e * (NM * NQ2);
            const int p = pk / NQ2, k = pk - p * NQ2;
            double x[NQ1];
#pragma unroll
            for (int j = 0; j < NQ1; ++j) x[j] = sB[e * NQP + p * LNP + k * J1 + j];
            if (IS_QUAD)
            {
                double y[NM];
                shp_tr<NQ1, NM>(tab.b1t, x, y);
#pragma unroll
                for (int q = 0; q < NM; ++q) sCin[e * NMT + q * NM + p] = y[q];
            }
            else if (IS_PRISM || IS_PYR)
            {
                double y[NM];
                shp_tr<NQ1, NM>(tab.b1t, x, y);
#pragma unroll
                for (int q = 0; q < NM; ++q) sA[e * NQP + (p * NM + q) * K2 + k] = y[q];
            }
            else
            {
                const int c0 = tri0(p), len = NM - p;
                const double *row = b1c + c0 * NQ1;
                for (int q = 0; q < len; ++q)
                {
                    double s = row[q * NQ1] * x[0];
#pragma unroll
                    for (int j = 1; j < NQ1; ++j) s = fma(row[q * NQ1 + j], x[j], s);
                    if (IS_TRI) sCin[e * NMT + c0 + q] = s;
                    else sA[e * NQP + (c0 + q) * K2 + k] = s;
                }
                if (IS_TET && p < 2)
                {
                    // transposed singular-vertex sums (IProductKernels.hpp:713-759)
                    double g1 = b1c[NQ1] * x[0], g0 = b1c[0] * x[0];
#pragma unroll
                    for (int j = 1; j < NQ1; ++j)
                    {
                        g1 = fma(b1c[NQ1 + j], x[j], g1);
                        g0 = fma(b1c[j], x[j], g0);
                    }
                    if (p == 0) sG[(e * 3 + 0) * NQ2 + k] = g1; // sum_j b1[1][j] f[0][j][k]
                    else
                    {
                        sG[(e * 3 + 1) * NQ2 + k] = g0;         // sum_j b1[0][j] f[1][j][k]
                        sG[(e * 3 + 2) * NQ2 + k] = g1;         // sum_j b1[1][j] f[1][j][k]
                    }
                }
            }
        }
        __syncthreads();
        if (IS_TRI)
        {
            // singular vertex: out[1] += sum_j b1[1][j] f[1][j]
            for (int e = tid; e < E; e += T)
            {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < NQ1; ++j) s = fma(b1c[NQ1 + j], sB[e * NQP + 1 * LNP + j], s);
                sCin[e * NMT + 1] += s;
            }
            __syncthreads();
        }
        // T1: k -> r (3-D)
        if (IS_TET)
        {
            for (int l = tid; l < E * NPAIR; l += T)
            {
                const int e = l / NPAIR, c = l - e * NPAIR;
                const int m0 = sPQm[c], len = NM - sPQp[c] - sPQq[c];
                const double *row = b2c + m0 * NQ2;
                double x[NQ2];
#pragma unroll
                for (int k = 0; k < NQ2; ++k) x[k] = sA[e * NQP + c * K2 + k];
                for (int r = 0; r < len; ++r)
                {
                    double s = row[r * NQ2] * x[0];
#pragma unroll
                    for (int k = 1; k < NQ2; ++k) s = fma(row[r * NQ2 + k], x[k], s);
                    sCin[e * NMT + m0 + r] = s;
                }
                if (c == 0)
                {
                    // mode 1 = (0,0,1): top vertex
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NQ2; ++k)
                        s = fma(b2c[NQ2 + k], sG[(e * 3 + 0) * NQ2 + k] + sG[(e * 3 + 1) * NQ2 + k] + sG[(e * 3 + 2) * NQ2 + k], s);
                    sCin[e * NMT + 1] += s;
                }
                if (c == 1)
                {
                    // modes NM + r = (0,1,r): bottom vertex (r = 0, b2 row 0) and singular edge (b2 row r+1)
                    for (int r = 0; r < NM - 1; ++r)
                    {
                        const double *rw = b2c + (r == 0 ? 0 : r + 1) * NQ2;
                        double s = 0.0;
#pragma unroll
                        for (int k = 0; k < NQ2; ++k) s = fma(rw[k], sG[(e * 3 + 2) * NQ2 + k], s);
                        sCin[e * NMT + NM + r] += s;
                    }
                }
            }
            __syncthreads();
        }
        if (IS_PRISM)
        {
            for (int l = tid; l < E * NM * NM; l += T)
            {
                const int e = l / (NM * NM), pq = l - e * (NM * NM);
                const int p = pq / NM, q = pq - p * NM;
                const int r0 = tri0(p), len = NM - p;
                const double *row = b2c + r0 * NQ2;
                double *cout = sCin + e * NMT + NM * r0 + q * len;
                double x[NQ2];
#pragma unroll
                for (int k = 0; k < NQ2; ++k) x[k] = sA[e * NQP + pq * K2 + k];
                for (int r = 0; r < len; ++r)
                {
                    double s = row[r * NQ2] * x[0];
#pragma unroll
                    for (int k = 1; k < NQ2; ++k) s = fma(row[r * NQ2 + k], x[k], s);
                    cout[r] = s;
                }
                if (p == 0)
                {
                    // singular edge: mode (0,q,1) += sum_k b2[(0,1)][k] fb[1][q][k]
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NQ2; ++k) s = fma(b2c[NQ2 + k], sA[e * NQP + (NM + q) * K2 + k], s);
                    cout[1] += s;
                }
            }
            __syncthreads();
        }
        if (IS_PYR)
        {
            for (int l = tid; l < E * NM * NM; l += T)
            {
                const int e = l / (NM * NM), pq = l - e * (NM * NM);
                const int p = pq / NM, q = pq - p * NM;
                const int m0 = sM0[pq], len = NM - (p > q ? p : q);
                const double *row = b2c + m0 * NQ2;
                double *cout = sCin + e * NMT + m0;
                double x[NQ2];
#pragma unroll
                for (int k = 0; k < NQ2; ++k) x[k] = sA[e * NQP + pq * K2 + k];
                for (int r = 0; r < len; ++r)
                {
                    double s = row[r * NQ2] * x[0];
#pragma unroll
                    for (int k = 1; k < NQ2; ++k) s = fma(row[r * NQ2 + k], x[k], s);
                    cout[r] = s;
                }
                if (pq == 0)
                {
                    // top vertex (IProductKernels.hpp:552-596): mode 1 += sum_k b2[1][k] (fb[0][1][k] + fb[1][0][k] + fb[1][1][k])
                    const double *fb = sA + e * NQP;
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NQ2; ++k)
                        s = fma(b2c[NQ2 + k], fb[1 * K2 + k] + fb[NM * K2 + k] + fb[(NM + 1) * K2 + k], s);
                    cout[1] += s;
                }
            }
            __syncthreads();
        }
        // ------------------------------------------------------------------ store coefficients
        if (COEFF_OUT)
        {
            double *dst = args.out0 + (size_t)e0 * NMT;
            for (int i = tid; i < ne * NMT; i += T) dst[i] = sCin[i];
            __syncthreads();
        }
        if (COEFF_IN)
        {
            double *t = sCin; sCin = sCinAlt; sCinAlt = t;
        }
    }
}

} // namespace nekmf
