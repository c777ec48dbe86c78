// hex_kernels.cuh -- sum-factorised hexahedral operators for sm_100a, one template for the five
// Collections operator types.
//
// Reference semantics (what is computed, not how):
//   BwdTrans              MatrixFreeOps/BwdTransKernels.hpp:302-372
//   IProductWRTBase       MatrixFreeOps/IProductKernels.hpp:236-314
//   PhysDeriv             MatrixFreeOps/PhysDerivKernels.hpp:219-372
//   Helmholtz             MatrixFreeOps/Helmholtz.h:764-993
//   IProductWRTDerivBase  MatrixFreeOps/IProductWRTDerivBase.h:1232-1345
//
// Design (B200-first, not the reference's lane=element SIMD):
//   * a persistent CTA processes batches of E elements; the element's coefficient block and its
//     per-quadrature-point geometric factors are staged into shared memory with TMA 1-D bulk
//     copies (cp.async.bulk + mbarrier), the next batch being requested as soon as its buffer
//     is free so the copy overlaps the remaining passes;
//   * every 1-D contraction is a "pencil pass": one thread owns one line of the tensor along the
//     contracted direction, keeps it in registers and produces the whole output line; the 1-D
//     basis / derivative matrices live in the kernel parameter (constant) bank so every DFMA
//     takes its matrix entry as a constant operand -- no shared-memory traffic for the matrices;
//   * Helmholtz is fused: derivative in the third direction, Laplacian metric, quadrature
//     weights and the transposed third-direction derivative happen in registers of the thread
//     that owns the (i,j) column; the three weak derivatives and the mass term are summed in
//     quadrature space so that ONE transposed-basis triple pass replaces the reference's four
//     IProduct calls.
#pragma once
#include "common.cuh"

namespace nekmf
{

enum HexOp
{
    HEX_BWD   = 0,
    HEX_HELM  = 1,
    HEX_IPROD = 2,
    HEX_IPWDB = 3,
    HEX_PD    = 4
};

template <int NM, int NQ> struct HexTab
{
    double B[NM * NQ]; // bdata[m*NQ+i]
    double D[NQ * NQ]; // D[k*NQ+i] = dh_k/dz(z_i)
    double w[NQ];      // quadrature weights
};

struct HexArgs
{
    const double *in0, *in1, *in2;
    double *out0, *out1, *out2;
    const double *jac; // [nElmt] | [nElmt*NQ3]
    const double *df;  // [9][nElmt] | [9][nElmt*NQ3]
    int nElmt;       // elements this launch covers (jac/df already point at the first of them)
    size_t dfStride; // distance between df rows: nElmt of the whole collection (* pitch when deformed)
    int in_aligned; // all input pointers 16-byte aligned
    double lambda;
};

template <int OP> struct HexOpTraits;
template <> struct HexOpTraits<HEX_BWD>
{
    static constexpr int NGEO = 0;
    static constexpr bool COEFF_IN = true, COEFF_OUT = false;
};
template <> struct HexOpTraits<HEX_HELM>
{
    static constexpr int NGEO = 10;
    static constexpr bool COEFF_IN = true, COEFF_OUT = true;
};
template <> struct HexOpTraits<HEX_IPROD>
{
    static constexpr int NGEO = 1;
    static constexpr bool COEFF_IN = false, COEFF_OUT = true;
};
template <> struct HexOpTraits<HEX_IPWDB>
{
    static constexpr int NGEO = 10;
    static constexpr bool COEFF_IN = false, COEFF_OUT = true;
};
template <> struct HexOpTraits<HEX_PD>
{
    static constexpr int NGEO = 9;
    static constexpr bool COEFF_IN = false, COEFF_OUT = false;
};

constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

template <int OP, int NM, int NQ, bool DEF> struct HexCfg
{
    static constexpr int NM3 = NM * NM * NM, NQ3 = NQ * NQ * NQ, NQ2 = NQ * NQ;
    static constexpr int NQ3P = round_up(NQ3, 2); // element pitch of the (internal) geometry arrays
    // deformed Helmholtz / PhysDeriv / IProductWRTDerivBase at high order: nine or ten staged factor arrays are 80-200 KB
    // per element, i.e. ONE CTA of 4-5 warps per SM.  GEO_DIRECT reads the factors in the column pass straight from
    // global memory instead (lanes = consecutive (j,i) points of a k-plane: fully coalesced), which frees the shared
    // memory for more resident CTAs.
    // Measured at nm = 9..11 (profiles/r02_sweep_hex_geodirect.jsonl against r02_final_sweep_hex.jsonl): faster for
    // Helmholtz at nm = 9 (3.44 -> 2.91 ms) and PhysDeriv / IProductWRTDerivBase at nm = 10 (1.90 -> 1.78, 3.01 -> 2.27 ms),
    // slower in the other six cells (the staged copy hides the latency better once two elements no longer fit) -- taken
    // exactly where it won.
    static constexpr bool GEO_DIRECT = DEF && ((OP == HEX_HELM && NM == 9) || ((OP == HEX_PD || OP == HEX_IPWDB) && NM == 10));
    static constexpr int NGEO = (DEF && !GEO_DIRECT) ? HexOpTraits<OP>::NGEO : 0;
    // work buffers actually live at the same time (the others alias them, see the kernel):
    //   BwdTrans: P1 -> sA, P2 -> sB, P3 -> sU = sA           IProductWRTBase: sU, sA, sB, sC = sA
    static constexpr int NBUF    = OP == HEX_BWD ? 2 : (OP == HEX_IPROD ? 3 : 4);
    // the coefficient staging block is dropped only for IProductWRTBase; PhysDeriv and IProductWRTDerivBase keep it
    // although they do not use it: with the smaller footprint more CTAs become resident and both measured
    // SLOWER (IPWDB nm=5 0.75 -> 1.18 ms, PhysDeriv nm=9 0.58 -> 0.76 ms), the pencil passes being bound by
    // shared-memory wavefronts, not by latency
    static constexpr bool HASCIN = OP != HEX_IPROD;
    // doubles of shared memory per element: work buffers + coefficient staging + geometry
    static constexpr int PER_ELMT = NBUF * NQ3 + (HASCIN ? round_up(NM3, 2) : 0) + NGEO * NQ3P;
    static constexpr int SMEM_BUDGET = 100 * 1024; // aim at >= 2 CTAs per SM
    static constexpr int E_FIT       = SMEM_BUDGET / (PER_ELMT * 8);
    // E even when possible (keeps every full batch of odd-sized coefficient blocks 16-byte aligned)
    static constexpr int E_RAW = E_FIT < 1 ? 1 : (E_FIT > 8 ? 8 : E_FIT);
    static constexpr int E_THR = (1024 / NQ2) < 1 ? 1 : (1024 / NQ2);
    // regular IProductWRTBase at nm = 11: two elements per CTA compile to 168 registers plus spills and ONE resident CTA of 9
    // warps (0.46 ms, slower than the deformed variant that streams the Jacobian on top); one element per CTA is the deformed
    // variant's shape (64 registers, four CTAs)
    static constexpr bool E_ONE = OP == HEX_IPROD && !DEF && NM == 11;
    static constexpr int E_MIN = E_ONE ? 1 : (E_RAW < E_THR ? E_RAW : E_THR);
    static constexpr int E     = (E_MIN >= 2 && (E_MIN % 2)) ? E_MIN - 1 : E_MIN;
    static constexpr int T     = round_up(E * NQ2, 32);
    static constexpr int CIN   = HASCIN ? round_up(E * NM3, 2) : 0;
    static constexpr int BUF   = round_up(E * NQ3, 2); // work-buffer pitch (keeps every buffer 16-byte aligned)
    static constexpr size_t SMEM = (size_t)(NBUF * BUF + CIN + NGEO * E * NQ3P + NQ) * 8 + 64;
    // regular-geometry Helmholtz is FP64-latency bound: the compiler hoists the per-element metric and
    // unrolls into >160 registers, which leaves ONE CTA per SM.  Bound the allocation so that two CTAs are
    // resident (measured: nm=5 1.25 -> 1.15 ms, nm=8 1.74 -> 1.18 ms, nm=11 2.42 -> 1.70 ms; the same bound
    // on IProductWRTDerivBase lost more than it won and is not applied).
    static constexpr int MINB_HELM = (!DEF && OP == HEX_HELM && 2 * SMEM <= 220 * 1024) ? 2 : 0; // 0 = no bound
    // regular PhysDeriv / IProductWRTBase / IProductWRTDerivBase at nm >= 7: ptxas settles at 147-234 registers, which
    // leaves ONE CTA of 5-9 warps per SM where shared memory has room for 2-4 (cuobjdump -res-usage of hex_nm*.o).
    // Bound the allocation to the number of CTAs shared memory admits, as long as >= 80 registers per thread remain.
    static constexpr int T_ALLOC = round_up(T, 128); // warps are allocated in fours
    static constexpr int SM_FIT  = (int)((220 * 1024) / SMEM) > 4 ? 4 : (int)((220 * 1024) / SMEM);
    static constexpr int minb_for(int m) { return m <= 1 ? 0 : ((65536 / (m * T_ALLOC) >= 80) ? m : minb_for(m - 1)); }
    // A/B over nm = 7..11 (profiles/r02_sweep_hex_minb_{A,B}.jsonl): IProductWRTDerivBase 1.16 -> 0.91, 1.23 -> 0.86, 1.11 -> 0.82,
    // 0.99 -> 0.90 ms at nm = 7, 9, 10, 11 and PhysDeriv 0.87 -> 0.68, 0.68 -> 0.61, 0.60 -> 0.53 ms at nm = 7, 10, 11; it LOSES at
    // nm = 8 (three CTAs: 96 registers, spills) and for IProductWRTBase, which keep the compiler's choice
    static constexpr bool MINB_ON = !DEF && NM >= 7 && NM != 8 && (OP == HEX_PD || OP == HEX_IPWDB);
    // deformed Helmholtz at nm = 9 with direct geometry: two CTAs (128 registers) 2.91 -> 2.61 ms; at nm = 7, 8 direct geometry
    // loses to the staged copy (1.89 -> 3.05, 1.77 -> 2.19 ms)
    static constexpr int MINB     = MINB_ON ? minb_for(SM_FIT) : ((GEO_DIRECT && OP == HEX_HELM) ? 2 : MINB_HELM);
};

// y[b] = sum_a M[a*NOUT+b] x[a]          (forward: basis / derivative evaluation)
template <int NIN, int NOUT> __device__ __forceinline__ void mat_fwd(const double *M, const double (&x)[NIN], double (&y)[NOUT])
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
// y[a] = sum_b M[a*NIN+b] x[b]           (transposed: projection onto modes / nodes)
template <int NIN, int NOUT> __device__ __forceinline__ void mat_tr(const double *M, const double (&x)[NIN], double (&y)[NOUT])
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

template <int OP, int NM, int NQ, bool DEF>
__global__ void __launch_bounds__(HexCfg<OP, NM, NQ, DEF>::T, HexCfg<OP, NM, NQ, DEF>::MINB)
    hex_op_kernel(const __grid_constant__ HexTab<NM, NQ> tab, const __grid_constant__ HexArgs args)
{
    using Cfg = HexCfg<OP, NM, NQ, DEF>;
    using Tr  = HexOpTraits<OP>;
    constexpr int E = Cfg::E, T = Cfg::T;
    constexpr int NM2 = NM * NM, NM3 = NM2 * NM, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ;
    // shared-memory pitches of the mode-indexed intermediates: lanes that own consecutive (k,j) / (r,q) lines read or
    // write them NM doubles apart, which for even NM is a 2..8-way bank conflict (NM = 8: all of a half-warp on two
    // banks; ncu-visible as the nm = 8 dip of every operator) -- an odd line pitch and a k-slab pitch that is not a
    // multiple of 16 doubles remove it; both still fit the NQ^3 work buffers (NM < NQ)
    constexpr int NMP = (NM % 2 == 0 && NM < NQ) ? NM + 1 : NM;
    constexpr int K1P = (NM2 % 16 == 0 && NM2 + 8 <= NQ2) ? NM2 + 8 : NM2;
    constexpr int COUT = NM2 * NMP; // element pitch of the staged result
    constexpr int NGEO = Cfg::NGEO, NQ3P = Cfg::NQ3P, GEOA = E * NQ3P; // GEOA: doubles per staged geometry array
    constexpr int INSZ = Tr::COEFF_IN ? NM3 : NQ3; // doubles per element of each input array
    constexpr int NIN  = OP == HEX_IPWDB ? 3 : 1;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    // BwdTrans: sU aliases sA (P3 reads sB; sA is dead after P2).  IProductWRTBase: sC aliases sA (P9 reads
    // sB; sA is dead after P8; the next write of sA is P7c of the following batch, two barriers later).
    double *sA   = reinterpret_cast<double *>(smem_raw);
    double *sB   = sA + Cfg::BUF;
    double *sU   = OP == HEX_BWD ? sA : sB + Cfg::BUF;
    double *sC   = OP == HEX_IPROD ? sA : sU + Cfg::BUF; // unused by BwdTrans
    double *sCin = reinterpret_cast<double *>(smem_raw) + Cfg::NBUF * Cfg::BUF;
    double *sGeo = sCin + Cfg::CIN;
    double *sW   = sGeo + NGEO * GEOA;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sW + NQ); // [0] inputs, [1] geometry

    const int tid      = threadIdx.x;
    const int nElmt    = args.nElmt;
    const int nBatches = (nElmt + E - 1) / E;

    if (tid < NQ) sW[tid] = tab.w[tid];
    if (tid == 0)
    {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto batch_ne = [&](int b) { int r = nElmt - b * E; return r < E ? r : E; };
    auto in_tma_ok = [&](int b) { return args.in_aligned && (((batch_ne(b) * INSZ) & 1) == 0) && (((b * E * INSZ) & 1) == 0); };
    // where input array a of a batch lands in shared memory
    auto in_dst = [&](int a) -> double * {
        if (Tr::COEFF_IN) return sCin;
        if (OP == HEX_IPWDB) return a == 0 ? sA : (a == 1 ? sB : sC);
        return sU;
    };
    auto in_src = [&](int a) -> const double * { return a == 0 ? args.in0 : (a == 1 ? args.in1 : args.in2); };
    auto issue_inputs = [&](int b) { // one thread
        if (!in_tma_ok(b)) return;
        const int ne         = batch_ne(b);
        const uint32_t bytes = (uint32_t)(ne * INSZ * 8);
        fence_proxy_async();
        mbar_expect_tx(&bars[0], bytes * NIN);
#pragma unroll
        for (int a = 0; a < NIN; ++a) tma_load_1d(in_dst(a), in_src(a) + (size_t)b * E * INSZ, bytes, &bars[0]);
    };
    auto issue_geo = [&](int b) { // one thread
        if (NGEO == 0) return;
        // internal geometry layout: [array][element][NQ3P] (op_internal.h: geo_pitch), always 16-byte aligned
        const int ne         = batch_ne(b);
        const uint32_t bytes = (uint32_t)(ne * NQ3P * 8);
        const size_t dfs     = args.dfStride;
        fence_proxy_async();
        mbar_expect_tx(&bars[1], bytes * NGEO);
        if (OP == HEX_IPROD)
            tma_load_1d(sGeo, args.jac + (size_t)b * GEOA, bytes, &bars[1]);
        else
        {
#pragma unroll
            for (int n = 0; n < 9; ++n)
                tma_load_1d(sGeo + n * GEOA, args.df + n * dfs + (size_t)b * GEOA, bytes, &bars[1]);
            if (NGEO == 10) tma_load_1d(sGeo + 9 * GEOA, args.jac + (size_t)b * GEOA, bytes, &bars[1]);
        }
    };

    uint32_t ph_in = 0, ph_geo = 0;
    if (tid == 0 && (int)blockIdx.x < nBatches)
    {
        issue_inputs(blockIdx.x);
        issue_geo(blockIdx.x);
    }

    // pencil ownership of this thread (one pencil per pass; T >= E*NQ2)
    const int pe  = tid / NQ2;       // element within the batch for NQ2-pencil passes
    const int pr  = tid - pe * NQ2;  // pencil id within the element
    const int pa  = pr / NQ;         // slow pencil coordinate
    const int pb  = pr - pa * NQ;    // fast pencil coordinate
    const bool pv = tid < E * NQ2;

    for (int b = blockIdx.x; b < nBatches; b += gridDim.x)
    {
        const int ne    = batch_ne(b);
        const int e0    = b * E;
        const int bnext = b + gridDim.x;
        // ---------------------------------------------------------------- inputs
        if (in_tma_ok(b))
        {
            mbar_wait(&bars[0], ph_in);
            ph_in ^= 1;
        }
        else
        {
#pragma unroll
            for (int a = 0; a < NIN; ++a)
            {
                const double *src = in_src(a) + (size_t)e0 * INSZ;
                double *dst       = in_dst(a);
                for (int i = tid; i < ne * INSZ; i += T) dst[i] = __ldg(src + i);
            }
            __syncthreads();
        }

        double acc[NQ]; // per-(j,i) column accumulator, lives across the D0^T / D1^T passes

        if (Tr::COEFF_IN)
        {
            // P1: r -> k.  pencils (q,p); t1[k][q][p] -> sA
            if (tid < E * NM2)
            {
                const int e = tid / NM2, qp = tid - e * NM2;
                double x[NM], y[NQ];
#pragma unroll
                for (int r = 0; r < NM; ++r) x[r] = sCin[e * NM3 + r * NM2 + qp];
                mat_fwd<NM, NQ>(tab.B, x, y);
#pragma unroll
                for (int k = 0; k < NQ; ++k) sA[e * NQ3 + k * K1P + qp] = y[k];
            }
            __syncthreads();
            if (tid == 0 && bnext < nBatches) issue_inputs(bnext); // sCin is free again
            // P2: q -> j.  pencils (k,p); t2[k][j][p] -> sB
            if (tid < E * NQ * NM)
            {
                const int e = tid / (NQ * NM), kp = tid - e * (NQ * NM);
                const int k = kp / NM, p = kp - k * NM;
                double x[NM], y[NQ];
#pragma unroll
                for (int q = 0; q < NM; ++q) x[q] = sA[e * NQ3 + k * K1P + q * NM + p];
                mat_fwd<NM, NQ>(tab.B, x, y);
#pragma unroll
                for (int j = 0; j < NQ; ++j) sB[e * NQ3 + (k * NQ + j) * NMP + p] = y[j];
            }
            __syncthreads();
            // P3: p -> i.  pencils (k,j); u[k][j][i] -> sU, and (Helmholtz) du/dxi0 -> sA
            if (pv)
            {
                double x[NM], y[NQ];
#pragma unroll
                for (int p = 0; p < NM; ++p) x[p] = sB[pe * NQ3 + pr * NMP + p];
                mat_fwd<NM, NQ>(tab.B, x, y);
#pragma unroll
                for (int i = 0; i < NQ; ++i) sU[pe * NQ3 + pr * NQ + i] = y[i];
                if (OP == HEX_HELM)
                {
                    double d[NQ];
                    mat_fwd<NQ, NQ>(tab.D, y, d);
#pragma unroll
                    for (int i = 0; i < NQ; ++i) sA[pe * NQ3 + pr * NQ + i] = d[i];
                }
            }
            __syncthreads();
        }

        if (OP == HEX_BWD)
        {
            double *dst = args.out0 + (size_t)e0 * NQ3;
            for (int i = tid; i < ne * NQ3; i += T) dst[i] = sU[i];
            __syncthreads();
            continue;
        }

        if (OP == HEX_PD)
        {
            // du/dxi0: pencils (k,j) along i
            if (pv)
            {
                double x[NQ], d[NQ];
#pragma unroll
                for (int i = 0; i < NQ; ++i) x[i] = sU[pe * NQ3 + pr * NQ + i];
                mat_fwd<NQ, NQ>(tab.D, x, d);
#pragma unroll
                for (int i = 0; i < NQ; ++i) sA[pe * NQ3 + pr * NQ + i] = d[i];
            }
        }
        if (OP == HEX_HELM || OP == HEX_PD)
        {
            // P4: du/dxi1: pencils (k,i) along j -> sB
            if (pv)
            {
                double x[NQ], d[NQ];
#pragma unroll
                for (int j = 0; j < NQ; ++j) x[j] = sU[pe * NQ3 + pa * NQ2 + j * NQ + pb];
                mat_fwd<NQ, NQ>(tab.D, x, d);
#pragma unroll
                for (int j = 0; j < NQ; ++j) sB[pe * NQ3 + pa * NQ2 + j * NQ + pb] = d[j];
            }
            __syncthreads();
        }

        // ------------------------------------------------------------ geometry arrival
        if (NGEO > 0)
        {
            mbar_wait(&bars[1], ph_geo);
            ph_geo ^= 1;
        }

        // ------------------------------------------------------------ column pass (j,i) along k
        // pa = j, pb = i
        if (pv)
        {
            const int col  = pe * NQ3 + pr;  // + k*NQ2
            const int gcol = pe * NQ3P + pr; // same point in the staged geometry arrays
            const int eg   = e0 + pe;
            const bool ev = pe < ne;
            // regular geometry: one set of factors per element
            double rdf[9], rjac = 0.0;
            if (!DEF && ev)
            {
                if (OP != HEX_IPROD)
                {
#pragma unroll
                    for (int n = 0; n < 9; ++n) rdf[n] = __ldg(args.df + (size_t)n * args.dfStride + eg);
                }
                if (OP != HEX_PD) rjac = __ldg(args.jac + eg);
            }
            const double wij = sW[pa] * sW[pb];
            // deformed factors of point gpt: staged copy, or (GEO_DIRECT) the internal [array][element][NQ3P] global layout
            const size_t gdir = (size_t)b * GEOA;
            auto gdf = [&](int n, int gpt) {
                return Cfg::GEO_DIRECT ? (ev ? __ldg(args.df + (size_t)n * args.dfStride + gdir + gpt) : 0.0) : sGeo[n * GEOA + gpt];
            };
            auto gjac = [&](int gpt) { return Cfg::GEO_DIRECT ? (ev ? __ldg(args.jac + gdir + gpt) : 0.0) : sGeo[9 * GEOA + gpt]; };

            if (OP == HEX_HELM)
            {
                double u[NQ], d2[NQ], v2[NQ];
#pragma unroll
                for (int k = 0; k < NQ; ++k) u[k] = sU[col + k * NQ2];
                mat_fwd<NQ, NQ>(tab.D, u, d2);
                double m00, m01, m02, m11, m12, m22;
                if (!DEF)
                {
                    m00 = rdf[0] * rdf[0] + rdf[3] * rdf[3] + rdf[6] * rdf[6];
                    m01 = rdf[0] * rdf[1] + rdf[3] * rdf[4] + rdf[6] * rdf[7];
                    m02 = rdf[0] * rdf[2] + rdf[3] * rdf[5] + rdf[6] * rdf[8];
                    m11 = rdf[1] * rdf[1] + rdf[4] * rdf[4] + rdf[7] * rdf[7];
                    m12 = rdf[1] * rdf[2] + rdf[4] * rdf[5] + rdf[7] * rdf[8];
                    m22 = rdf[2] * rdf[2] + rdf[5] * rdf[5] + rdf[8] * rdf[8];
                }
#pragma unroll
                for (int k = 0; k < NQ; ++k)
                {
                    const int pt = col + k * NQ2, gpt = gcol + k * NQ2;
                    double jw;
                    if (DEF)
                    {
                        double f[9];
#pragma unroll
                        for (int n = 0; n < 9; ++n) f[n] = gdf(n, gpt);
                        jw  = gjac(gpt) * (wij * tab.w[k]);
                        m00 = f[0] * f[0] + f[3] * f[3] + f[6] * f[6];
                        m01 = f[0] * f[1] + f[3] * f[4] + f[6] * f[7];
                        m02 = f[0] * f[2] + f[3] * f[5] + f[6] * f[8];
                        m11 = f[1] * f[1] + f[4] * f[4] + f[7] * f[7];
                        m12 = f[1] * f[2] + f[4] * f[5] + f[7] * f[8];
                        m22 = f[2] * f[2] + f[5] * f[5] + f[8] * f[8];
                    }
                    else
                        jw = rjac * (wij * tab.w[k]);
                    const double g0 = sA[pt], g1 = sB[pt], g2 = d2[k];
                    sA[pt] = jw * (m00 * g0 + m01 * g1 + m02 * g2);
                    sB[pt] = jw * (m01 * g0 + m11 * g1 + m12 * g2);
                    v2[k]  = jw * (m02 * g0 + m12 * g1 + m22 * g2);
                    acc[k] = (args.lambda * jw) * u[k];
                }
                double t[NQ];
                mat_tr<NQ, NQ>(tab.D, v2, t);
#pragma unroll
                for (int k = 0; k < NQ; ++k) acc[k] += t[k];
            }
            else if (OP == HEX_IPROD)
            {
#pragma unroll
                for (int k = 0; k < NQ; ++k)
                {
                    const int pt = col + k * NQ2, gpt = gcol + k * NQ2;
                    const double j = DEF ? sGeo[gpt] : rjac;
                    acc[k]         = sU[pt] * (j * (wij * tab.w[k]));
                }
            }
            else if (OP == HEX_IPWDB)
            {
                double v2[NQ];
#pragma unroll
                for (int k = 0; k < NQ; ++k)
                {
                    const int pt = col + k * NQ2, gpt = gcol + k * NQ2;
                    double f[9], jw;
                    if (DEF)
                    {
#pragma unroll
                        for (int n = 0; n < 9; ++n) f[n] = gdf(n, gpt);
                        jw = gjac(gpt) * (wij * tab.w[k]);
                    }
                    else
                    {
#pragma unroll
                        for (int n = 0; n < 9; ++n) f[n] = rdf[n];
                        jw = rjac * (wij * tab.w[k]);
                    }
                    const double a = sA[pt], bb = sB[pt], c = sC[pt];
                    sA[pt] = jw * (f[0] * a + f[3] * bb + f[6] * c);
                    sB[pt] = jw * (f[1] * a + f[4] * bb + f[7] * c);
                    v2[k]  = jw * (f[2] * a + f[5] * bb + f[8] * c);
                }
                mat_tr<NQ, NQ>(tab.D, v2, acc);
            }
            else if (OP == HEX_PD)
            {
                double u[NQ], d2[NQ];
#pragma unroll
                for (int k = 0; k < NQ; ++k) u[k] = sU[col + k * NQ2];
                mat_fwd<NQ, NQ>(tab.D, u, d2);
#pragma unroll
                for (int k = 0; k < NQ; ++k)
                {
                    const int pt = col + k * NQ2, gpt = gcol + k * NQ2;
                    double f[9];
                    if (DEF)
                    {
#pragma unroll
                        for (int n = 0; n < 9; ++n) f[n] = gdf(n, gpt);
                    }
                    else
                    {
#pragma unroll
                        for (int n = 0; n < 9; ++n) f[n] = rdf[n];
                    }
                    const double g0 = sA[pt], g1 = sB[pt], g2 = d2[k];
                    sA[pt] = f[0] * g0 + f[1] * g1 + f[2] * g2;
                    sB[pt] = f[3] * g0 + f[4] * g1 + f[5] * g2;
                    sC[pt] = f[6] * g0 + f[7] * g1 + f[8] * g2;
                }
            }
        }
        __syncthreads();
        // geometry (and phys-space inputs that are fully consumed) can be re-requested now
        if (tid == 0 && bnext < nBatches)
        {
            issue_geo(bnext);
            if (OP == HEX_PD || OP == HEX_IPROD) issue_inputs(bnext);
        }

        if (OP == HEX_PD)
        {
            double *o0 = args.out0 + (size_t)e0 * NQ3, *o1 = args.out1 + (size_t)e0 * NQ3,
                   *o2 = args.out2 + (size_t)e0 * NQ3;
            for (int i = tid; i < ne * NQ3; i += T)
            {
                o0[i] = sA[i];
                o1[i] = sB[i];
                o2[i] = sC[i];
            }
            __syncthreads();
            continue;
        }

        if (OP == HEX_HELM || OP == HEX_IPWDB)
        {
            // P7a: D0^T on v0: pencils (k,j) along i: sA -> sC
            // P7b: D1^T on v1: pencils (k,i) along j: sB -> sU
            if (pv)
            {
                double x[NQ], y[NQ];
#pragma unroll
                for (int i = 0; i < NQ; ++i) x[i] = sA[pe * NQ3 + pr * NQ + i];
                mat_tr<NQ, NQ>(tab.D, x, y);
#pragma unroll
                for (int i = 0; i < NQ; ++i) sC[pe * NQ3 + pr * NQ + i] = y[i];
#pragma unroll
                for (int j = 0; j < NQ; ++j) x[j] = sB[pe * NQ3 + pa * NQ2 + j * NQ + pb];
                mat_tr<NQ, NQ>(tab.D, x, y);
#pragma unroll
                for (int j = 0; j < NQ; ++j) sU[pe * NQ3 + pa * NQ2 + j * NQ + pb] = y[j];
            }
            __syncthreads();
        }

        // P7c: sum the weak-derivative contributions and project along k: t3[r][j][i] -> sA
        if (pv)
        {
            const int col = pe * NQ3 + pr;
            if (OP == HEX_HELM || OP == HEX_IPWDB)
            {
#pragma unroll
                for (int k = 0; k < NQ; ++k) acc[k] += sC[col + k * NQ2] + sU[col + k * NQ2];
            }
            double y[NM];
            mat_tr<NQ, NM>(tab.B, acc, y);
#pragma unroll
            for (int r = 0; r < NM; ++r) sA[col + r * NQ2] = y[r];
        }
        __syncthreads();
        // P8: j -> q.  pencils (r,i); t4[r][q][i] -> sB
        if (tid < E * NM * NQ)
        {
            const int e = tid / (NM * NQ), ri = tid - e * (NM * NQ);
            const int r = ri / NQ, i = ri - r * NQ;
            double x[NQ], y[NM];
#pragma unroll
            for (int j = 0; j < NQ; ++j) x[j] = sA[e * NQ3 + r * NQ2 + j * NQ + i];
            mat_tr<NQ, NM>(tab.B, x, y);
#pragma unroll
            for (int q = 0; q < NM; ++q) sB[e * NQ3 + r * (NM * NQ) + q * NQ + i] = y[q];
        }
        __syncthreads();
        // P9: i -> p.  pencils (r,q); out[r][q][p] -> sC (element stride NM3)
        if (tid < E * NM2)
        {
            const int e = tid / NM2, rq = tid - e * NM2;
            double x[NQ], y[NM];
#pragma unroll
            for (int i = 0; i < NQ; ++i) x[i] = sB[e * NQ3 + rq * NQ + i];
            mat_tr<NQ, NM>(tab.B, x, y);
#pragma unroll
            for (int p = 0; p < NM; ++p) sC[e * COUT + rq * NMP + p] = y[p];
        }
        __syncthreads();
        {
            double *dst = args.out0 + (size_t)e0 * NM3;
            if (NMP == NM)
            {
                for (int i = tid; i < ne * NM3; i += T) dst[i] = sC[i];
            }
            else
            {
                for (int i = tid; i < ne * NM3; i += T)
                {
                    const int rq = i / NM, p = i - rq * NM; // rq runs over (element, r, q): COUT = NM2 * NMP
                    dst[i] = sC[rq * NMP + p];
                }
            }
        }
        // sC is next written in P7a/P9 of the following batch, several barriers away -- except for
        // IProductWRTDerivBase whose next inputs land in sA/sB/sC
        if (OP == HEX_IPWDB)
        {
            __syncthreads();
            if (tid == 0 && bnext < nBatches) issue_inputs(bnext);
        }
    }
}

} // namespace nekmf
