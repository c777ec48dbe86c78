// oracle/ref_driver.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin C-ABI driver around the REFERENCE's own matrix-free kernels.  It is
// compiled (see oracle/Makefile) with -I/root/reference/library so that the
// arithmetic comes from the reference headers where they lie:
//     MatrixFreeOps/BwdTransKernels.hpp, IProductKernels.hpp,
//     PhysDerivKernels.hpp, LibUtilities/SimdLib/*, LibUtilities/Polylib/Polylib.cpp
// No reference source is copied into this repository.  The output library goes
// to oracle/_ref/ (git-ignored) and is used (a) to pin the plain-C restatement
// in oracle/mf_oracle.c, and (b) as the "reference" CPU baseline of bench.py.
//
// What cannot be included (needs Boost/Nektar infrastructure) and is therefore
// restated here, following the cited reference lines:
//   * the per-block driver loops of the operator classes
//       BwdTrans.h:712-728, IProduct.h:743-772, PhysDeriv.h (Impl loops),
//       Helmholtz.h:138-275 (Quad) 506-635 (Tri) 764-993 (Hex)
//       1291-1458 (Prism) 2266-2448 (Tet), IProductWRTDerivBase.h:1232-1345
//   * Helper<DIM>: weight scaling for Gauss-Radau points (Operator.hpp:244-258)
//   * CoalescedGeomData interleaving (CoalescedGeomData.cpp:115-198, 315-403)
//   * zero padding to the SIMD width (Collections/MatrixFreeBase.h:68-93)
#include <cstddef>
#include <cstring>
#include <vector>
#include <algorithm>

#include <LibUtilities/BasicConst/NektarUnivTypeDefs.hpp>
using Nektar::NekDouble;
#include <boost/core/ignore_unused.hpp>
#include <LibUtilities/Polylib/Polylib.h>
#include <MatrixFreeOps/BwdTransKernels.hpp>
#include <MatrixFreeOps/IProductKernels.hpp>
#include <MatrixFreeOps/PhysDerivKernels.hpp>

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace Nektar::MatrixFree;
using VecVec = std::vector<vec_t, allocator<vec_t>>;

namespace nekref
{

enum { SH_QUAD = 0, SH_TRI = 1, SH_HEX = 2, SH_PRISM = 3, SH_PYR = 4, SH_TET = 5, SH_SEG = 6 };
enum { OP_BWD = 0, OP_HELM = 1, OP_IPROD = 2, OP_IPWDB = 3, OP_PHYSDERIV = 4 };

struct Ctx
{
    int dim = 0, coordim = 0, nmTot = 0, nqTot = 0, nBlocks = 0, nElmt = 0;
    VecVec bdata[3], dbdata[3], D[3], Z[3], w[3];
    VecVec h0, h1, h2, h3;
    VecVec jac, df;
    double lambda = 1.0;
    const double *in[3] = {nullptr, nullptr, nullptr};
    double *out[3]      = {nullptr, nullptr, nullptr};
    int nthreads        = 1;
};

inline void bcast(VecVec &dst, const double *src, int n, double fac = 1.0)
{
    dst.resize(n);
    for (int i = 0; i < n; ++i) dst[i] = vec_t(fac * src[i]);
}

// CoalescedGeomData.cpp:115-198
inline void interleave_jac(VecVec &out, const double *jac, int nElmtReal, int nBlocks, int nq, bool deformed)
{
    constexpr int W = vec_t::width;
    alignas(vec_t::alignment) double tmp[W];
    if (deformed)
    {
        const long jacsize = (long)nElmtReal * nq;
        out.resize((size_t)nBlocks * nq);
        for (long b = 0; b < nBlocks; ++b)
            for (int q = 0; q < nq; ++q)
            {
                for (int j = 0; j < W; ++j)
                {
                    long idx = b * nq * W + (long)nq * j + q;
                    tmp[j]   = idx < jacsize ? jac[idx] : 0.0;
                }
                out[b * nq + q].load(tmp);
            }
    }
    else
    {
        out.resize(nBlocks);
        for (long b = 0; b < nBlocks; ++b)
        {
            for (int j = 0; j < W; ++j)
            {
                long idx = (long)W * b + j;
                tmp[j]   = idx < nElmtReal ? jac[idx] : 0.0;
            }
            out[b].load(tmp);
        }
    }
}

// CoalescedGeomData.cpp:315-403 ; df given as [ndf][nElmt*(nq|1)] row-major
inline void interleave_df(VecVec &out, const double *df, int ndf, int nElmtReal, int nBlocks, int nq, bool deformed)
{
    constexpr int W = vec_t::width;
    alignas(vec_t::alignment) double tmp[W];
    if (deformed)
    {
        const long cols = (long)nElmtReal * nq;
        out.resize((size_t)nBlocks * ndf * nq);
        size_t o = 0;
        for (long e = 0; e < nBlocks; ++e)
            for (int q = 0; q < nq; ++q)
                for (int dir = 0; dir < ndf; ++dir, ++o)
                {
                    for (int j = 0; j < W; ++j)
                    {
                        long idx = ((long)W * e + j) * nq + q;
                        tmp[j]   = idx < cols ? df[(long)dir * cols + idx] : 0.0;
                    }
                    out[o].load(tmp);
                }
    }
    else
    {
        const long cols = nElmtReal;
        out.resize((size_t)nBlocks * ndf);
        for (long e = 0; e < nBlocks; ++e)
            for (int dir = 0; dir < ndf; ++dir)
            {
                for (int j = 0; j < W; ++j)
                {
                    long idx = (long)W * e + j;
                    tmp[j]   = idx < cols ? df[(long)dir * cols + idx] : 0.0;
                }
                out[e * ndf + dir].load(tmp);
            }
    }
}

#define PAR_BLOCKS_BEGIN(c)                                                    \
    _Pragma("omp parallel num_threads((c).nthreads)")                          \
    {
#define PAR_FOR _Pragma("omp for schedule(static)")
#define PAR_BLOCKS_END }

// =============================================================== QUAD
template <int NM, int NQ, bool DEF> struct QuadOps
{
    static constexpr int nmTot = NM * NM, nqTot = NQ * NQ, ndf = 4;

    static void bwd(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        vec_t wsp[NQ * NM];
        VecVec tmpIn(nmTot), tmpOut(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            load_interleave(c.in[0] + (size_t)e * nmTot * W, nmTot, tmpIn);
            BwdTransQuadKernel<NM, NM, NQ, NQ>(tmpIn, c.bdata[0], c.bdata[1], wsp, tmpOut);
            deinterleave_store(tmpOut, nqTot, c.out[0] + (size_t)e * nqTot * W);
        }
        PAR_BLOCKS_END
    }
    static void iprod(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        vec_t sums_j[NQ];
        VecVec tmpIn(nqTot), tmpOut(nmTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            IProductQuadKernel<NM, NM, NQ, NQ, false, false, DEF>(
                tmpIn, c.bdata[0], c.bdata[1], c.w[0], c.w[1], jac_ptr, sums_j, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
    static void physderiv(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        VecVec tmpIn(nqTot), o0(nqTot), o1(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr = &c.df[e * dfSize];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            PhysDerivQuadKernel<NQ, NQ, DEF>(tmpIn, c.Z[0], c.Z[1], c.D[0], c.D[1], df_ptr, o0, o1);
            deinterleave_store(o0, nqTot, c.out[0] + (size_t)e * nqTot * W);
            deinterleave_store(o1, nqTot, c.out[1] + (size_t)e * nqTot * W);
        }
        PAR_BLOCKS_END
    }
    // Helmholtz.h:138-275
    static void helm(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        constexpr int wspSize = NQ > NQ * NM ? NQ : NQ * NM;
        vec_t wsp[wspSize];
        VecVec tmpIn(nmTot), tmpOut(nmTot), bwd(nqTot), deriv0(nqTot), deriv1(nqTot);
        vec_t df0, df1, df2, df3, metric00, metric01, metric11;
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr = &c.df[e * dfSize];
            const vec_t *jac_ptr;
            load_interleave(c.in[0] + (size_t)e * nmTot * W, nmTot, tmpIn);
            if (!DEF)
            {
                df0 = df_ptr[0]; df1 = df_ptr[1]; df2 = df_ptr[2]; df3 = df_ptr[3];
                metric00 = df0 * df0; metric00.fma(df2, df2);
                metric01 = df0 * df1; metric01.fma(df2, df3);
                metric11 = df1 * df1; metric11.fma(df3, df3);
                jac_ptr  = &c.jac[e];
            }
            else
            {
                jac_ptr = &c.jac[(size_t)e * nqTot];
            }
            BwdTransQuadKernel<NM, NM, NQ, NQ>(tmpIn, c.bdata[0], c.bdata[1], wsp, bwd);
            IProductQuadKernel<NM, NM, NQ, NQ, true, false, DEF>(
                bwd, c.bdata[0], c.bdata[1], c.w[0], c.w[1], jac_ptr, wsp, tmpOut, c.lambda);
            PhysDerivTensor2DKernel<NQ, NQ>(bwd, c.D[0], c.D[1], deriv0, deriv1);
            for (int cnt = 0; cnt < nqTot; ++cnt)
            {
                if (DEF)
                {
                    df0 = df_ptr[cnt * ndf]; df1 = df_ptr[cnt * ndf + 1];
                    df2 = df_ptr[cnt * ndf + 2]; df3 = df_ptr[cnt * ndf + 3];
                    metric00 = df0 * df0; metric00.fma(df2, df2);
                    metric01 = df0 * df1; metric01.fma(df2, df3);
                    metric11 = df1 * df1; metric11.fma(df3, df3);
                }
                vec_t d0 = deriv0[cnt], d1 = deriv1[cnt];
                vec_t tmp = metric00 * d0; tmp.fma(metric01, d1); bwd[cnt] = tmp;
                tmp = metric01 * d0; tmp.fma(metric11, d1); deriv0[cnt] = tmp;
            }
            IProductQuadKernel<NM, NM, NQ, NQ, false, true, DEF>(
                bwd, c.dbdata[0], c.bdata[1], c.w[0], c.w[1], jac_ptr, wsp, tmpOut);
            IProductQuadKernel<NM, NM, NQ, NQ, false, true, DEF>(
                deriv0, c.bdata[0], c.dbdata[1], c.w[0], c.w[1], jac_ptr, wsp, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
    // IProductWRTDerivBase.h (Quad Impl, same pattern as Hex :1232-1345)
    static void ipwdb(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        vec_t sums_j[NQ];
        VecVec i0(nqTot), i1(nqTot), t0(nqTot), t1(nqTot), tmpOut(nmTot);
        vec_t df0, df1, df2, df3;
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr  = &c.df[e * dfSize];
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, i0);
            load_interleave(c.in[1] + (size_t)e * nqTot * W, nqTot, i1);
            if (!DEF) { df0 = df_ptr[0]; df1 = df_ptr[1]; df2 = df_ptr[2]; df3 = df_ptr[3]; }
            for (int i = 0; i < nqTot; ++i)
            {
                if (DEF)
                {
                    df0 = df_ptr[i * ndf]; df1 = df_ptr[i * ndf + 1];
                    df2 = df_ptr[i * ndf + 2]; df3 = df_ptr[i * ndf + 3];
                }
                t0[i] = df0 * i0[i] + df2 * i1[i];
                t1[i] = df1 * i0[i] + df3 * i1[i];
            }
            IProductQuadKernel<NM, NM, NQ, NQ, false, false, DEF>(
                t0, c.dbdata[0], c.bdata[1], c.w[0], c.w[1], jac_ptr, sums_j, tmpOut);
            IProductQuadKernel<NM, NM, NQ, NQ, false, true, DEF>(
                t1, c.bdata[0], c.dbdata[1], c.w[0], c.w[1], jac_ptr, sums_j, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
};

// =============================================================== TRI  (nq1 = nq0 - 1)
template <int NM, int NQ, bool DEF> struct TriOps
{
    static constexpr int NQ1 = NQ - 1;
    static constexpr int nmTot = NM * (NM + 1) / 2, nqTot = NQ * NQ1, ndf = 4;
    static constexpr bool CORRECT = true; // eModified_A

    static void bwd(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        vec_t wsp[NM * NQ];
        VecVec tmpIn(nmTot), tmpOut(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            load_interleave(c.in[0] + (size_t)e * nmTot * W, nmTot, tmpIn);
            BwdTransTriKernel<NM, NM, NQ, NQ1, CORRECT>(tmpIn, c.bdata[0], c.bdata[1], wsp, tmpOut);
            deinterleave_store(tmpOut, nqTot, c.out[0] + (size_t)e * nqTot * W);
        }
        PAR_BLOCKS_END
    }
    static void iprod(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        vec_t wsp[NQ];
        VecVec tmpIn(nqTot), tmpOut(nmTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            IProductTriKernel<NM, NM, NQ, NQ1, CORRECT, false, false, DEF>(
                tmpIn, c.bdata[0], c.bdata[1], c.w[0], c.w[1], jac_ptr, wsp, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
    static void physderiv(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        VecVec tmpIn(nqTot), o0(nqTot), o1(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr = &c.df[e * dfSize];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            PhysDerivTriKernel<NQ, NQ1, DEF>(tmpIn, c.Z[0], c.Z[1], c.D[0], c.D[1], df_ptr, o0, o1);
            deinterleave_store(o0, nqTot, c.out[0] + (size_t)e * nqTot * W);
            deinterleave_store(o1, nqTot, c.out[1] + (size_t)e * nqTot * W);
        }
        PAR_BLOCKS_END
    }
    // Helmholtz.h:506-635
    static void helm(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        constexpr int wspSize = NQ1 > NM ? NQ1 : NM;
        vec_t wsp[wspSize];
        VecVec tmpIn(nmTot), tmpOut(nmTot), bwd(nqTot), deriv0(nqTot), deriv1(nqTot);
        vec_t df0, df1, df2, df3, metric00, metric01, metric11;
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr = &c.df[e * dfSize];
            const vec_t *jac_ptr;
            load_interleave(c.in[0] + (size_t)e * nmTot * W, nmTot, tmpIn);
            if (!DEF)
            {
                df0 = df_ptr[0]; df1 = df_ptr[1]; df2 = df_ptr[2]; df3 = df_ptr[3];
                jac_ptr = &c.jac[e];
            }
            else
            {
                jac_ptr = &c.jac[(size_t)e * nqTot];
            }
            BwdTransTriKernel<NM, NM, NQ, NQ1, CORRECT>(tmpIn, c.bdata[0], c.bdata[1], wsp, bwd);
            IProductTriKernel<NM, NM, NQ, NQ1, CORRECT, true, false, DEF>(
                bwd, c.bdata[0], c.bdata[1], c.w[0], c.w[1], jac_ptr, wsp, tmpOut, c.lambda);
            PhysDerivTensor2DKernel<NQ, NQ1>(bwd, c.D[0], c.D[1], deriv0, deriv1);
            for (size_t j = 0, cnt = 0; j < NQ1; ++j)
            {
                vec_t h1j = c.h1[j];
                for (size_t i = 0; i < NQ; ++i, ++cnt)
                {
                    if (DEF)
                    {
                        df0 = df_ptr[cnt * ndf]; df1 = df_ptr[cnt * ndf + 1];
                        df2 = df_ptr[cnt * ndf + 2]; df3 = df_ptr[cnt * ndf + 3];
                    }
                    vec_t h0i = c.h0[i];
                    metric00  = h1j * (df0 + h0i * df1);
                    metric01  = metric00 * df1;
                    metric00  = metric00 * metric00;
                    vec_t tmp = h1j * (df2 + h0i * df3);
                    metric01.fma(tmp, df3);
                    metric00.fma(tmp, tmp);
                    metric11 = df1 * df1;
                    metric11.fma(df3, df3);
                    vec_t d0 = deriv0[cnt], d1 = deriv1[cnt];
                    tmp = metric00 * d0; tmp.fma(metric01, d1); bwd[cnt] = tmp;
                    tmp = metric01 * d0; tmp.fma(metric11, d1); deriv0[cnt] = tmp;
                }
            }
            IProductTriKernel<NM, NM, NQ, NQ1, CORRECT, false, true, DEF>(
                bwd, c.dbdata[0], c.bdata[1], c.w[0], c.w[1], jac_ptr, wsp, tmpOut);
            IProductTriKernel<NM, NM, NQ, NQ1, CORRECT, false, true, DEF>(
                deriv0, c.bdata[0], c.dbdata[1], c.w[0], c.w[1], jac_ptr, wsp, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
    // IProductWRTDerivBase.h:891-1060 (coordim 2)
    static void ipwdb(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        vec_t sums_j[NQ1 > NM ? NQ1 : NM];
        VecVec i0(nqTot), i1(nqTot), t0v(nqTot), t1v(nqTot), tmpOut(nmTot);
        vec_t df0, df1, df2, df3;
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr  = &c.df[e * dfSize];
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, i0);
            load_interleave(c.in[1] + (size_t)e * nqTot * W, nqTot, i1);
            if (!DEF) { df0 = df_ptr[0]; df1 = df_ptr[1]; df2 = df_ptr[2]; df3 = df_ptr[3]; }
            size_t cnt = 0;
            for (size_t j = 0; j < NQ1; ++j)
            {
                vec_t f0 = 2.0 / (1.0 - c.Z[1][j]);
                for (size_t i = 0; i < NQ; ++i, ++cnt)
                {
                    if (DEF)
                    {
                        df0 = df_ptr[cnt * ndf]; df1 = df_ptr[cnt * ndf + 1];
                        df2 = df_ptr[cnt * ndf + 2]; df3 = df_ptr[cnt * ndf + 3];
                    }
                    vec_t a = i0[cnt], b = i1[cnt];
                    vec_t t0 = df0 * a + df2 * b;
                    vec_t t1 = df1 * a + df3 * b;
                    vec_t hf1 = 0.5 * (1.0 + c.Z[0][i]);
                    t0 *= f0;
                    vec_t c1 = hf1 * t1;
                    t0.fma(c1, f0);
                    t0v[cnt] = t0;
                    t1v[cnt] = t1;
                }
            }
            IProductTriKernel<NM, NM, NQ, NQ1, CORRECT, false, false, DEF>(
                t0v, c.dbdata[0], c.bdata[1], c.w[0], c.w[1], jac_ptr, sums_j, tmpOut);
            IProductTriKernel<NM, NM, NQ, NQ1, CORRECT, false, true, DEF>(
                t1v, c.bdata[0], c.dbdata[1], c.w[0], c.w[1], jac_ptr, sums_j, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
};

// =============================================================== HEX
template <int NM, int NQ, bool DEF> struct HexOps
{
    static constexpr int nmTot = NM * NM * NM, nqTot = NQ * NQ * NQ, ndf = 9;

    static void bwd(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        VecVec s1(nqTot), s2(nqTot);
        VecVec tmpIn(nmTot), tmpOut(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            load_interleave(c.in[0] + (size_t)e * nmTot * W, nmTot, tmpIn);
            BwdTransHexKernel<NM, NM, NM, NQ, NQ, NQ>(tmpIn, c.bdata[0], c.bdata[1], c.bdata[2],
                                                     s1.data(), s2.data(), tmpOut);
            deinterleave_store(tmpOut, nqTot, c.out[0] + (size_t)e * nqTot * W);
        }
        PAR_BLOCKS_END
    }
    static void iprod(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        vec_t sums_kj[NQ * NQ], sums_k[NQ];
        VecVec tmpIn(nqTot), tmpOut(nmTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            IProductHexKernel<NM, NM, NM, NQ, NQ, NQ, false, false, DEF>(
                tmpIn, c.bdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr,
                sums_kj, sums_k, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
    static void physderiv(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        VecVec tmpIn(nqTot), o0(nqTot), o1(nqTot), o2(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr = &c.df[e * dfSize];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            PhysDerivHexKernel<NQ, NQ, NQ, DEF>(tmpIn, c.Z[0], c.Z[1], c.Z[2], c.D[0], c.D[1],
                                                c.D[2], df_ptr, o0, o1, o2);
            deinterleave_store(o0, nqTot, c.out[0] + (size_t)e * nqTot * W);
            deinterleave_store(o1, nqTot, c.out[1] + (size_t)e * nqTot * W);
            deinterleave_store(o2, nqTot, c.out[2] + (size_t)e * nqTot * W);
        }
        PAR_BLOCKS_END
    }
    // Helmholtz.h:764-993
    static void helm(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        VecVec wsp1(nqTot), wsp2(nqTot);
        VecVec tmpIn(nmTot), tmpOut(nmTot), bwd(nqTot), deriv0(nqTot), deriv1(nqTot), deriv2(nqTot);
        vec_t df0, df1, df2, df3, df4, df5, df6, df7, df8;
        vec_t metric00, metric01, metric02, metric11, metric12, metric22;
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr = &c.df[e * dfSize];
            const vec_t *jac_ptr;
            load_interleave(c.in[0] + (size_t)e * nmTot * W, nmTot, tmpIn);
            auto metrics = [&]() {
                metric00 = df0 * df0; metric00.fma(df3, df3); metric00.fma(df6, df6);
                metric01 = df0 * df1; metric01.fma(df3, df4); metric01.fma(df6, df7);
                metric02 = df0 * df2; metric02.fma(df3, df5); metric02.fma(df6, df8);
                metric11 = df1 * df1; metric11.fma(df4, df4); metric11.fma(df7, df7);
                metric12 = df1 * df2; metric12.fma(df4, df5); metric12.fma(df7, df8);
                metric22 = df2 * df2; metric22.fma(df5, df5); metric22.fma(df8, df8);
            };
            if (!DEF)
            {
                df0 = df_ptr[0]; df1 = df_ptr[1]; df2 = df_ptr[2];
                df3 = df_ptr[3]; df4 = df_ptr[4]; df5 = df_ptr[5];
                df6 = df_ptr[6]; df7 = df_ptr[7]; df8 = df_ptr[8];
                metrics();
                jac_ptr = &c.jac[e];
            }
            else
            {
                jac_ptr = &c.jac[(size_t)e * nqTot];
            }
            BwdTransHexKernel<NM, NM, NM, NQ, NQ, NQ>(tmpIn, c.bdata[0], c.bdata[0], c.bdata[0],
                                                     wsp1.data(), wsp2.data(), bwd);
            IProductHexKernel<NM, NM, NM, NQ, NQ, NQ, true, false, DEF>(
                bwd, c.bdata[0], c.bdata[0], c.bdata[0], c.w[0], c.w[0], c.w[0], jac_ptr,
                wsp1.data(), wsp2.data(), tmpOut, c.lambda);
            PhysDerivTensor3DKernel<NQ, NQ, NQ>(bwd, c.D[0], c.D[0], c.D[0], deriv0, deriv1, deriv2);
            for (int cnt = 0; cnt < nqTot; ++cnt)
            {
                if (DEF)
                {
                    df0 = df_ptr[cnt * ndf]; df1 = df_ptr[cnt * ndf + 1]; df2 = df_ptr[cnt * ndf + 2];
                    df3 = df_ptr[cnt * ndf + 3]; df4 = df_ptr[cnt * ndf + 4]; df5 = df_ptr[cnt * ndf + 5];
                    df6 = df_ptr[cnt * ndf + 6]; df7 = df_ptr[cnt * ndf + 7]; df8 = df_ptr[cnt * ndf + 8];
                    metrics();
                }
                vec_t d0 = deriv0[cnt], d1 = deriv1[cnt], d2 = deriv2[cnt];
                vec_t tmp = metric00 * d0; tmp.fma(metric01, d1); tmp.fma(metric02, d2); deriv0[cnt] = tmp;
                tmp = metric01 * d0; tmp.fma(metric11, d1); tmp.fma(metric12, d2); deriv1[cnt] = tmp;
                tmp = metric02 * d0; tmp.fma(metric12, d1); tmp.fma(metric22, d2); deriv2[cnt] = tmp;
            }
            IProductHexKernel<NM, NM, NM, NQ, NQ, NQ, false, true, DEF>(
                deriv0, c.dbdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr,
                wsp1.data(), wsp2.data(), tmpOut);
            IProductHexKernel<NM, NM, NM, NQ, NQ, NQ, false, true, DEF>(
                deriv1, c.bdata[0], c.dbdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr,
                wsp1.data(), wsp2.data(), tmpOut);
            IProductHexKernel<NM, NM, NM, NQ, NQ, NQ, false, true, DEF>(
                deriv2, c.bdata[0], c.bdata[1], c.dbdata[2], c.w[0], c.w[1], c.w[2], jac_ptr,
                wsp1.data(), wsp2.data(), tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
    // IProductWRTDerivBase.h:1232-1345
    static void ipwdb(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        vec_t sums_kj[NQ * NQ], sums_k[NQ];
        VecVec i0(nqTot), i1(nqTot), i2(nqTot), t0(nqTot), t1(nqTot), t2(nqTot), tmpOut(nmTot);
        vec_t df0, df1, df2, df3, df4, df5, df6, df7, df8;
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr  = &c.df[e * dfSize];
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, i0);
            load_interleave(c.in[1] + (size_t)e * nqTot * W, nqTot, i1);
            load_interleave(c.in[2] + (size_t)e * nqTot * W, nqTot, i2);
            if (!DEF)
            {
                df0 = df_ptr[0]; df1 = df_ptr[1]; df2 = df_ptr[2];
                df3 = df_ptr[3]; df4 = df_ptr[4]; df5 = df_ptr[5];
                df6 = df_ptr[6]; df7 = df_ptr[7]; df8 = df_ptr[8];
            }
            for (int i = 0; i < nqTot; ++i)
            {
                if (DEF)
                {
                    df0 = df_ptr[i * ndf]; df1 = df_ptr[i * ndf + 1]; df2 = df_ptr[i * ndf + 2];
                    df3 = df_ptr[i * ndf + 3]; df4 = df_ptr[i * ndf + 4]; df5 = df_ptr[i * ndf + 5];
                    df6 = df_ptr[i * ndf + 6]; df7 = df_ptr[i * ndf + 7]; df8 = df_ptr[i * ndf + 8];
                }
                vec_t a = i0[i], b = i1[i], cc = i2[i];
                t0[i] = df0 * a + df3 * b + df6 * cc;
                t1[i] = df1 * a + df4 * b + df7 * cc;
                t2[i] = df2 * a + df5 * b + df8 * cc;
            }
            IProductHexKernel<NM, NM, NM, NQ, NQ, NQ, false, false, DEF>(
                t0, c.dbdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, sums_kj, sums_k, tmpOut);
            IProductHexKernel<NM, NM, NM, NQ, NQ, NQ, false, true, DEF>(
                t1, c.bdata[0], c.dbdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, sums_kj, sums_k, tmpOut);
            IProductHexKernel<NM, NM, NM, NQ, NQ, NQ, false, true, DEF>(
                t2, c.bdata[0], c.bdata[1], c.dbdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, sums_kj, sums_k, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
};

// =============================================================== PRISM (nq = NQ,NQ,NQ-1)
template <int NM, int NQ, bool DEF> struct PrismOps
{
    static constexpr int NQ2 = NQ - 1;
    static constexpr int nmTot = NM * NM * (NM + 1) / 2, nqTot = NQ * NQ * NQ2, ndf = 9;
    static constexpr bool CORRECT = true;

    static void bwd(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        vec_t fpq[NM * NM], fp[NM];
        VecVec tmpIn(nmTot), tmpOut(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            load_interleave(c.in[0] + (size_t)e * nmTot * W, nmTot, tmpIn);
            BwdTransPrismKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT>(
                tmpIn, c.bdata[0], c.bdata[1], c.bdata[2], fpq, fp, tmpOut);
            deinterleave_store(tmpOut, nqTot, c.out[0] + (size_t)e * nqTot * W);
        }
        PAR_BLOCKS_END
    }
    static void iprod(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        vec_t sums_kj[NQ * NQ2], sums_k[NQ2], corr_q[NM];
        VecVec tmpIn(nqTot), tmpOut(nmTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            IProductPrismKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, false, DEF>(
                tmpIn, c.bdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr,
                sums_kj, sums_k, corr_q, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
    static void physderiv(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        VecVec tmpIn(nqTot), o0(nqTot), o1(nqTot), o2(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr = &c.df[e * dfSize];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            PhysDerivPrismKernel<NQ, NQ, NQ2, DEF>(tmpIn, c.Z[0], c.Z[1], c.Z[2], c.D[0], c.D[1],
                                                   c.D[2], df_ptr, o0, o1, o2);
            deinterleave_store(o0, nqTot, c.out[0] + (size_t)e * nqTot * W);
            deinterleave_store(o1, nqTot, c.out[1] + (size_t)e * nqTot * W);
            deinterleave_store(o2, nqTot, c.out[2] + (size_t)e * nqTot * W);
        }
        PAR_BLOCKS_END
    }
    // Helmholtz.h:1291-1458
    static void helm(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        vec_t wsp1[NQ * NQ2 > NM * NM ? NQ * NQ2 : NM * NM], wsp2[NQ2 > NM ? NQ2 : NM], wsp3[NM];
        VecVec tmpIn(nmTot), tmpOut(nmTot), bwd(nqTot), deriv0(nqTot), deriv1(nqTot), deriv2(nqTot);
        vec_t df0, df1, df2, df3, df4, df5, df6, df7, df8;
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr = &c.df[e * dfSize];
            const vec_t *jac_ptr;
            load_interleave(c.in[0] + (size_t)e * nmTot * W, nmTot, tmpIn);
            if (!DEF)
            {
                df0 = df_ptr[0]; df1 = df_ptr[1]; df2 = df_ptr[2];
                df3 = df_ptr[3]; df4 = df_ptr[4]; df5 = df_ptr[5];
                df6 = df_ptr[6]; df7 = df_ptr[7]; df8 = df_ptr[8];
                jac_ptr = &c.jac[e];
            }
            else
            {
                jac_ptr = &c.jac[(size_t)e * nqTot];
            }
            BwdTransPrismKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT>(
                tmpIn, c.bdata[0], c.bdata[1], c.bdata[2], wsp1, wsp2, bwd);
            IProductPrismKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, true, false, DEF>(
                bwd, c.bdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr,
                wsp1, wsp2, wsp3, tmpOut, c.lambda);
            PhysDerivTensor3DKernel<NQ, NQ, NQ2>(bwd, c.D[0], c.D[1], c.D[2], deriv0, deriv1, deriv2);
            for (size_t k = 0, cnt = 0; k < NQ2; ++k)
            {
                vec_t h1 = c.h1[k];
                for (size_t j = 0; j < NQ; ++j)
                    for (size_t i = 0; i < NQ; ++i, cnt++)
                    {
                        vec_t h0 = c.h0[i];
                        if (DEF)
                        {
                            df0 = df_ptr[cnt * ndf]; df1 = df_ptr[cnt * ndf + 1]; df2 = df_ptr[cnt * ndf + 2];
                            df3 = df_ptr[cnt * ndf + 3]; df4 = df_ptr[cnt * ndf + 4]; df5 = df_ptr[cnt * ndf + 5];
                            df6 = df_ptr[cnt * ndf + 6]; df7 = df_ptr[cnt * ndf + 7]; df8 = df_ptr[cnt * ndf + 8];
                        }
                        vec_t tmp1 = h1 * (h0 * df2 + df0);
                        vec_t tmp2 = h1 * (h0 * df5 + df3);
                        vec_t tmp3 = h1 * (h0 * df8 + df6);
                        vec_t g0 = tmp1 * tmp1; g0.fma(tmp2, tmp2); g0.fma(tmp3, tmp3);
                        vec_t g3 = df1 * tmp1; g3.fma(df4, tmp2); g3.fma(df7, tmp3);
                        vec_t g4 = df2 * tmp1; g4.fma(df5, tmp2); g4.fma(df8, tmp3);
                        vec_t g1 = df1 * df1; g1.fma(df4, df4); g1.fma(df7, df7);
                        vec_t g2 = df2 * df2; g2.fma(df5, df5); g2.fma(df8, df8);
                        vec_t g5 = df1 * df2; g5.fma(df4, df5); g5.fma(df7, df8);
                        vec_t d0 = deriv0[cnt], d1 = deriv1[cnt], d2 = deriv2[cnt];
                        tmp1 = g0 * d0; tmp1.fma(g3, d1); tmp1.fma(g4, d2); deriv0[cnt] = tmp1;
                        tmp2 = g3 * d0; tmp2.fma(g1, d1); tmp2.fma(g5, d2); deriv1[cnt] = tmp2;
                        tmp3 = g4 * d0; tmp3.fma(g5, d1); tmp3.fma(g2, d2); deriv2[cnt] = tmp3;
                    }
            }
            IProductPrismKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, true, DEF>(
                deriv0, c.dbdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr,
                wsp1, wsp2, wsp3, tmpOut);
            IProductPrismKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, true, DEF>(
                deriv1, c.bdata[0], c.dbdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr,
                wsp1, wsp2, wsp3, tmpOut);
            IProductPrismKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, true, DEF>(
                deriv2, c.bdata[0], c.bdata[1], c.dbdata[2], c.w[0], c.w[1], c.w[2], jac_ptr,
                wsp1, wsp2, wsp3, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
    // IProductWRTDerivBase.h:1630-1778
    static void ipwdb(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        vec_t wsp1[NQ * NQ2 > NM * NM ? NQ * NQ2 : NM * NM], wsp2[NQ2 > NM ? NQ2 : NM], wsp3[NM];
        VecVec i0(nqTot), i1(nqTot), i2(nqTot), t0v(nqTot), t1v(nqTot), t2v(nqTot), tmpOut(nmTot);
        vec_t df0, df1, df2, df3, df4, df5, df6, df7, df8;
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr  = &c.df[e * dfSize];
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, i0);
            load_interleave(c.in[1] + (size_t)e * nqTot * W, nqTot, i1);
            load_interleave(c.in[2] + (size_t)e * nqTot * W, nqTot, i2);
            if (!DEF)
            {
                df0 = df_ptr[0]; df1 = df_ptr[1]; df2 = df_ptr[2];
                df3 = df_ptr[3]; df4 = df_ptr[4]; df5 = df_ptr[5];
                df6 = df_ptr[6]; df7 = df_ptr[7]; df8 = df_ptr[8];
            }
            for (size_t k = 0, cnt = 0; k < NQ2; ++k)
            {
                vec_t f0 = 2.0 / (1.0 - c.Z[2][k]);
                for (size_t j = 0; j < NQ; ++j)
                    for (size_t i = 0; i < NQ; ++i, ++cnt)
                    {
                        if (DEF)
                        {
                            df0 = df_ptr[cnt * ndf]; df1 = df_ptr[cnt * ndf + 1]; df2 = df_ptr[cnt * ndf + 2];
                            df3 = df_ptr[cnt * ndf + 3]; df4 = df_ptr[cnt * ndf + 4]; df5 = df_ptr[cnt * ndf + 5];
                            df6 = df_ptr[cnt * ndf + 6]; df7 = df_ptr[cnt * ndf + 7]; df8 = df_ptr[cnt * ndf + 8];
                        }
                        vec_t a = i0[cnt], b = i1[cnt], g = i2[cnt];
                        vec_t t0 = df0 * a + df3 * b + df6 * g;
                        vec_t t1 = df1 * a + df4 * b + df7 * g;
                        vec_t t2 = df2 * a + df5 * b + df8 * g;
                        vec_t hf1 = 0.5 * (1.0 + c.Z[0][i]);
                        t0 *= f0;
                        vec_t f1t2 = hf1 * t2;
                        t0.fma(f1t2, f0);
                        t0v[cnt] = t0; t1v[cnt] = t1; t2v[cnt] = t2;
                    }
            }
            IProductPrismKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, false, DEF>(
                t0v, c.dbdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, wsp2, wsp3, tmpOut);
            IProductPrismKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, true, DEF>(
                t1v, c.bdata[0], c.dbdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, wsp2, wsp3, tmpOut);
            IProductPrismKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, true, DEF>(
                t2v, c.bdata[0], c.bdata[1], c.dbdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, wsp2, wsp3, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
};

// =============================================================== TET (nq = NQ,NQ-1,NQ-1)
template <int NM, int NQ, bool DEF> struct TetOps
{
    static constexpr int NQ1 = NQ - 1, NQ2 = NQ - 1;
    static constexpr int nmTot = NM * (NM + 1) * (NM + 2) / 6, nqTot = NQ * NQ1 * NQ2, ndf = 9;
    static constexpr bool CORRECT = true;

    static void bwd(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        vec_t fpq[NM * NM], fp[NM];
        VecVec tmpIn(nmTot), tmpOut(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            load_interleave(c.in[0] + (size_t)e * nmTot * W, nmTot, tmpIn);
            BwdTransTetKernel<NM, NM, NM, NQ, NQ1, NQ2, CORRECT>(
                tmpIn, c.bdata[0], c.bdata[1], c.bdata[2], fpq, fp, tmpOut);
            deinterleave_store(tmpOut, nqTot, c.out[0] + (size_t)e * nqTot * W);
        }
        PAR_BLOCKS_END
    }
    static void iprod(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        vec_t wsp[NQ1 * NQ2 + NQ2];
        VecVec tmpIn(nqTot), tmpOut(nmTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            IProductTetKernel<NM, NM, NM, NQ, NQ1, NQ2, CORRECT, false, false, DEF>(
                tmpIn, c.bdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
    static void physderiv(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        VecVec tmpIn(nqTot), o0(nqTot), o1(nqTot), o2(nqTot), d0(nqTot), d1(nqTot), d2(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr = &c.df[e * dfSize];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            PhysDerivTetKernel<NQ, NQ1, NQ2, DEF>(tmpIn, c.Z[0], c.Z[1], c.Z[2], c.D[0], c.D[1],
                                                  c.D[2], df_ptr, d0, d1, d2, o0, o1, o2);
            deinterleave_store(o0, nqTot, c.out[0] + (size_t)e * nqTot * W);
            deinterleave_store(o1, nqTot, c.out[1] + (size_t)e * nqTot * W);
            deinterleave_store(o2, nqTot, c.out[2] + (size_t)e * nqTot * W);
        }
        PAR_BLOCKS_END
    }
    // Helmholtz.h:2266-2448
    static void helm(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        vec_t wsp1[(NQ1 * NQ2 + NQ2) > NM * NM ? (NQ1 * NQ2 + NQ2) : NM * NM], wsp2[NM];
        VecVec tmpIn(nmTot), tmpOut(nmTot), bwd(nqTot), deriv0(nqTot), deriv1(nqTot), deriv2(nqTot);
        vec_t df0, df1, df2, df3, df4, df5, df6, df7, df8;
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr = &c.df[e * dfSize];
            const vec_t *jac_ptr;
            load_interleave(c.in[0] + (size_t)e * nmTot * W, nmTot, tmpIn);
            if (!DEF)
            {
                df0 = df_ptr[0]; df1 = df_ptr[1]; df2 = df_ptr[2];
                df3 = df_ptr[3]; df4 = df_ptr[4]; df5 = df_ptr[5];
                df6 = df_ptr[6]; df7 = df_ptr[7]; df8 = df_ptr[8];
                jac_ptr = &c.jac[e];
            }
            else
            {
                jac_ptr = &c.jac[(size_t)e * nqTot];
            }
            BwdTransTetKernel<NM, NM, NM, NQ, NQ1, NQ2, CORRECT>(
                tmpIn, c.bdata[0], c.bdata[1], c.bdata[2], wsp1, wsp2, bwd);
            IProductTetKernel<NM, NM, NM, NQ, NQ1, NQ2, CORRECT, true, false, DEF>(
                bwd, c.bdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, tmpOut, c.lambda);
            PhysDerivTensor3DKernel<NQ, NQ1, NQ2>(bwd, c.D[0], c.D[1], c.D[2], deriv0, deriv1, deriv2);
            for (size_t k = 0, cnt = 0; k < NQ2; ++k)
            {
                vec_t h3 = c.h3[k];
                for (size_t j = 0; j < NQ1; ++j)
                {
                    vec_t h1 = c.h1[j], h2 = c.h2[j];
                    vec_t h2h3 = h2 * h3, h1h3 = h1 * h3;
                    for (int i = 0; i < NQ; ++i, ++cnt)
                    {
                        if (DEF)
                        {
                            df0 = df_ptr[cnt * ndf]; df1 = df_ptr[cnt * ndf + 1]; df2 = df_ptr[cnt * ndf + 2];
                            df3 = df_ptr[cnt * ndf + 3]; df4 = df_ptr[cnt * ndf + 4]; df5 = df_ptr[cnt * ndf + 5];
                            df6 = df_ptr[cnt * ndf + 6]; df7 = df_ptr[cnt * ndf + 7]; df8 = df_ptr[cnt * ndf + 8];
                        }
                        vec_t h0h2h3 = c.h0[i] * h2h3;
                        vec_t tmp1 = h0h2h3 * (df1 + df2); tmp1.fma(df0, h2h3);
                        vec_t tmp2 = h0h2h3 * (df4 + df5); tmp2.fma(df3, h2h3);
                        vec_t tmp3 = h0h2h3 * (df7 + df8); tmp3.fma(df6, h2h3);
                        vec_t g0 = tmp1 * tmp1; g0.fma(tmp2, tmp2); g0.fma(tmp3, tmp3);
                        vec_t g4 = df2 * tmp1; g4.fma(df5, tmp2); g4.fma(df8, tmp3);
                        vec_t tmp4 = df1 * h3; tmp4.fma(df2, h1h3);
                        vec_t tmp5 = df4 * h3; tmp5.fma(df5, h1h3);
                        vec_t tmp6 = df7 * h3; tmp6.fma(df8, h1h3);
                        vec_t g3 = tmp1 * tmp4; g3.fma(tmp2, tmp5); g3.fma(tmp3, tmp6);
                        vec_t g1 = tmp4 * tmp4; g1.fma(tmp5, tmp5); g1.fma(tmp6, tmp6);
                        vec_t g5 = df2 * tmp4; g5.fma(df5, tmp5); g5.fma(df8, tmp6);
                        vec_t g2 = df2 * df2; g2.fma(df5, df5); g2.fma(df8, df8);
                        vec_t d0 = deriv0[cnt], d1 = deriv1[cnt], d2 = deriv2[cnt];
                        tmp1 = g0 * d0; tmp1.fma(g3, d1); tmp1.fma(g4, d2); deriv0[cnt] = tmp1;
                        tmp2 = g3 * d0; tmp2.fma(g1, d1); tmp2.fma(g5, d2); deriv1[cnt] = tmp2;
                        tmp3 = g4 * d0; tmp3.fma(g5, d1); tmp3.fma(g2, d2); deriv2[cnt] = tmp3;
                    }
                }
            }
            IProductTetKernel<NM, NM, NM, NQ, NQ1, NQ2, CORRECT, false, true, DEF>(
                deriv0, c.dbdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, tmpOut);
            IProductTetKernel<NM, NM, NM, NQ, NQ1, NQ2, CORRECT, false, true, DEF>(
                deriv1, c.bdata[0], c.dbdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, tmpOut);
            IProductTetKernel<NM, NM, NM, NQ, NQ1, NQ2, CORRECT, false, true, DEF>(
                deriv2, c.bdata[0], c.bdata[1], c.dbdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
    // IProductWRTDerivBase.h:2484-2640
    static void ipwdb(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        vec_t wsp[NQ1 * NQ2 + NQ2];
        VecVec i0(nqTot), i1(nqTot), i2(nqTot), t0v(nqTot), t1v(nqTot), t2v(nqTot), tmpOut(nmTot);
        vec_t df0, df1, df2, df3, df4, df5, df6, df7, df8;
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr  = &c.df[e * dfSize];
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, i0);
            load_interleave(c.in[1] + (size_t)e * nqTot * W, nqTot, i1);
            load_interleave(c.in[2] + (size_t)e * nqTot * W, nqTot, i2);
            if (!DEF)
            {
                df0 = df_ptr[0]; df1 = df_ptr[1]; df2 = df_ptr[2];
                df3 = df_ptr[3]; df4 = df_ptr[4]; df5 = df_ptr[5];
                df6 = df_ptr[6]; df7 = df_ptr[7]; df8 = df_ptr[8];
            }
            for (size_t k = 0, cnt = 0; k < NQ2; ++k)
            {
                vec_t f2 = 2.0 / (1.0 - c.Z[2][k]);
                for (size_t j = 0; j < NQ1; ++j)
                {
                    vec_t Z1Load = c.Z[1][j];
                    vec_t f3 = 0.5 * (1.0 + Z1Load);
                    vec_t f0 = 2.0 * f2 / (1.0 - Z1Load);
                    for (size_t i = 0; i < NQ; ++i, ++cnt)
                    {
                        if (DEF)
                        {
                            df0 = df_ptr[cnt * ndf]; df1 = df_ptr[cnt * ndf + 1]; df2 = df_ptr[cnt * ndf + 2];
                            df3 = df_ptr[cnt * ndf + 3]; df4 = df_ptr[cnt * ndf + 4]; df5 = df_ptr[cnt * ndf + 5];
                            df6 = df_ptr[cnt * ndf + 6]; df7 = df_ptr[cnt * ndf + 7]; df8 = df_ptr[cnt * ndf + 8];
                        }
                        vec_t a = i0[cnt], b = i1[cnt], g = i2[cnt];
                        vec_t t0 = df0 * a + df3 * b + df6 * g;
                        vec_t t1 = df1 * a + df4 * b + df7 * g;
                        vec_t t2 = df2 * a + df5 * b + df8 * g;
                        vec_t f1 = 0.5 * (1.0 + c.Z[0][i]);
                        t0.fma(t1 + t2, f1);
                        t0 *= f0;
                        t1.fma(t2, f3);
                        t1 *= f2;
                        t0v[cnt] = t0; t1v[cnt] = t1; t2v[cnt] = t2;
                    }
                }
            }
            IProductTetKernel<NM, NM, NM, NQ, NQ1, NQ2, CORRECT, false, false, DEF>(
                t0v, c.dbdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp, tmpOut);
            IProductTetKernel<NM, NM, NM, NQ, NQ1, NQ2, CORRECT, false, true, DEF>(
                t1v, c.bdata[0], c.dbdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp, tmpOut);
            IProductTetKernel<NM, NM, NM, NQ, NQ1, NQ2, CORRECT, false, true, DEF>(
                t2v, c.bdata[0], c.bdata[1], c.dbdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
};

// =============================================================== PYR  (nq2 = nq0 - 1)
template <int NM, int NQ, bool DEF> struct PyrOps
{
    static constexpr int NQ2 = NQ - 1;
    static constexpr int nmTot = NM * (NM + 1) * (2 * NM + 1) / 6, nqTot = NQ * NQ * NQ2, ndf = 9;
    static constexpr bool CORRECT = true;

    static void bwd(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        vec_t fpq[NM * NM], fp[NM];
        VecVec tmpIn(nmTot), tmpOut(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            load_interleave(c.in[0] + (size_t)e * nmTot * W, nmTot, tmpIn);
            BwdTransPyrKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT>(tmpIn, c.bdata[0], c.bdata[1], c.bdata[2], fpq, fp, tmpOut);
            deinterleave_store(tmpOut, nqTot, c.out[0] + (size_t)e * nqTot * W);
        }
        PAR_BLOCKS_END
    }
    static void iprod(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        vec_t wsp1[NQ * NQ2], wsp2[NQ2];
        VecVec tmpIn(nqTot), tmpOut(nmTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            IProductPyrKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, false, DEF>(
                tmpIn, c.bdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, wsp2, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
    static void physderiv(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        VecVec tmpIn(nqTot), o0(nqTot), o1(nqTot), o2(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr = &c.df[e * dfSize];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            PhysDerivPyrKernel<NQ, NQ, NQ2, DEF>(tmpIn, c.Z[0], c.Z[1], c.Z[2], c.D[0], c.D[1], c.D[2], df_ptr, o0, o1, o2);
            deinterleave_store(o0, nqTot, c.out[0] + (size_t)e * nqTot * W);
            deinterleave_store(o1, nqTot, c.out[1] + (size_t)e * nqTot * W);
            deinterleave_store(o2, nqTot, c.out[2] + (size_t)e * nqTot * W);
        }
        PAR_BLOCKS_END
    }
    // Helmholtz.h:1771-1955
    static void helm(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        vec_t wsp1[NQ * NQ2 > NM * NM ? NQ * NQ2 : NM * NM], wsp2[NQ2 > NM ? NQ2 : NM];
        VecVec tmpIn(nmTot), tmpOut(nmTot), bwd(nqTot), deriv0(nqTot), deriv1(nqTot), deriv2(nqTot);
        vec_t df0, df1, df2, df3, df4, df5, df6, df7, df8;
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr = &c.df[e * dfSize];
            const vec_t *jac_ptr;
            load_interleave(c.in[0] + (size_t)e * nmTot * W, nmTot, tmpIn);
            if (!DEF)
            {
                df0 = df_ptr[0]; df1 = df_ptr[1]; df2 = df_ptr[2];
                df3 = df_ptr[3]; df4 = df_ptr[4]; df5 = df_ptr[5];
                df6 = df_ptr[6]; df7 = df_ptr[7]; df8 = df_ptr[8];
                jac_ptr = &c.jac[e];
            }
            else
            {
                jac_ptr = &c.jac[(size_t)e * nqTot];
            }
            BwdTransPyrKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT>(tmpIn, c.bdata[0], c.bdata[1], c.bdata[2], wsp1, wsp2, bwd);
            IProductPyrKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, true, false, DEF>(
                bwd, c.bdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, wsp2, tmpOut, c.lambda);
            PhysDerivTensor3DKernel<NQ, NQ, NQ2>(bwd, c.D[0], c.D[1], c.D[2], deriv0, deriv1, deriv2);
            for (size_t k = 0, cnt = 0; k < NQ2; ++k)
            {
                vec_t h2 = c.h2[k];
                for (size_t j = 0; j < NQ; ++j)
                {
                    vec_t h1 = c.h1[j];
                    vec_t h1h2 = h1 * h2;
                    for (size_t i = 0; i < NQ; ++i, cnt++)
                    {
                        vec_t h0 = c.h0[i];
                        vec_t h0h2 = h0 * h2;
                        if (DEF)
                        {
                            df0 = df_ptr[cnt * ndf]; df1 = df_ptr[cnt * ndf + 1]; df2 = df_ptr[cnt * ndf + 2];
                            df3 = df_ptr[cnt * ndf + 3]; df4 = df_ptr[cnt * ndf + 4]; df5 = df_ptr[cnt * ndf + 5];
                            df6 = df_ptr[cnt * ndf + 6]; df7 = df_ptr[cnt * ndf + 7]; df8 = df_ptr[cnt * ndf + 8];
                        }
                        vec_t tmp0 = h2 * df0; tmp0.fma(h0h2, df2);
                        vec_t tmp1 = h2 * df3; tmp1.fma(h0h2, df5);
                        vec_t tmp2 = h2 * df6; tmp2.fma(h0h2, df8);
                        vec_t tmp3 = h2 * df1; tmp3.fma(h1h2, df2);
                        vec_t tmp4 = h2 * df4; tmp4.fma(h1h2, df5);
                        vec_t tmp5 = h2 * df7; tmp5.fma(h1h2, df8);
                        vec_t g0 = tmp0 * tmp0; g0.fma(tmp1, tmp1); g0.fma(tmp2, tmp2);
                        vec_t g1 = tmp3 * tmp3; g1.fma(tmp4, tmp4); g1.fma(tmp5, tmp5);
                        vec_t g2 = df2 * df2; g2.fma(df5, df5); g2.fma(df8, df8);
                        vec_t g3 = tmp0 * tmp3; g3.fma(tmp1, tmp4); g3.fma(tmp2, tmp5);
                        vec_t g4 = df2 * tmp0; g4.fma(df5, tmp1); g4.fma(df8, tmp2);
                        vec_t g5 = df2 * tmp3; g5.fma(df5, tmp4); g5.fma(df8, tmp5);
                        vec_t d0 = deriv0[cnt], d1 = deriv1[cnt], d2 = deriv2[cnt];
                        tmp1 = g0 * d0; tmp1.fma(g3, d1); tmp1.fma(g4, d2); deriv0[cnt] = tmp1;
                        tmp2 = g3 * d0; tmp2.fma(g1, d1); tmp2.fma(g5, d2); deriv1[cnt] = tmp2;
                        tmp3 = g4 * d0; tmp3.fma(g5, d1); tmp3.fma(g2, d2); deriv2[cnt] = tmp3;
                    }
                }
            }
            IProductPyrKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, true, DEF>(
                deriv0, c.dbdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, wsp2, tmpOut);
            IProductPyrKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, true, DEF>(
                deriv1, c.bdata[0], c.dbdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, wsp2, tmpOut);
            IProductPyrKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, true, DEF>(
                deriv2, c.bdata[0], c.bdata[1], c.dbdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, wsp2, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
    // IProductWRTDerivBase.h:2056-2200
    static void ipwdb(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        vec_t wsp1[NQ * NQ2], wsp2[NQ2];
        VecVec i0(nqTot), i1(nqTot), i2(nqTot), t0v(nqTot), t1v(nqTot), t2v(nqTot), tmpOut(nmTot);
        vec_t df0, df1, df2, df3, df4, df5, df6, df7, df8;
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr  = &c.df[e * dfSize];
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, i0);
            load_interleave(c.in[1] + (size_t)e * nqTot * W, nqTot, i1);
            load_interleave(c.in[2] + (size_t)e * nqTot * W, nqTot, i2);
            if (!DEF)
            {
                df0 = df_ptr[0]; df1 = df_ptr[1]; df2 = df_ptr[2];
                df3 = df_ptr[3]; df4 = df_ptr[4]; df5 = df_ptr[5];
                df6 = df_ptr[6]; df7 = df_ptr[7]; df8 = df_ptr[8];
            }
            for (size_t k = 0, cnt = 0; k < NQ2; ++k)
            {
                vec_t f0 = 2.0 / (1.0 - c.Z[2][k]);
                for (size_t j = 0; j < NQ; ++j)
                {
                    vec_t hf2 = 0.5 * (1.0 + c.Z[1][j]);
                    for (size_t i = 0; i < NQ; ++i, ++cnt)
                    {
                        if (DEF)
                        {
                            df0 = df_ptr[cnt * ndf]; df1 = df_ptr[cnt * ndf + 1]; df2 = df_ptr[cnt * ndf + 2];
                            df3 = df_ptr[cnt * ndf + 3]; df4 = df_ptr[cnt * ndf + 4]; df5 = df_ptr[cnt * ndf + 5];
                            df6 = df_ptr[cnt * ndf + 6]; df7 = df_ptr[cnt * ndf + 7]; df8 = df_ptr[cnt * ndf + 8];
                        }
                        vec_t a = i0[cnt], b = i1[cnt], g = i2[cnt];
                        vec_t t0 = df0 * a + df3 * b + df6 * g;
                        vec_t t1 = df1 * a + df4 * b + df7 * g;
                        vec_t t2 = df2 * a + df5 * b + df8 * g;
                        t0 *= f0;
                        vec_t hf1 = 0.5 * (1.0 + c.Z[0][i]);
                        vec_t f1t2 = hf1 * t2;
                        t0.fma(f1t2, f0);
                        t1 *= f0;
                        f1t2 = hf2 * t2;
                        t1.fma(f1t2, f0);
                        t0v[cnt] = t0; t1v[cnt] = t1; t2v[cnt] = t2;
                    }
                }
            }
            IProductPyrKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, false, DEF>(
                t0v, c.dbdata[0], c.bdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, wsp2, tmpOut);
            IProductPyrKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, true, DEF>(
                t1v, c.bdata[0], c.dbdata[1], c.bdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, wsp2, tmpOut);
            IProductPyrKernel<NM, NM, NM, NQ, NQ, NQ2, CORRECT, false, true, DEF>(
                t2v, c.bdata[0], c.bdata[1], c.dbdata[2], c.w[0], c.w[1], c.w[2], jac_ptr, wsp1, wsp2, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
};

// =============================================================== SEG  (coordim 1..3)
template <int NM, int NQ, bool DEF> struct SegOps
{
    static constexpr int nmTot = NM, nqTot = NQ;
    static void bwd(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        VecVec tmpIn(nmTot), tmpOut(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            load_interleave(c.in[0] + (size_t)e * nmTot * W, nmTot, tmpIn);
            BwdTransSegKernel<NM, NQ>(tmpIn, c.bdata[0], tmpOut);
            deinterleave_store(tmpOut, nqTot, c.out[0] + (size_t)e * nqTot * W);
        }
        PAR_BLOCKS_END
    }
    static void iprod(Ctx &c)
    {
        constexpr int W = vec_t::width;
        PAR_BLOCKS_BEGIN(c)
        VecVec tmpIn(nqTot), tmpOut(nmTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            IProductSegKernel<NM, NQ, false, false, DEF>(tmpIn, c.bdata[0], c.w[0], jac_ptr, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
    // PhysDeriv.h:60-250: tensor derivative, then one product per space dimension
    static void physderiv(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const int ndf       = c.coordim;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        VecVec tmpIn(nqTot), d0(nqTot), o(nqTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr = &c.df[e * dfSize];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, tmpIn);
            PhysDerivTensor1DKernel<NQ>(tmpIn, c.D[0], d0);
            for (int dir = 0; dir < ndf; ++dir)
            {
                for (int j = 0; j < NQ; ++j) o[j] = d0[j] * (DEF ? df_ptr[j * ndf + dir] : df_ptr[dir]);
                deinterleave_store(o, nqTot, c.out[dir] + (size_t)e * nqTot * W);
            }
        }
        PAR_BLOCKS_END
    }
    static void helm(Ctx &) {}
    // IProductWRTDerivBase.h:182-365 (incl. the regular coordim-3 branch reading df[1] twice, :321-323)
    static void ipwdb(Ctx &c)
    {
        constexpr int W = vec_t::width;
        const int ndf       = c.coordim;
        const size_t dfSize = DEF ? (size_t)ndf * nqTot : ndf;
        PAR_BLOCKS_BEGIN(c)
        VecVec i0(nqTot), i1(nqTot), i2(nqTot), t0(nqTot), tmpOut(nmTot);
        PAR_FOR
        for (int e = 0; e < c.nBlocks; ++e)
        {
            const vec_t *df_ptr  = &c.df[e * dfSize];
            const vec_t *jac_ptr = DEF ? &c.jac[(size_t)nqTot * e] : &c.jac[e];
            load_interleave(c.in[0] + (size_t)e * nqTot * W, nqTot, i0);
            if (ndf >= 2) load_interleave(c.in[1] + (size_t)e * nqTot * W, nqTot, i1);
            if (ndf == 3) load_interleave(c.in[2] + (size_t)e * nqTot * W, nqTot, i2);
            for (int i = 0; i < nqTot; ++i)
            {
                vec_t df0 = DEF ? df_ptr[i * ndf] : df_ptr[0];
                if (ndf == 1)
                    t0[i] = df0 * i0[i];
                else if (ndf == 2)
                {
                    vec_t df1 = DEF ? df_ptr[i * ndf + 1] : df_ptr[1];
                    t0[i]     = df0 * i0[i] + df1 * i1[i];
                }
                else
                {
                    vec_t df1 = DEF ? df_ptr[i * ndf + 1] : df_ptr[1];
                    vec_t df2 = DEF ? df_ptr[i * ndf + 2] : df_ptr[1];
                    t0[i]     = df0 * i0[i] + df1 * i1[i] + df2 * i2[i];
                }
            }
            IProductSegKernel<NM, NQ, false, false, DEF>(t0, c.dbdata[0], c.w[0], jac_ptr, tmpOut);
            deinterleave_store(tmpOut, nmTot, c.out[0] + (size_t)e * nmTot * W);
        }
        PAR_BLOCKS_END
    }
};

template <class Ops> int run_op(int op, Ctx &c)
{
    switch (op)
    {
        case OP_BWD: Ops::bwd(c); return 0;
        case OP_HELM: Ops::helm(c); return 0;
        case OP_IPROD: Ops::iprod(c); return 0;
        case OP_IPWDB: Ops::ipwdb(c); return 0;
        case OP_PHYSDERIV: Ops::physderiv(c); return 0;
    }
    return -2;
}

template <template <int, int, bool> class Ops, int NM, int NQ> int run_def(int op, bool def, Ctx &c)
{
    return def ? run_op<Ops<NM, NQ, true>>(op, c) : run_op<Ops<NM, NQ, false>>(op, c);
}

// (nm, nq0) pairs instantiated; the reference class dispatch covers nm 2..8, nq nm..2nm
// (e.g. Helmholtz.h:669-761); nm 9..11 (hex only) are direct instantiations of the same
// reference kernel templates.  One translation unit per shape (-DREF_TU_SHAPE=n)
// keeps the build parallel.
#define REF_PAIRS_BASE(X) X(2, 3) X(3, 4) X(4, 5) X(5, 6) X(6, 7) X(7, 8) X(8, 9) X(4, 6) X(5, 8)
#define REF_PAIRS_HEX(X) REF_PAIRS_BASE(X) X(9, 10) X(10, 11) X(11, 12) X(5, 10)

int dispatch_quad(int nm, int nq, int op, bool def, Ctx &c);
int dispatch_tri(int nm, int nq, int op, bool def, Ctx &c);
int dispatch_hex(int nm, int nq, int op, bool def, Ctx &c);
int dispatch_prism(int nm, int nq, int op, bool def, Ctx &c);
int dispatch_tet(int nm, int nq, int op, bool def, Ctx &c);
int dispatch_pyr(int nm, int nq, int op, bool def, Ctx &c);
int dispatch_seg(int nm, int nq, int op, bool def, Ctx &c);

#define X(A, B)                                                                \
    if (nm == A && nq == B) return run_def<OPS, A, B>(op, def, c);
#if defined(REF_TU_SHAPE) && REF_TU_SHAPE == 0
#define OPS QuadOps
int dispatch_quad(int nm, int nq, int op, bool def, Ctx &c) { REF_PAIRS_BASE(X) return -3; }
#elif defined(REF_TU_SHAPE) && REF_TU_SHAPE == 1
#define OPS TriOps
int dispatch_tri(int nm, int nq, int op, bool def, Ctx &c) { REF_PAIRS_BASE(X) return -3; }
#elif defined(REF_TU_SHAPE) && REF_TU_SHAPE == 2
#define OPS HexOps
int dispatch_hex(int nm, int nq, int op, bool def, Ctx &c) { REF_PAIRS_HEX(X) return -3; }
#elif defined(REF_TU_SHAPE) && REF_TU_SHAPE == 3
#define OPS PrismOps
int dispatch_prism(int nm, int nq, int op, bool def, Ctx &c) { REF_PAIRS_BASE(X) return -3; }
#elif defined(REF_TU_SHAPE) && REF_TU_SHAPE == 6
#define OPS SegOps
int dispatch_seg(int nm, int nq, int op, bool def, Ctx &c) { REF_PAIRS_BASE(X) return -3; }
#elif defined(REF_TU_SHAPE) && REF_TU_SHAPE == 4
#define OPS PyrOps
int dispatch_pyr(int nm, int nq, int op, bool def, Ctx &c) { REF_PAIRS_BASE(X) return -3; }
#elif defined(REF_TU_SHAPE) && REF_TU_SHAPE == 5
#define OPS TetOps
int dispatch_tet(int nm, int nq, int op, bool def, Ctx &c) { REF_PAIRS_BASE(X) return -3; }
#endif
#undef X

} // namespace nekref

#ifndef REF_TU_SHAPE
using namespace nekref;
extern "C"
{

int nekref_width(void) { return vec_t::width; }
int nekref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---- Polylib pass-throughs (reference LibUtilities/Polylib/Polylib.cpp)
void nekref_zwglj(double *z, double *w, int np, double a, double b) { Polylib::zwglj(z, w, np, a, b); }
void nekref_zwgrjm(double *z, double *w, int np, double a, double b) { Polylib::zwgrjm(z, w, np, a, b); }
void nekref_Dglj(double *D, const double *z, int np, double a, double b) { Polylib::Dglj(D, z, np, a, b); }
void nekref_Dgrjm(double *D, const double *z, int np, double a, double b) { Polylib::Dgrjm(D, z, np, a, b); }
void nekref_jacobfd(int np, const double *z, double *p, double *pd, int n, double a, double b)
{
    Polylib::jacobfd(np, z, p, pd, n, a, b);
}

// Handle-based API so that timing loops exclude set-up (table broadcast, geometry interleave).
//   shape: 0 quad 1 tri 2 hex 3 prism 4 pyr 5 tet ; op: Collections::OperatorType order
//   tables per direction d<dim: bdata[d] (blen[d] doubles), dbdata[d], D[d] (nq^2), Z[d], w[d] (raw
//   quadrature weights) and ptype[d] (0 GLL, 1 Gauss-Radau-M alpha=1, 2 alpha=2)
//   jac: [nElmt] or [nElmt*nqTot]; df: [ndf][nElmt] or [ndf][nElmt*nqTot] (Nektar Array<TwoD> order)
struct RefHandle
{
    Ctx c;
    int op, shape, nm, nq0, deformed, nPad, nin, nout, nins, nouts;
    std::vector<double> pin[3], pout[3];
};

void *nekref_create2(int op, int shape, int nm, int nq0, int deformed, const double *const *bdata,
                     const double *const *dbdata, const double *const *D, const double *const *Z,
                     const double *const *w, const int *blen, const int *nqd, const int *ptype,
                     int nElmt, const double *jac, const double *df, int coordim);
void *nekref_create(int op, int shape, int nm, int nq0, int deformed, const double *const *bdata,
                    const double *const *dbdata, const double *const *D, const double *const *Z,
                    const double *const *w, const int *blen, const int *nqd, const int *ptype,
                    int nElmt, const double *jac, const double *df)
{
    const int dim = shape == SH_SEG ? 1 : ((shape == SH_QUAD || shape == SH_TRI) ? 2 : 3);
    return nekref_create2(op, shape, nm, nq0, deformed, bdata, dbdata, D, Z, w, blen, nqd, ptype, nElmt, jac, df, dim);
}
// coordim: number of space dimensions (only segments may differ from the element dimension)
void *nekref_create2(int op, int shape, int nm, int nq0, int deformed, const double *const *bdata,
                     const double *const *dbdata, const double *const *D, const double *const *Z,
                     const double *const *w, const int *blen, const int *nqd, const int *ptype,
                     int nElmt, const double *jac, const double *df, int coordim)
{
    RefHandle *h = new RefHandle;
    Ctx &c       = h->c;
    constexpr int W = vec_t::width;
    h->op = op; h->shape = shape; h->nm = nm; h->nq0 = nq0; h->deformed = deformed;
    c.dim     = shape == SH_SEG ? 1 : ((shape == SH_QUAD || shape == SH_TRI) ? 2 : 3);
    c.coordim = shape == SH_SEG ? coordim : c.dim;
    c.nqTot   = 1;
    for (int d = 0; d < c.dim; ++d) c.nqTot *= nqd[d];
    switch (shape)
    {
        case SH_QUAD: c.nmTot = nm * nm; break;
        case SH_TRI: c.nmTot = nm * (nm + 1) / 2; break;
        case SH_HEX: c.nmTot = nm * nm * nm; break;
        case SH_PRISM: c.nmTot = nm * nm * (nm + 1) / 2; break;
        case SH_TET: c.nmTot = nm * (nm + 1) * (nm + 2) / 6; break;
        case SH_PYR: c.nmTot = nm * (nm + 1) * (2 * nm + 1) / 6; break;
        case SH_SEG: c.nmTot = nm; break;
        default: delete h; return nullptr;
    }
    for (int d = 0; d < c.dim; ++d)
    {
        // Operator.hpp:244-258
        double fac = ptype[d] == 1 ? 0.5 : (ptype[d] == 2 ? 0.25 : 1.0);
        bcast(c.bdata[d], bdata[d], blen[d]);
        bcast(c.dbdata[d], dbdata[d], blen[d]);
        bcast(c.D[d], D[d], nqd[d] * nqd[d]);
        bcast(c.Z[d], Z[d], nqd[d]);
        bcast(c.w[d], w[d], nqd[d], fac);
    }
    // collapsed-coordinate factor tables (Helmholtz.h:304-315, 1017-1028, 1985-2004)
    if (shape == SH_TRI)
    {
        c.h0.resize(nqd[0]); c.h1.resize(nqd[1]);
        for (int i = 0; i < nqd[0]; ++i) c.h0[i] = vec_t(0.5 * (1 + Z[0][i]));
        for (int j = 0; j < nqd[1]; ++j) c.h1[j] = vec_t(2.0 / (1 - Z[1][j]));
    }
    else if (shape == SH_PRISM)
    {
        c.h0.resize(nqd[0]); c.h1.resize(nqd[2]);
        for (int i = 0; i < nqd[0]; ++i) c.h0[i] = vec_t(0.5 * (1 + Z[0][i]));
        for (int k = 0; k < nqd[2]; ++k) c.h1[k] = vec_t(2.0 / (1 - Z[2][k]));
    }
    else if (shape == SH_PYR)
    {
        // Helmholtz.h:1489-1507
        c.h0.resize(nqd[0]); c.h1.resize(nqd[1]); c.h2.resize(nqd[2]);
        for (int i = 0; i < nqd[0]; ++i) c.h0[i] = vec_t(0.5 * (1 + Z[0][i]));
        for (int j = 0; j < nqd[1]; ++j) c.h1[j] = vec_t(0.5 * (1 + Z[1][j]));
        for (int k = 0; k < nqd[2]; ++k) c.h2[k] = vec_t(2.0 / (1 - Z[2][k]));
    }
    else if (shape == SH_TET)
    {
        c.h0.resize(nqd[0]); c.h1.resize(nqd[1]); c.h2.resize(nqd[1]); c.h3.resize(nqd[2]);
        for (int i = 0; i < nqd[0]; ++i) c.h0[i] = vec_t(0.5 * (1 + Z[0][i]));
        for (int j = 0; j < nqd[1]; ++j)
        {
            c.h1[j] = vec_t(0.5 * (1 + Z[1][j]));
            c.h2[j] = vec_t(2.0 / (1 - Z[1][j]));
        }
        for (int k = 0; k < nqd[2]; ++k) c.h3[k] = vec_t(2.0 / (1 - Z[2][k]));
    }

    // padding to SIMD width (MatrixFreeBase.h:68-93)
    h->nPad       = (nElmt + W - 1) / W * W;
    c.nBlocks     = h->nPad / W;
    c.nElmt       = nElmt;
    const int ndf = c.dim * c.coordim;
    if (jac) interleave_jac(c.jac, jac, nElmt, c.nBlocks, c.nqTot, deformed != 0);
    if (df) interleave_df(c.df, df, ndf, nElmt, c.nBlocks, c.nqTot, deformed != 0);

    h->nins = 1; h->nouts = 1;
    switch (op)
    {
        case OP_BWD: h->nin = c.nmTot; h->nout = c.nqTot; break;
        case OP_HELM: h->nin = c.nmTot; h->nout = c.nmTot; break;
        case OP_IPROD: h->nin = c.nqTot; h->nout = c.nmTot; break;
        case OP_IPWDB: h->nin = c.nqTot; h->nout = c.nmTot; h->nins = c.coordim; break;
        case OP_PHYSDERIV: h->nin = c.nqTot; h->nout = c.nqTot; h->nouts = c.coordim; break;
        default: delete h; return nullptr;
    }
    if (h->nPad != nElmt)
    {
        for (int a = 0; a < h->nins; ++a) h->pin[a].assign((size_t)h->nin * h->nPad, 0.0);
        for (int a = 0; a < h->nouts; ++a) h->pout[a].assign((size_t)h->nout * h->nPad, 0.0);
    }
    return h;
}

int nekref_run(void *handle, const double *in0, const double *in1, const double *in2, double *out0,
               double *out1, double *out2, double lambda, int nthreads)
{
    RefHandle *h = static_cast<RefHandle *>(handle);
    Ctx &c       = h->c;
    c.lambda     = lambda;
    c.nthreads   = nthreads > 0 ? nthreads : 1;
    const double *ins[3] = {in0, in1, in2};
    double *outs[3]      = {out0, out1, out2};
    const bool padded    = h->nPad != c.nElmt;
    for (int a = 0; a < h->nins; ++a)
    {
        if (padded)
        {
            std::memcpy(h->pin[a].data(), ins[a], sizeof(double) * (size_t)h->nin * c.nElmt);
            c.in[a] = h->pin[a].data();
        }
        else
            c.in[a] = ins[a];
    }
    for (int a = 0; a < h->nouts; ++a) c.out[a] = padded ? h->pout[a].data() : outs[a];

    int rc = -1;
    const bool def = h->deformed != 0;
    switch (h->shape)
    {
        case SH_SEG: rc = dispatch_seg(h->nm, h->nq0, h->op, def, c); break;
        case SH_QUAD: rc = dispatch_quad(h->nm, h->nq0, h->op, def, c); break;
        case SH_TRI: rc = dispatch_tri(h->nm, h->nq0, h->op, def, c); break;
        case SH_HEX: rc = dispatch_hex(h->nm, h->nq0, h->op, def, c); break;
        case SH_PRISM: rc = dispatch_prism(h->nm, h->nq0, h->op, def, c); break;
        case SH_PYR: rc = dispatch_pyr(h->nm, h->nq0, h->op, def, c); break;
        case SH_TET: rc = dispatch_tet(h->nm, h->nq0, h->op, def, c); break;
    }
    if (rc == 0 && padded)
        for (int a = 0; a < h->nouts; ++a)
            std::memcpy(outs[a], h->pout[a].data(), sizeof(double) * (size_t)h->nout * c.nElmt);
    return rc;
}

void nekref_destroy(void *handle) { delete static_cast<RefHandle *>(handle); }

} // extern "C"
#endif // !REF_TU_SHAPE
