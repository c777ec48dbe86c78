// TestCollectionB200.cpp -- C++ parity tests through the Collections mirror, written after the
// pattern of library/UnitTests/Collections/TestHexCollection.cpp:3649-3744: build ONE element,
// replicate it nelmts times, `CollectionOptimisation colOpt(dummySession, eB200)`,
// `Collection c(CollExp, impTypes); c.Initialise(op); c.ApplyOperator(...)` and compare with a second
// implementation of the same operator -- here the CPU oracle (oracle/mf_oracle.h; test
// infrastructure) instead of the LocalRegions routines.  Tolerance 1e-12 relative (the reference
// tests use BOOST_CHECK_CLOSE 1e-8 percent = 1e-10).
#include "NekStandIn.hpp"
#include "../../oracle/mf_oracle.h"
#include <cmath>
#include <cstdio>
#include <random>

using namespace Nektar;
using namespace Nektar::Collections;
typedef Array<OneD, NekDouble> DArray;

static int g_fail = 0, g_run = 0;
#define CHECK(cond, what)                                                  \
    do                                                                     \
    {                                                                      \
        ++g_run;                                                           \
        if (!(cond))                                                       \
        {                                                                  \
            ++g_fail;                                                      \
            printf("FAIL %s:%d %s\n", __FILE__, __LINE__, what);           \
        }                                                                  \
    } while (0)

static double relerr(const DArray &a, const std::vector<double> &b)
{
    double num = 0, den = 0, mx = 0, mb = 0;
    for (size_t i = 0; i < b.size(); ++i)
    {
        num += (a[i] - b[i]) * (a[i] - b[i]);
        den += b[i] * b[i];
        mx = std::fmax(mx, std::fabs(a[i] - b[i]));
        mb = std::fmax(mb, std::fabs(b[i]));
    }
    return std::fmax(std::sqrt(num / den), mx / mb);
}

// trilinear hexahedron: geometric factors at the tensor quadrature points (GeomFactors.cpp:399-474:
// df[c*3+d] = d xi_d / d x_c, jac = det(d x / d xi))
static void HexFactors(const double v[8][3], const DArray &z, bool deformed, DArray &jac, DArray &df)
{
    const int nq = z.num_elements(), n = deformed ? nq * nq * nq : 1;
    jac = DArray(n);
    df  = DArray(9 * n);
    for (int k = 0; k < (deformed ? nq : 1); ++k)
        for (int j = 0; j < (deformed ? nq : 1); ++j)
            for (int i = 0; i < (deformed ? nq : 1); ++i)
            {
                const double xi[3] = {deformed ? z[i] : 0.0, deformed ? z[j] : 0.0, deformed ? z[k] : 0.0};
                double F[3][3] = {{0}};
                for (int a = 0; a < 8; ++a)
                {
                    const double s[3] = {(a & 1) ? 1.0 : -1.0, (a & 2) ? 1.0 : -1.0, (a & 4) ? 1.0 : -1.0};
                    for (int d = 0; d < 3; ++d)
                    {
                        double dN = 0.125 * s[d];
                        for (int o = 0; o < 3; ++o)
                            if (o != d) dN *= (1.0 + s[o] * xi[o]);
                        for (int c = 0; c < 3; ++c) F[c][d] += v[a][c] * dN;
                    }
                }
                const double det = F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) -
                                   F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
                                   F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
                const int pt = deformed ? (k * nq + j) * nq + i : 0;
                jac[pt]      = det;
                for (int d = 0; d < 3; ++d)
                    for (int c = 0; c < 3; ++c)
                    {
                        const int r0 = (c + 1) % 3, r1 = (c + 2) % 3, c0 = (d + 1) % 3, c1 = (d + 2) % 3;
                        // inverse(F)[d][c] = cofactor(F)[c][d] / det
                        df[(c * 3 + d) * n + pt] = (F[r0][c0] * F[r1][c1] - F[r0][c1] * F[r1][c0]) / det;
                    }
            }
}

struct Case
{
    std::vector<StdRegions::StdExpansionSharedPtr> CollExp;
    mfo_elem *el;
    int nelmts, ncoeffs, nq;
    bool deformed;
    std::vector<double> jac, df; // oracle layout: jac[nel(*nq)], df[ndf][nel(*nq)]
};

static Case MakeHex(int nm, int nq0, int nelmts, bool deformed)
{
    using namespace LibUtilities;
    StdRegions::StdExpansion proto(eHexahedron, nm, nq0);
    // the reference's test hex: unit cube corners, "deformed" moves the last vertex to (2,3,4)
    // (TestHexCollection.cpp:3663-3664)
    double v[8][3] = {{-1, -1, -1}, {1, -1, -1}, {-1, 1, -1}, {1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {-1, 1, 1}, {1, 1, 1}};
    if (deformed) { v[7][0] = 2; v[7][1] = 3; v[7][2] = 4; }
    else { for (auto &p : v) { p[0] = 0.5 * p[0] + 0.2 * p[1]; p[2] = 1.5 * p[2] + 0.1 * p[0]; } } // parallelepiped
    DArray jac, df;
    HexFactors(v, proto.GetBasis(0)->GetZ(), deformed, jac, df);
    Case c;
    c.nelmts = nelmts; c.deformed = deformed; c.ncoeffs = proto.GetNcoeffs(); c.nq = proto.GetTotPoints();
    c.el = mfo_create(MFO_HEX, nm, nq0);
    for (int e = 0; e < nelmts; ++e) c.CollExp.push_back(std::make_shared<StdRegions::StdExpansion>(proto, deformed, jac, df));
    const int n = deformed ? c.nq : 1;
    c.jac.resize((size_t)n * nelmts);
    c.df.resize((size_t)9 * n * nelmts);
    for (int e = 0; e < nelmts; ++e)
        for (int q = 0; q < n; ++q)
        {
            c.jac[e * n + q] = jac[q];
            for (int r = 0; r < 9; ++r) c.df[(size_t)r * n * nelmts + e * n + q] = df[r * n + q];
        }
    return c;
}

static void RunHex(int nm, int nq0, int nelmts, bool deformed)
{
    Case c = MakeHex(nm, nq0, nelmts, deformed);
    void *dummySession = nullptr;
    CollectionOptimisation colOpt(dummySession, eB200);
    OperatorImpMap impTypes = colOpt.GetOperatorImpMap(c.CollExp[0]);
    Collection col(c.CollExp, impTypes);
    std::mt19937_64 rng(nm * 100 + nq0);
    std::uniform_real_distribution<double> U(-1, 1);
    const size_t nc = (size_t)c.nelmts * c.ncoeffs, np = (size_t)c.nelmts * c.nq;
    DArray coeffs(nc), phys(np), f1(np), f2(np);
    for (size_t i = 0; i < nc; ++i) coeffs[i] = U(rng);
    for (size_t i = 0; i < np; ++i) { phys[i] = U(rng); f1[i] = U(rng); f2[i] = U(rng); }
    std::vector<double> ref(np), ref1(np), ref2(np), refc(nc);
    char what[128];

    col.Initialise(eBwdTrans);
    DArray out(np);
    col.ApplyOperator(eBwdTrans, coeffs, out);
    mfo_bwdtrans(c.el, c.nelmts, coeffs.get(), ref.data());
    snprintf(what, sizeof(what), "BwdTrans hex nm=%d def=%d err=%.2e", nm, deformed, relerr(out, ref));
    CHECK(relerr(out, ref) < 1e-12, what);

    col.Initialise(eIProductWRTBase);
    DArray outc(nc);
    col.ApplyOperator(eIProductWRTBase, phys, outc);
    mfo_iproduct(c.el, c.nelmts, deformed, c.jac.data(), phys.get(), refc.data());
    snprintf(what, sizeof(what), "IProductWRTBase hex nm=%d def=%d err=%.2e", nm, deformed, relerr(outc, refc));
    CHECK(relerr(outc, refc) < 1e-12, what);

    col.Initialise(ePhysDeriv);
    DArray d0(np), d1(np), d2(np), dd(np);
    col.ApplyOperator(ePhysDeriv, phys, d0, d1, d2);
    mfo_physderiv(c.el, c.nelmts, deformed, c.df.data(), phys.get(), ref.data(), ref1.data(), ref2.data());
    CHECK(relerr(d0, ref) < 1e-12 && relerr(d1, ref1) < 1e-12 && relerr(d2, ref2) < 1e-12, "PhysDeriv hex");
    col.ApplyOperator(ePhysDeriv, 1, phys, dd);
    CHECK(relerr(dd, ref1) < 1e-12, "PhysDeriv hex dir=1");

    col.Initialise(eHelmholtz);
    StdRegions::ConstFactorMap factors;
    factors[StdRegions::eFactorLambda] = 1.5;
    col.ApplyOperator(eHelmholtz, coeffs, outc, factors);
    mfo_helmholtz(c.el, c.nelmts, deformed, c.jac.data(), c.df.data(), 1.5, coeffs.get(), refc.data());
    snprintf(what, sizeof(what), "Helmholtz hex nm=%d nq=%d def=%d err=%.2e", nm, nq0, deformed, relerr(outc, refc));
    CHECK(relerr(outc, refc) < 1e-12, what);

    col.Initialise(eIProductWRTDerivBase);
    col.ApplyOperator(eIProductWRTDerivBase, phys, f1, f2, outc);
    mfo_iproductwrtderivbase(c.el, c.nelmts, deformed, c.jac.data(), c.df.data(), phys.get(), f1.get(), f2.get(), refc.data());
    CHECK(relerr(outc, refc) < 1e-12, "IProductWRTDerivBase hex");

    // error behaviour
    bool threw = false;
    try { col.ApplyOperator(eHelmholtz, 0, coeffs, outc); } catch (const ErrorUtil::NekError &) { threw = true; }
    CHECK(threw, "operator()(dir,...) on Helmholtz must throw NekError");
    threw = false;
    try { col.ApplyOperator(eHelmholtz, coeffs, outc); } catch (const ErrorUtil::NekError &) { threw = true; }
    CHECK(threw, "Helmholtz without eFactorLambda must throw NekError");
    mfo_destroy(c.el);
}

static void TestUnregisteredImplementation()
{
    Case c = MakeHex(4, 5, 2, false);
    OperatorImpMap impTypes = SetFixedImpType(eStdMat);
    Collection col(c.CollExp, impTypes);
    bool threw = false;
    try { col.Initialise(eBwdTrans); } catch (const ErrorUtil::NekError &) { threw = true; }
    CHECK(threw, "factory must reject (Hex, BwdTrans, StdMat)");
    CHECK(GetOperatorFactory().ModuleExists(OperatorKey(LibUtilities::eTetrahedron, eHelmholtz, eB200, false)), "Tet Helmholtz registered");
    // every shape x operator pair the reference registers for eMatrixFree (Collections/*.cpp m_typeArr) exists
    CHECK(GetOperatorFactory().ModuleExists(OperatorKey(LibUtilities::eTetrahedron, eIProductWRTDerivBase, eB200, false)),
          "Tet IProductWRTDerivBase registered");
    CHECK(GetOperatorFactory().ModuleExists(OperatorKey(LibUtilities::ePyramid, eHelmholtz, eB200, false)), "Pyr Helmholtz registered");
    CHECK(!GetOperatorFactory().ModuleExists(OperatorKey(LibUtilities::eTetrahedron, eHelmholtz, eStdMat, false)),
          "only eB200 is registered here");
    mfo_destroy(c.el);
}

int main()
{
    if (nekmf_device_count() < 1)
    {
        printf("no CUDA device: nothing to run (the B200 path has no CPU fallback)\n");
        return 77;
    }
    // TestHexCollection.cpp variants: UniformP (nm 4, nq 5) / VariableP-like (nm 4..8) / OverInt (nq up to 2 nm) /
    // MultiElmt (10 elements: not a multiple of any SIMD width) / Undeformed + Deformed
    for (int deformed = 0; deformed < 2; ++deformed)
    {
        RunHex(4, 5, 1, deformed);
        RunHex(4, 5, 10, deformed);
        RunHex(5, 6, 10, deformed);
        RunHex(8, 9, 10, deformed);
        RunHex(4, 8, 10, deformed); // over-integration
    }
    TestUnregisteredImplementation();
    printf("%s: %d checks, %d failures\n", g_fail ? "FAILED" : "PASSED", g_run, g_fail);
    return g_fail ? 1 : 0;
}
