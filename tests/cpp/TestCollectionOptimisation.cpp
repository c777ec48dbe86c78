// TestCollectionOptimisation.cpp -- the selection logic of Collections::CollectionOptimisation
// (/root/reference/library/Collections/CollectionOptimisation.cpp:52-316) in the C++ host mirror
// (ithaca-sem_b200/host/NekB200Collections.hpp): constructor defaults, the <COLLECTIONS> block of a session
// document, the (shape, order) -> shape default -> eNoCollection lookup and the reference's error messages.
// Needs no GPU: nothing here creates an operator.
#include "NekStandIn.hpp"
#include <cstdio>
#include <cstring>

using namespace Nektar;
using namespace Nektar::Collections;

static int g_fail = 0;
#define CHECK(cond, what)                                   \
    do                                                      \
    {                                                       \
        if (!(cond))                                        \
        {                                                   \
            printf("FAILED: %s (%s)\n", what, #cond);       \
            ++g_fail;                                       \
        }                                                   \
    } while (0)

static bool throws(const std::string &xml, const char *msg)
{
    try
    {
        CollectionOptimisation opt(LibUtilities::SessionReader::CreateInstance(xml));
    }
    catch (const ErrorUtil::NekError &e)
    {
        return strstr(e.what(), msg) != nullptr;
    }
    return false;
}

int main()
{
    auto exp = [](LibUtilities::ShapeType s, int nm) { return std::make_shared<StdRegions::StdExpansion>(s, nm); };
    auto hex5 = exp(LibUtilities::eHexahedron, 5), hex7 = exp(LibUtilities::eHexahedron, 7), hex3 = exp(LibUtilities::eHexahedron, 3);
    auto tet2 = exp(LibUtilities::eTetrahedron, 2), tet6 = exp(LibUtilities::eTetrahedron, 6);
    {
        void *dummySession = nullptr; // TestHexCollection.cpp:3699-3704
        CollectionOptimisation opt(dummySession, eB200);
        OperatorImpMap m = opt.GetOperatorImpMap(hex5);
        bool all = true;
        for (auto &it : m) all = all && it.second == eB200;
        CHECK(all && m.size() == SIZE_OperatorType, "dummy session, fixed type");
        CHECK(!opt.SetByXml() && !opt.IsUsingAutotuning() && opt.GetMaxCollectionSize() == 0, "dummy session flags");
    }
    {
        CollectionOptimisation opt(LibUtilities::SessionReaderSharedPtr(), eNoImpType);
        CHECK(opt.GetDefaultImplementationType() == eIterPerExp, "default type");
        OperatorImpMap lo = opt.GetOperatorImpMap(tet2), hi = opt.GetOperatorImpMap(tet6);
        CHECK(lo[eBwdTrans] == eStdMat && lo[ePhysDeriv] == eSumFac, "low-order defaults");
        CHECK(hi[eHelmholtz] == eIterPerExp && hi[ePhysDeriv] == eNoCollection, "high-order defaults");
    }
    const std::string xml = "<?xml version='1.0'?>\n<NEKTAR>\n <!-- selection -->\n"
                            " <COLLECTIONS DEFAULT=\"b200\" MAXSIZE=\"64\">\n"
                            "  <OPERATOR TYPE=\"Helmholtz\">\n"
                            "   <ELEMENT TYPE=\"H\" ORDER=\"2-4,7\" IMPTYPE=\"MatrixFree\" />\n"
                            "   <ELEMENT TYPE='A' ORDER='*' IMPTYPE='StdMat'/>\n"
                            "  </OPERATOR>\n </COLLECTIONS>\n <EXPANSIONS><E COMPOSITE=\"C[0]\"/></EXPANSIONS>\n</NEKTAR>\n";
    {
        CollectionOptimisation opt(LibUtilities::SessionReader::CreateInstance(xml));
        CHECK(opt.GetDefaultImplementationType() == eB200 && opt.GetMaxCollectionSize() == 64 && opt.SetByXml(), "session flags");
        CHECK(opt.GetOperatorImpMap(hex5)[eHelmholtz] == eB200, "order 5 not in 2-4,7");
        CHECK(opt.GetOperatorImpMap(hex7)[eHelmholtz] == eMatrixFree, "order 7 in 2-4,7");
        CHECK(opt.GetOperatorImpMap(hex3)[eHelmholtz] == eMatrixFree && opt.GetOperatorImpMap(hex3)[eBwdTrans] == eB200, "order 3");
        CHECK(opt.GetOperatorImpMap(tet6)[eHelmholtz] == eStdMat && opt.GetOperatorImpMap(tet6)[eBwdTrans] == eB200, "ORDER=*");
        CollectionOptimisation fixed(LibUtilities::SessionReader::CreateInstance(xml), eMatrixFree);
        CHECK(fixed.GetOperatorImpMap(hex5)[eBwdTrans] == eMatrixFree, "constructor type wins over DEFAULT");
        CollectionOptimisation au(LibUtilities::SessionReader::CreateInstance("<NEKTAR><COLLECTIONS DEFAULT=\"auto\"/></NEKTAR>"));
        CHECK(au.IsUsingAutotuning(), "DEFAULT=auto");
    }
    CHECK(throws("<FOO/>", "Unable to find NEKTAR tag"), "no NEKTAR tag");
    CHECK(throws("<NEKTAR><COLLECTIONS DEFAULT=\"Fast\"/></NEKTAR>", "Unknown default collection scheme: Fast"), "unknown default");
    CHECK(throws("<NEKTAR><COLLECTIONS><THING/></COLLECTIONS></NEKTAR>", "Only OPERATOR tags"), "non-OPERATOR child");
    CHECK(throws("<NEKTAR><COLLECTIONS><OPERATOR/></COLLECTIONS></NEKTAR>", "Missing TYPE in OPERATOR tag"), "missing TYPE");
    CHECK(throws("<NEKTAR><COLLECTIONS><OPERATOR TYPE=\"Mass\"/></COLLECTIONS></NEKTAR>", "Unknown OPERATOR type Mass"), "unknown operator");
    CHECK(throws("<NEKTAR><COLLECTIONS><OPERATOR TYPE=\"BwdTrans\"><ELEMENT TYPE=\"X\" ORDER=\"*\" IMPTYPE=\"B200\"/></OPERATOR>"
                 "</COLLECTIONS></NEKTAR>", "Unknown element type X"), "unknown element");
    CHECK(throws("<NEKTAR><COLLECTIONS><OPERATOR TYPE=\"BwdTrans\"><ELEMENT TYPE=\"H\" ORDER=\"*\" IMPTYPE=\"Cuda\"/></OPERATOR>"
                 "</COLLECTIONS></NEKTAR>", "Unknown IMPTYPE type Cuda"), "unknown imptype");
    CHECK(throws("<NEKTAR><COLLECTIONS><OPERATOR TYPE=\"BwdTrans\"><ELEMENT TYPE=\"H\" IMPTYPE=\"B200\"/></OPERATOR>"
                 "</COLLECTIONS></NEKTAR>", "Missing ORDER in ELEMENT tag"), "missing order");
    CHECK(throws("<NEKTAR><COLLECTIONS></NEKTAR>", "XML: mismatched"), "malformed document");
    {
        // a Collection built from the map refuses implementation types that are not registered here (NekFactory.hpp:145-209)
        CollectionOptimisation opt(LibUtilities::SessionReader::CreateInstance(xml));
        std::vector<StdRegions::StdExpansionSharedPtr> v(3, tet6);
        OperatorImpMap imp = opt.GetOperatorImpMap(tet6);
        Collection c(v, imp);
        bool refused = false;
        try { c.Initialise(eHelmholtz); }
        catch (const ErrorUtil::NekError &e) { refused = strstr(e.what(), "No such module") != nullptr; }
        CHECK(refused, "unregistered implementation type");
    }
    {
        // MultiRegions::ExpList::CreateCollections (ExpList.cpp:5005-5151): per-shape passes in ShapeType order, a
        // collection ends at a gap in the arrays, a change of nCoeffs / nPhys / deformed-ness, or at collmax members
        auto hex4 = exp(LibUtilities::eHexahedron, 4);
        auto tet4 = exp(LibUtilities::eTetrahedron, 4);
        auto reg = [](const StdRegions::StdExpansionSharedPtr &b) {
            const int d = b->GetShapeDimension();
            return std::make_shared<StdRegions::StdExpansion>(*b, false, Array<OneD, NekDouble>(1, 1.0), Array<OneD, NekDouble>(d * d, 1.0));
        };
        auto def = [](const StdRegions::StdExpansionSharedPtr &b) {
            const int d = b->GetShapeDimension(), n = b->GetTotPoints();
            return std::make_shared<StdRegions::StdExpansion>(*b, true, Array<OneD, NekDouble>(n, 1.0), Array<OneD, NekDouble>(d * d * n, 1.0));
        };
        std::vector<StdRegions::StdExpansionSharedPtr> mesh{reg(hex4), reg(hex4), reg(hex4), def(hex4), def(hex4), reg(tet4),
                                                          reg(tet4), reg(hex5), reg(hex4), reg(hex4)};
        MultiRegions::ExpList list(mesh);
        list.CreateCollections(eB200);
        auto &c = list.GetCollections();
        const int nc4 = hex4->GetNcoeffs(), nq4 = hex4->GetTotPoints(), nct = tet4->GetNcoeffs(), nqt = tet4->GetTotPoints();
        const size_t sizes[5] = {2, 3, 2, 1, 2};
        bool ok = c.size() == 5;
        for (size_t i = 0; ok && i < 5; ++i) ok = c[i].GetNumElmt() == sizes[i];
        CHECK(ok, "CreateCollections group sizes");
        CHECK(ok && c[0].GetExp(0)->DetShapeType() == LibUtilities::eTetrahedron && c[2].GetExp(0)->IsDeformed() &&
                  c[3].GetExp(0)->GetNcoeffs() == hex5->GetNcoeffs(), "CreateCollections group contents");
        const std::vector<int> co{5 * nc4, 0, 3 * nc4, 5 * nc4 + 2 * nct, 5 * nc4 + 2 * nct + hex5->GetNcoeffs()};
        const std::vector<int> po{5 * nq4, 0, 3 * nq4, 5 * nq4 + 2 * nqt, 5 * nq4 + 2 * nqt + hex5->GetTotPoints()};
        CHECK(list.GetCollCoeffOffset() == co && list.GetCollPhysOffset() == po, "CreateCollections offsets");
        MultiRegions::ExpList capped(mesh, LibUtilities::SessionReader::CreateInstance(
                                               "<NEKTAR><COLLECTIONS DEFAULT=\"B200\" MAXSIZE=\"2\"/></NEKTAR>"));
        capped.CreateCollections();
        const size_t csz[6] = {2, 2, 1, 2, 1, 2};
        bool okc = capped.GetCollections().size() == 6;
        for (size_t i = 0; okc && i < 6; ++i)
            okc = capped.GetCollections()[i].GetNumElmt() == csz[i] && capped.GetCollections()[i].GetImpTypes().at(eHelmholtz) == eB200;
        CHECK(okc, "CreateCollections with MAXSIZE");
    }
    printf(g_fail ? "%d check(s) FAILED\n" : "PASSED%.0d\n", g_fail);
    return g_fail ? 1 : 0;
}
