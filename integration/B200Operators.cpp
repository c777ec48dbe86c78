///////////////////////////////////////////////////////////////////////////////
// File: B200Operators.cpp  --  drop into library/Collections/, link with -lnekmf_b200
//
// The eB200 ImplementationType of the five matrix-free Collections operators: BwdTrans, IProductWRTBase,
// PhysDeriv, Helmholtz and IProductWRTDerivBase on Seg / Quad / Tri / Hex / Prism / Pyr / Tet, each a thin class
// over one nekmf_op_t of the C ABI (include/nekmf_b200.h).  They take the place of the *_MatrixFree classes
//     Collections/BwdTrans.cpp:140-240, IProductWRTBase.cpp:160-272, PhysDeriv.cpp:270-425,
//     Helmholtz.cpp:380-505, IProductWRTDerivBase.cpp:274-432
// and of their MatrixFreeBase.h:46-186 helpers (no SIMD padding: the device kernels take any element count).
// Registration is the reference's own mechanism: static m_typeArr[] initialisers calling
// GetOperatorFactory().RegisterCreatorFunction(OperatorKey(shape, op, eB200, false), ...).
//
// Besides this file a maintainer adds `eB200` before SIZE_ImplementationType and "B200" to
// ImplementationTypeMap (Collections/Operator.h:84-103); see INTEGRATION.md.
//
// This repository compiles the same file against tests/cpp/NekStandIn.hpp (-DNEKB200_STANDIN), a test double of
// the Nektar++ headers used below, and runs it on the GPU (tests/cpp/TestCollectionB200.cpp).
///////////////////////////////////////////////////////////////////////////////
#ifdef NEKB200_STANDIN
#include "NekStandIn.hpp"
#else
#include <Collections/CoalescedGeomData.h>
#include <Collections/Collection.h>
#include <Collections/Operator.h>
#include <LibUtilities/BasicUtils/Vmath.hpp>
#endif
#include <nekmf_b200.h>

using namespace std;

namespace Nektar
{
namespace Collections
{

using LibUtilities::eHexahedron;
using LibUtilities::ePrism;
using LibUtilities::ePyramid;
using LibUtilities::eQuadrilateral;
using LibUtilities::eSegment;
using LibUtilities::eTetrahedron;
using LibUtilities::eTriangle;

namespace
{
int ShapeToAbi(LibUtilities::ShapeType s)
{
    switch (s)
    {
        case eSegment:       return NEKMF_SEG;
        case eQuadrilateral: return NEKMF_QUAD;
        case eTriangle:      return NEKMF_TRI;
        case eHexahedron:    return NEKMF_HEX;
        case ePrism:         return NEKMF_PRISM;
        case ePyramid:       return NEKMF_PYR;
        case eTetrahedron:   return NEKMF_TET;
        default:
            NEKERROR(ErrorUtil::efatal, "B200 operators: shape not supported");
            return -1;
    }
}

int BasisToAbi(LibUtilities::BasisType b)
{
    switch (b)
    {
        case LibUtilities::eModified_A:    return NEKMF_MODIFIED_A;
        case LibUtilities::eModified_B:    return NEKMF_MODIFIED_B;
        case LibUtilities::eModified_C:    return NEKMF_MODIFIED_C;
        case LibUtilities::eModifiedPyr_C: return NEKMF_MODIFIEDPYR_C;
        default:
            NEKERROR(ErrorUtil::efatal, "B200 operators: only the modified C0 bases are supported "
                                        "(the MatrixFree operators have the same restriction)");
            return -1;
    }
}

int PointsToAbi(LibUtilities::PointsType p)
{
    switch (p)
    {
        case LibUtilities::eGaussLobattoLegendre:   return NEKMF_GLL;
        case LibUtilities::eGaussRadauMAlpha1Beta0: return NEKMF_GRJM_A1B0;
        case LibUtilities::eGaussRadauMAlpha2Beta0: return NEKMF_GRJM_A2B0;
        default:
            NEKERROR(ErrorUtil::efatal, "B200 operators: points distribution not supported");
            return -1;
    }
}

void Check(int rc)
{
    if (rc != NEKMF_OK)
    {
        NEKERROR(ErrorUtil::efatal, nekmf_last_error());
    }
}

/**
 * @brief What the five operators share: the device operator object, created from the first expansion of the
 * collection (bases, points, weights, derivative matrices: MatrixFreeOps/Operator.hpp:223-273 reads the same
 * accessors) and the collection's coalesced geometric factors in their plain, non-interleaved layout.
 */
class B200Base
{
protected:
    nekmf_op_t m_op;
    int m_dim;      ///< shape dimension
    int m_coordim;  ///< number of physical directions (== m_dim except for segments)
    unsigned int m_nIn, m_nOut;

    B200Base(int optype, vector<StdRegions::StdExpansionSharedPtr> &pCollExp,
             CoalescedGeomDataSharedPtr &pGeomData, bool needJac, bool needDF)
        : m_op(nullptr)
    {
        StdRegions::StdExpansionSharedPtr stdExp = pCollExp[0]->GetStdExp();
        m_dim     = stdExp->GetShapeDimension();
        m_coordim = pCollExp[0]->GetCoordim();

        int nm[3] = {1, 1, 1}, nq[3] = {1, 1, 1}, bt[3] = {0, 0, 0}, pt[3] = {0, 0, 0};
        const double *b[3] = {nullptr, nullptr, nullptr}, *db[3] = {nullptr, nullptr, nullptr};
        const double *D[3] = {nullptr, nullptr, nullptr}, *Z[3] = {nullptr, nullptr, nullptr};
        const double *W[3] = {nullptr, nullptr, nullptr};
        // keep the derivative matrices alive until nekmf_op_create has copied them
        std::shared_ptr<void> keepD[3];
        for (int d = 0; d < m_dim; ++d)
        {
            LibUtilities::BasisSharedPtr bas = pCollExp[0]->GetBasis(d);
            nm[d] = bas->GetNumModes();
            nq[d] = bas->GetNumPoints();
            bt[d] = BasisToAbi(bas->GetBasisType());
            pt[d] = PointsToAbi(bas->GetPointsType());
            b[d]  = bas->GetBdata().get();
            db[d] = bas->GetDbdata().get();
            auto Dm  = bas->GetD();
            keepD[d] = Dm;
            D[d]     = Dm->GetPtr().get();
            Z[d]     = bas->GetZ().get();
            W[d]     = bas->GetW().get();
        }
        const bool deformed = pGeomData->IsDeformed(pCollExp);
        const int nElmt     = (int)pCollExp.size();
        Check(nekmf_op_create(ShapeToAbi(stdExp->DetShapeType()), optype, nm, nq, bt, pt, b, db, D, Z, W, nElmt,
                              deformed ? 1 : 0, m_coordim, &m_op));

        if (needJac || needDF)
        {
            const double *jac = nullptr;
            Array<OneD, NekDouble> dfFlat;
            if (needJac)
            {
                jac = pGeomData->GetJac(pCollExp).get();
            }
            if (needDF)
            {
                // Array<TwoD>[dim*coordim][nElmt (* nq)] -> the ABI's contiguous rows
                const Array<TwoD, const NekDouble> &df = pGeomData->GetDerivFactors(pCollExp);
                const int ndf = m_dim * m_coordim, n = (int)df.GetColumns();
                dfFlat        = Array<OneD, NekDouble>(ndf * n);
                for (int r = 0; r < ndf; ++r)
                {
                    Vmath::Vcopy(n, &df[r][0], 1, &dfFlat[r * n], 1);
                }
            }
            Check(nekmf_op_set_geom(m_op, jac, needDF ? dfFlat.get() : nullptr, NEKMF_HOST));
        }
        m_nIn  = nElmt * stdExp->GetNcoeffs();
        m_nOut = nElmt * stdExp->GetTotPoints();
    }

    ~B200Base()
    {
        nekmf_op_destroy(m_op);
    }

    /// "operator()(dir, ...) is not valid": what the reference's one-direction overloads raise for the
    /// operators that have none (e.g. Collections/BwdTrans.cpp:201-207)
    static void NoDirectionalForm(const char *name)
    {
        NEKERROR(ErrorUtil::efatal, std::string(name) + ": operator()(dir, ...) is not valid for this operator.");
    }
};
} // namespace

/**
 * @brief Backward transform on the device (replaces BwdTrans_MatrixFree, Collections/BwdTrans.cpp:140-240).
 */
class BwdTrans_B200 : public Operator, B200Base
{
public:
    OPERATOR_CREATE(BwdTrans_B200)

    ~BwdTrans_B200()
    {
    }

    void operator()(const Array<OneD, const NekDouble> &input, Array<OneD, NekDouble> &output0,
                    Array<OneD, NekDouble> &output1, Array<OneD, NekDouble> &output2, Array<OneD, NekDouble> &wsp,
                    const StdRegions::ConstFactorMap &factors) override
    {
        (void)output1; (void)output2; (void)wsp; (void)factors;
        Check(nekmf_op_apply(m_op, input.get(), nullptr, nullptr, output0.get(), nullptr, nullptr, NEKMF_HOST));
    }

    void operator()(int dir, const Array<OneD, const NekDouble> &input, Array<OneD, NekDouble> &output,
                    Array<OneD, NekDouble> &wsp) override
    {
        (void)dir; (void)input; (void)output; (void)wsp;
        NoDirectionalForm("BwdTrans_B200");
    }

private:
    BwdTrans_B200(vector<StdRegions::StdExpansionSharedPtr> pCollExp, CoalescedGeomDataSharedPtr pGeomData)
        : Operator(pCollExp, pGeomData), B200Base(NEKMF_BWDTRANS, pCollExp, pGeomData, false, false)
    {
    }
};

OperatorKey BwdTrans_B200::m_typeArr[] = {
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eSegment, eBwdTrans, eB200, false),
                                                 BwdTrans_B200::create, "BwdTrans_B200_Seg"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eQuadrilateral, eBwdTrans, eB200, false),
                                                 BwdTrans_B200::create, "BwdTrans_B200_Quad"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eTriangle, eBwdTrans, eB200, false),
                                                 BwdTrans_B200::create, "BwdTrans_B200_Tri"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eHexahedron, eBwdTrans, eB200, false),
                                                 BwdTrans_B200::create, "BwdTrans_B200_Hex"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(ePrism, eBwdTrans, eB200, false),
                                                 BwdTrans_B200::create, "BwdTrans_B200_Prism"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eTetrahedron, eBwdTrans, eB200, false),
                                                 BwdTrans_B200::create, "BwdTrans_B200_Tet"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(ePyramid, eBwdTrans, eB200, false),
                                                 BwdTrans_B200::create, "BwdTrans_B200_Pyr")};

/**
 * @brief Inner product with the basis on the device (replaces IProductWRTBase_MatrixFree,
 * Collections/IProductWRTBase.cpp:160-272).
 */
class IProductWRTBase_B200 : public Operator, B200Base
{
public:
    OPERATOR_CREATE(IProductWRTBase_B200)

    ~IProductWRTBase_B200()
    {
    }

    void operator()(const Array<OneD, const NekDouble> &input, Array<OneD, NekDouble> &output0,
                    Array<OneD, NekDouble> &output1, Array<OneD, NekDouble> &output2, Array<OneD, NekDouble> &wsp,
                    const StdRegions::ConstFactorMap &factors) override
    {
        (void)output1; (void)output2; (void)wsp; (void)factors;
        Check(nekmf_op_apply(m_op, input.get(), nullptr, nullptr, output0.get(), nullptr, nullptr, NEKMF_HOST));
    }

    void operator()(int dir, const Array<OneD, const NekDouble> &input, Array<OneD, NekDouble> &output,
                    Array<OneD, NekDouble> &wsp) override
    {
        (void)dir; (void)input; (void)output; (void)wsp;
        NoDirectionalForm("IProductWRTBase_B200");
    }

private:
    IProductWRTBase_B200(vector<StdRegions::StdExpansionSharedPtr> pCollExp, CoalescedGeomDataSharedPtr pGeomData)
        : Operator(pCollExp, pGeomData), B200Base(NEKMF_IPRODUCTWRTBASE, pCollExp, pGeomData, true, false)
    {
    }
};

OperatorKey IProductWRTBase_B200::m_typeArr[] = {
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eSegment, eIProductWRTBase, eB200, false),
                                                 IProductWRTBase_B200::create, "IProductWRTBase_B200_Seg"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eQuadrilateral, eIProductWRTBase, eB200, false),
                                                 IProductWRTBase_B200::create, "IProductWRTBase_B200_Quad"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eTriangle, eIProductWRTBase, eB200, false),
                                                 IProductWRTBase_B200::create, "IProductWRTBase_B200_Tri"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eHexahedron, eIProductWRTBase, eB200, false),
                                                 IProductWRTBase_B200::create, "IProductWRTBase_B200_Hex"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(ePrism, eIProductWRTBase, eB200, false),
                                                 IProductWRTBase_B200::create, "IProductWRTBase_B200_Prism"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(ePyramid, eIProductWRTBase, eB200, false),
                                                 IProductWRTBase_B200::create, "IProductWRTBase_B200_Pyr"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eTetrahedron, eIProductWRTBase, eB200, false),
                                                 IProductWRTBase_B200::create, "IProductWRTBase_B200_Tet")};

/**
 * @brief Physical derivatives on the device (replaces PhysDeriv_MatrixFree, Collections/PhysDeriv.cpp:270-425).
 */
class PhysDeriv_B200 : public Operator, B200Base
{
public:
    OPERATOR_CREATE(PhysDeriv_B200)

    ~PhysDeriv_B200()
    {
    }

    void operator()(const Array<OneD, const NekDouble> &input, Array<OneD, NekDouble> &output0,
                    Array<OneD, NekDouble> &output1, Array<OneD, NekDouble> &output2, Array<OneD, NekDouble> &wsp,
                    const StdRegions::ConstFactorMap &factors) override
    {
        (void)wsp; (void)factors;
        // PhysDeriv.cpp:300-320: as many outputs as coordinate directions
        switch (m_coordim)
        {
            case 1:
                Check(nekmf_op_apply(m_op, input.get(), nullptr, nullptr, output0.get(), nullptr, nullptr, NEKMF_HOST));
                break;
            case 2:
                Check(nekmf_op_apply(m_op, input.get(), nullptr, nullptr, output0.get(), output1.get(), nullptr,
                                     NEKMF_HOST));
                break;
            case 3:
                Check(nekmf_op_apply(m_op, input.get(), nullptr, nullptr, output0.get(), output1.get(), output2.get(),
                                     NEKMF_HOST));
                break;
            default:
                NEKERROR(ErrorUtil::efatal, "Unknown coordinate dimension");
                break;
        }
    }

    void operator()(int dir, const Array<OneD, const NekDouble> &input, Array<OneD, NekDouble> &output,
                    Array<OneD, NekDouble> &wsp) override
    {
        (void)wsp;
        // PhysDeriv.cpp:323-341: every direction is computed, one is handed back
        ASSERTL0(dir >= 0 && dir < m_coordim, "PhysDeriv_B200: direction out of range");
        Array<OneD, NekDouble> tmp[3];
        for (int d = 0; d < m_coordim; ++d)
        {
            tmp[d] = d == dir ? output : Array<OneD, NekDouble>(m_nOut);
        }
        Check(nekmf_op_apply(m_op, input.get(), nullptr, nullptr, tmp[0].get(), m_coordim > 1 ? tmp[1].get() : nullptr,
                             m_coordim > 2 ? tmp[2].get() : nullptr, NEKMF_HOST));
    }

private:
    PhysDeriv_B200(vector<StdRegions::StdExpansionSharedPtr> pCollExp, CoalescedGeomDataSharedPtr pGeomData)
        : Operator(pCollExp, pGeomData), B200Base(NEKMF_PHYSDERIV, pCollExp, pGeomData, false, true)
    {
    }
};

OperatorKey PhysDeriv_B200::m_typeArr[] = {
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eSegment, ePhysDeriv, eB200, false),
                                                 PhysDeriv_B200::create, "PhysDeriv_B200_Seg"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eTriangle, ePhysDeriv, eB200, false),
                                                 PhysDeriv_B200::create, "PhysDeriv_B200_Tri"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eQuadrilateral, ePhysDeriv, eB200, false),
                                                 PhysDeriv_B200::create, "PhysDeriv_B200_Quad"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eHexahedron, ePhysDeriv, eB200, false),
                                                 PhysDeriv_B200::create, "PhysDeriv_B200_Hex"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(ePrism, ePhysDeriv, eB200, false),
                                                 PhysDeriv_B200::create, "PhysDeriv_B200_Prism"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(ePyramid, ePhysDeriv, eB200, false),
                                                 PhysDeriv_B200::create, "PhysDeriv_B200_Pyr"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eTetrahedron, ePhysDeriv, eB200, false),
                                                 PhysDeriv_B200::create, "PhysDeriv_B200_Tet")};

/**
 * @brief Helmholtz operator on the device (replaces Helmholtz_MatrixFree, Collections/Helmholtz.cpp:380-505).
 */
class Helmholtz_B200 : public Operator, B200Base
{
public:
    OPERATOR_CREATE(Helmholtz_B200)

    ~Helmholtz_B200()
    {
    }

    void operator()(const Array<OneD, const NekDouble> &input, Array<OneD, NekDouble> &output0,
                    Array<OneD, NekDouble> &output1, Array<OneD, NekDouble> &output2, Array<OneD, NekDouble> &wsp,
                    const StdRegions::ConstFactorMap &factors) override
    {
        (void)output1; (void)output2; (void)wsp;
        // Helmholtz.cpp:400-428: lambda is taken from the factor map on every call
        auto x = factors.find(StdRegions::eFactorLambda);
        ASSERTL0(x != factors.end(), "Helmholtz_B200: eFactorLambda is missing from the constant factor map");
        Check(nekmf_op_set_lambda(m_op, x->second));
        Check(nekmf_op_apply(m_op, input.get(), nullptr, nullptr, output0.get(), nullptr, nullptr, NEKMF_HOST));
    }

    void operator()(int dir, const Array<OneD, const NekDouble> &input, Array<OneD, NekDouble> &output,
                    Array<OneD, NekDouble> &wsp) override
    {
        (void)dir; (void)input; (void)output; (void)wsp;
        NoDirectionalForm("Helmholtz_B200");
    }

private:
    Helmholtz_B200(vector<StdRegions::StdExpansionSharedPtr> pCollExp, CoalescedGeomDataSharedPtr pGeomData)
        : Operator(pCollExp, pGeomData), B200Base(NEKMF_HELMHOLTZ, pCollExp, pGeomData, true, true)
    {
    }
};

OperatorKey Helmholtz_B200::m_typeArr[] = {
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eQuadrilateral, eHelmholtz, eB200, false),
                                                 Helmholtz_B200::create, "Helmholtz_B200_Quad"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eTriangle, eHelmholtz, eB200, false),
                                                 Helmholtz_B200::create, "Helmholtz_B200_Tri"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eHexahedron, eHelmholtz, eB200, false),
                                                 Helmholtz_B200::create, "Helmholtz_B200_Hex"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(ePrism, eHelmholtz, eB200, false),
                                                 Helmholtz_B200::create, "Helmholtz_B200_Prism"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(ePyramid, eHelmholtz, eB200, false),
                                                 Helmholtz_B200::create, "Helmholtz_B200_Pyr"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eTetrahedron, eHelmholtz, eB200, false),
                                                 Helmholtz_B200::create, "Helmholtz_B200_Tet")};

/**
 * @brief Inner product with the derivatives of the basis on the device (replaces IProductWRTDerivBase_MatrixFree,
 * Collections/IProductWRTDerivBase.cpp:274-432).
 */
class IProductWRTDerivBase_B200 : public Operator, B200Base
{
public:
    OPERATOR_CREATE(IProductWRTDerivBase_B200)

    ~IProductWRTDerivBase_B200()
    {
    }

    /// The reference's calling convention (IProductWRTDerivBase.cpp:283-353): coordim inputs followed by the
    /// output -- (in0, out), (in0, in1, out) or (in0, in1, in2, out).
    void operator()(const Array<OneD, const NekDouble> &entry0, Array<OneD, NekDouble> &entry1,
                    Array<OneD, NekDouble> &entry2, Array<OneD, NekDouble> &entry3, Array<OneD, NekDouble> &wsp,
                    const StdRegions::ConstFactorMap &factors) override
    {
        (void)wsp; (void)factors;
        switch (m_coordim)
        {
            case 1:
                Check(nekmf_op_apply(m_op, entry0.get(), nullptr, nullptr, entry1.get(), nullptr, nullptr, NEKMF_HOST));
                break;
            case 2:
                Check(nekmf_op_apply(m_op, entry0.get(), entry1.get(), nullptr, entry2.get(), nullptr, nullptr,
                                     NEKMF_HOST));
                break;
            case 3:
                Check(nekmf_op_apply(m_op, entry0.get(), entry1.get(), entry2.get(), entry3.get(), nullptr, nullptr,
                                     NEKMF_HOST));
                break;
            default:
                NEKERROR(ErrorUtil::efatal, "coordim not valid");
                break;
        }
    }

    void operator()(int dir, const Array<OneD, const NekDouble> &input, Array<OneD, NekDouble> &output,
                    Array<OneD, NekDouble> &wsp) override
    {
        (void)dir; (void)input; (void)output; (void)wsp;
        NoDirectionalForm("IProductWRTDerivBase_B200");
    }

private:
    IProductWRTDerivBase_B200(vector<StdRegions::StdExpansionSharedPtr> pCollExp, CoalescedGeomDataSharedPtr pGeomData)
        : Operator(pCollExp, pGeomData), B200Base(NEKMF_IPRODUCTWRTDERIVBASE, pCollExp, pGeomData, true, true)
    {
    }
};

OperatorKey IProductWRTDerivBase_B200::m_typeArr[] = {
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eSegment, eIProductWRTDerivBase, eB200, false),
                                                 IProductWRTDerivBase_B200::create, "IProductWRTDerivBase_B200_Seg"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eQuadrilateral, eIProductWRTDerivBase, eB200, false),
                                                 IProductWRTDerivBase_B200::create, "IProductWRTDerivBase_B200_Quad"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eTriangle, eIProductWRTDerivBase, eB200, false),
                                                 IProductWRTDerivBase_B200::create, "IProductWRTDerivBase_B200_Tri"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eHexahedron, eIProductWRTDerivBase, eB200, false),
                                                 IProductWRTDerivBase_B200::create, "IProductWRTDerivBase_B200_Hex"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(ePrism, eIProductWRTDerivBase, eB200, false),
                                                 IProductWRTDerivBase_B200::create, "IProductWRTDerivBase_B200_Prism"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(ePyramid, eIProductWRTDerivBase, eB200, false),
                                                 IProductWRTDerivBase_B200::create, "IProductWRTDerivBase_B200_Pyr"),
    GetOperatorFactory().RegisterCreatorFunction(OperatorKey(eTetrahedron, eIProductWRTDerivBase, eB200, false),
                                                 IProductWRTDerivBase_B200::create, "IProductWRTDerivBase_B200_Tet")};

} // namespace Collections
} // namespace Nektar
