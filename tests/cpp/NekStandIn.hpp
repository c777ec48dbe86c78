// NekStandIn.hpp -- TEST DOUBLE, not product code.
//
// A stand-in for the handful of ITHACA-SEM / Nektar++ headers that integration/B200Operators.cpp (the adapter a
// maintainer drops into library/Collections) and its callers need.  The reference's own headers cannot be compiled
// here (they include Boost and the build's generated config), so the types the Collections interface exchanges are
// re-declared with the same names, argument meaning and error behaviour, and the selection logic around the
// operator factory is RESTATED from the reference so that the adapter can be compiled, linked and run on the GPU
// without Nektar++ (tests/cpp/*.cpp, -DNEKB200_STANDIN):
//
//   Array<OneD, T>, Array<TwoD, T>  LibUtilities/BasicUtils/SharedArray.hpp  (ref-counted pointer + offset)
//   MemoryManager<T>                LibUtilities/Memory/NekMemoryManager.hpp (AllocateSharedPtr only)
//   NEKERROR / ASSERTL0             LibUtilities/BasicUtils/ErrorUtil.hpp
//   Vmath::Vcopy                    LibUtilities/BasicUtils/Vmath.hpp
//   LibUtilities::Basis             LibUtilities/Foundations/Basis.h         (GetBdata/GetDbdata/GetD/GetZ/GetW ...)
//   StdRegions::StdExpansion        StdRegions/StdExpansion.h                (GetBasis, DetShapeType, GetNcoeffs ...)
//                                   + the per-element geometric factors a LocalRegions::Expansion would own
//   Collections::OperatorType / ImplementationType / OperatorKey / Operator / OperatorFactory /
//   GetOperatorFactory / CoalescedGeomData / Collection / CollectionOptimisation / SetFixedImpType
//                                   Collections/Operator.h:45-191, Collection.h:53-110,
//                                   CoalescedGeomData.cpp:53-424, CollectionOptimisation.cpp:52-281  (restated)
//   MultiRegions::ExpList           MultiRegions/ExpList.cpp:5005-5151 (CreateCollections) + the four call sites
//                                   (restated)
//
// No operator is registered here: the eB200 operators register themselves from integration/B200Operators.cpp's
// static m_typeArr[] initialisers, exactly as the reference's *_MatrixFree classes do.  Asking the factory for any
// other key throws NekError as the reference's NekFactory does for an unregistered key (NekFactory.hpp:145-209).
#pragma once
#include "../../include/nekmf_b200.h"
#include <cstring>
#include <type_traits>
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace Nektar
{
typedef double NekDouble;

namespace ErrorUtil
{
struct NekError : public std::runtime_error
{
    explicit NekError(const std::string &m) : std::runtime_error(m) {}
};
} // namespace ErrorUtil
#define NEKB200_ERROR(msg)                                                              \
    do                                                                                  \
    {                                                                                   \
        std::ostringstream _s;                                                          \
        _s << "Fatal   : " << msg;                                                      \
        throw ::Nektar::ErrorUtil::NekError(_s.str());                                  \
    } while (0)

// ErrorUtil.hpp:88-186, 246-262
namespace ErrorUtil
{
enum ErrType { efatal, ewarning };
}
#define NEKERROR(type, msg) NEKB200_ERROR(msg)
#define ASSERTL0(cond, msg)                                                             \
    do                                                                                  \
    {                                                                                   \
        if (!(cond)) NEKB200_ERROR(msg);                                                \
    } while (0)

// NekMemoryManager.hpp: only the call the OPERATOR_CREATE macro makes
template <typename T> struct MemoryManager
{
    template <typename... A> static std::shared_ptr<T> AllocateSharedPtr(A &&...a)
    {
        return std::shared_ptr<T>(new T(std::forward<A>(a)...));
    }
};

// ------------------------------------------------------------------------------------------ Array<OneD>
struct OneD {};
struct TwoD {};
template <typename Dim, typename T> class Array;
template <typename T> class Array<OneD, T>
{
public:
    typedef typename std::remove_const<T>::type V;
    Array() : m_size(0), m_off(0) {}
    explicit Array(size_t n, V init = V()) : m_data(new V[n ? n : 1], std::default_delete<V[]>()), m_size(n), m_off(0)
    {
        for (size_t i = 0; i < n; ++i) m_data.get()[i] = init;
    }
    Array(size_t n, const V *src) : Array(n)
    {
        for (size_t i = 0; i < n; ++i) m_data.get()[i] = src[i];
    }
    // const view of a non-const array
    template <typename U> Array(const Array<OneD, U> &o) : m_data(o.m_data), m_size(o.m_size), m_off(o.m_off) {}
    size_t num_elements() const { return m_size - m_off; }
    T *get() const { return m_data ? m_data.get() + m_off : nullptr; }
    T &operator[](size_t i) const { return get()[i]; }
    Array operator+(size_t off) const
    {
        Array r(*this);
        r.m_off += off;
        return r;
    }
    std::shared_ptr<V> m_data;
    size_t m_size, m_off;
};
static Array<OneD, NekDouble> NullNekDouble1DArray;
// rows x columns over one block; operator[] gives the row pointer (what `&df[r][0]` needs)
template <typename T> class Array<TwoD, T>
{
public:
    typedef typename std::remove_const<T>::type V;
    Array() : m_rows(0), m_cols(0) {}
    Array(size_t rows, size_t cols) : m_data(rows * cols), m_rows(rows), m_cols(cols) {}
    template <typename U> Array(const Array<TwoD, U> &o) : m_data(o.m_data), m_rows(o.m_rows), m_cols(o.m_cols) {}
    size_t GetRows() const { return m_rows; }
    size_t GetColumns() const { return m_cols; }
    T *operator[](size_t r) const { return m_data.get() + r * m_cols; }
    T *get() const { return m_data.get(); }
    size_t num_elements() const { return m_rows * m_cols; }

    Array<OneD, V> m_data;
    size_t m_rows, m_cols;
};
} // namespace Nektar
namespace Vmath
{
template <class T> inline void Vcopy(int n, const T *x, int incx, T *y, int incy) // Vmath.hpp:1098-1108
{
    if (incx == 1 && incy == 1) memcpy(y, x, n * sizeof(T));
    else
        for (int i = 0; i < n; ++i) y[i * incy] = x[i * incx];
}
} // namespace Vmath
namespace Nektar
{

// ------------------------------------------------------------------------------------------ LibUtilities
namespace LibUtilities
{
enum ShapeType { eQuadrilateral = NEKMF_QUAD, eTriangle = NEKMF_TRI, eHexahedron = NEKMF_HEX, ePrism = NEKMF_PRISM,
                 ePyramid = NEKMF_PYR, eTetrahedron = NEKMF_TET, eSegment = NEKMF_SEG };
enum BasisType { eModified_A = NEKMF_MODIFIED_A, eModified_B = NEKMF_MODIFIED_B, eModified_C = NEKMF_MODIFIED_C,
                 eModifiedPyr_C = NEKMF_MODIFIEDPYR_C };
enum PointsType { eGaussLobattoLegendre = NEKMF_GLL, eGaussRadauMAlpha1Beta0 = NEKMF_GRJM_A1B0,
                  eGaussRadauMAlpha2Beta0 = NEKMF_GRJM_A2B0 };
static const char *const ShapeTypeMap[] = {"Quadrilateral", "Triangle", "Hexahedron", "Prism", "Pyramid", "Tetrahedron", "Segment"};

// the piece of NekMatrix<NekDouble> the Helper reads: basis->GetD()->GetPtr() (MatrixFreeOps/Operator.hpp:262)
struct DMat
{
    Array<OneD, NekDouble> m_v;
    const Array<OneD, NekDouble> &GetPtr() const { return m_v; }
    const NekDouble *GetRawPtr() const { return m_v.get(); }
};
class Basis
{
public:
    Basis(BasisType bt, int nm, PointsType pt, int nq) : m_bt(bt), m_pt(pt), m_nm(nm), m_nq(nq), m_z(nq), m_w(nq), m_D(nq * nq)
    {
        if (nekmf_points(pt, nq, m_z.get(), m_w.get(), m_D.get())) NEKB200_ERROR("bad points key");
        const int rows = nekmf_basis_rows(bt, nm);
        m_b  = Array<OneD, NekDouble>(rows * nq);
        m_db = Array<OneD, NekDouble>(rows * nq);
        if (nekmf_basis(bt, nm, nq, m_z.get(), m_D.get(), m_b.get(), m_db.get())) NEKB200_ERROR("bad basis key");
    }
    const Array<OneD, NekDouble> &GetBdata() const { return m_b; }
    const Array<OneD, NekDouble> &GetDbdata() const { return m_db; }
    std::shared_ptr<DMat> GetD() const
    {
        auto d = std::make_shared<DMat>();
        d->m_v = m_D;
        return d;
    }
    const Array<OneD, NekDouble> &GetZ() const { return m_z; }
    const Array<OneD, NekDouble> &GetW() const { return m_w; }
    int GetNumModes() const { return m_nm; }
    int GetNumPoints() const { return m_nq; }
    BasisType GetBasisType() const { return m_bt; }
    PointsType GetPointsType() const { return m_pt; }

private:
    BasisType m_bt;
    PointsType m_pt;
    int m_nm, m_nq;
    Array<OneD, NekDouble> m_z, m_w, m_D, m_b, m_db;
};
typedef std::shared_ptr<Basis> BasisSharedPtr;
} // namespace LibUtilities

// ------------------------------------------------------------------------------------------ StdRegions
namespace StdRegions
{
enum ConstFactorType { eFactorLambda, eFactorTau };
typedef std::map<ConstFactorType, NekDouble> ConstFactorMap;
static const ConstFactorMap NullConstFactorMap;

// the expansion of ONE element: reference-element bases + that element's geometric factors
// (jac: 1 or nq values; df: ndf x (1 or nq), df[c*dim+d] = d xi_d / d x_c -- GeomFactors.cpp:399-474)
class StdExpansion : public std::enable_shared_from_this<StdExpansion>
{
public:
    // a LocalRegions::Expansion hands out its reference-element part and its coordinate dimension
    // (StdExpansion.h:381-384, 682-685); the stand-in is both at once
    std::shared_ptr<StdExpansion> GetStdExp() { return shared_from_this(); }
    int GetCoordim() const { return m_dim; }
    StdExpansion(LibUtilities::ShapeType shape, int nummodes, int numpoints0 = -1) : m_shape(shape), m_deformed(false)
    {
        using namespace LibUtilities;
        const int nq0 = numpoints0 > 0 ? numpoints0 : nummodes + 1; // MeshGraph.cpp:1609-1762 defaults
        m_dim         = (shape == eQuadrilateral || shape == eTriangle) ? 2 : 3;
        BasisType bt[3]  = {eModified_A, eModified_A, eModified_A};
        PointsType pt[3] = {eGaussLobattoLegendre, eGaussLobattoLegendre, eGaussLobattoLegendre};
        int nq[3]        = {nq0, nq0, nq0};
        if (shape == eTriangle) { bt[1] = eModified_B; pt[1] = eGaussRadauMAlpha1Beta0; nq[1] = nq0 - 1; }
        else if (shape == ePrism) { bt[2] = eModified_B; pt[2] = eGaussRadauMAlpha1Beta0; nq[2] = nq0 - 1; }
        else if (shape == ePyramid) { bt[2] = eModifiedPyr_C; pt[2] = eGaussRadauMAlpha2Beta0; nq[2] = nq0 - 1; }
        else if (shape == eTetrahedron)
        {
            bt[1] = eModified_B; pt[1] = eGaussRadauMAlpha1Beta0; nq[1] = nq0 - 1;
            bt[2] = eModified_C; pt[2] = eGaussRadauMAlpha2Beta0; nq[2] = nq0 - 1;
        }
        else if (shape != eQuadrilateral && shape != eHexahedron)
            NEKB200_ERROR("shape " << ShapeTypeMap[shape] << " not supported");
        m_ntot = 1;
        for (int d = 0; d < m_dim; ++d)
        {
            m_base.push_back(std::make_shared<Basis>(bt[d], nummodes, pt[d], nq[d]));
            m_ntot *= nq[d];
        }
        const int n = nummodes;
        m_ncoeffs   = shape == eQuadrilateral ? n * n : shape == eTriangle ? n * (n + 1) / 2 : shape == eHexahedron ? n * n * n
                      : shape == ePrism ? n * n * (n + 1) / 2 : shape == ePyramid ? n * (n + 1) * (2 * n + 1) / 6
                      : n * (n + 1) * (n + 2) / 6;
    }
    // share the bases of an existing expansion, new geometry (what ExpList does for every element)
    StdExpansion(const StdExpansion &o, bool deformed, const Array<OneD, NekDouble> &jac, const Array<OneD, NekDouble> &df)
        : m_shape(o.m_shape), m_dim(o.m_dim), m_ncoeffs(o.m_ncoeffs), m_ntot(o.m_ntot), m_base(o.m_base),
          m_deformed(deformed), m_jac(jac), m_df(df)
    {
        const size_t n = deformed ? m_ntot : 1;
        if (jac.num_elements() != n || df.num_elements() != n * m_dim * m_dim)
            NEKB200_ERROR("StdExpansion: geometric factor arrays have the wrong size");
    }
    const LibUtilities::BasisSharedPtr &GetBasis(int d) const { return m_base[d]; }
    LibUtilities::ShapeType DetShapeType() const { return m_shape; }
    int GetShapeDimension() const { return m_dim; }
    int GetNcoeffs() const { return m_ncoeffs; }
    int GetTotPoints() const { return m_ntot; }
    bool IsDeformed() const { return m_deformed; }
    const Array<OneD, NekDouble> &GetJac() const { return m_jac; }
    const Array<OneD, NekDouble> &GetDerivFactors() const { return m_df; }

private:
    LibUtilities::ShapeType m_shape;
    int m_dim, m_ncoeffs, m_ntot;
    std::vector<LibUtilities::BasisSharedPtr> m_base;
    bool m_deformed;
    Array<OneD, NekDouble> m_jac, m_df;
};
typedef std::shared_ptr<StdExpansion> StdExpansionSharedPtr;
} // namespace StdRegions

// ------------------------------------------------------------------------------------------ Collections
namespace Collections
{
enum OperatorType { eBwdTrans, eHelmholtz, eIProductWRTBase, eIProductWRTDerivBase, ePhysDeriv, SIZE_OperatorType };
static const char *const OperatorTypeMap[] = {"BwdTrans", "Helmholtz", "IProductWRTBase", "IProductWRTDerivBase", "PhysDeriv"};
enum ImplementationType { eNoImpType, eNoCollection, eIterPerExp, eStdMat, eSumFac, eMatrixFree, eB200, SIZE_ImplementationType };
static const char *const ImplementationTypeMap[] = {"NoImplementationType", "NoCollection", "IterPerExp", "StdMat",
                                                    "SumFac", "MatrixFree", "B200"};
typedef bool ExpansionIsNodal;
typedef std::map<OperatorType, ImplementationType> OperatorImpMap;
inline OperatorImpMap SetFixedImpType(ImplementationType t) // Operator.cpp:128-138
{
    OperatorImpMap m;
    for (int i = 0; i < SIZE_OperatorType; ++i) m[(OperatorType)i] = t;
    return m;
}

// CoalescedGeomData.cpp:53-113 (GetJac), 251-313 (GetDerivFactors), 406-424 (IsDeformed)
class CoalescedGeomData
{
public:
    bool IsDeformed(const std::vector<StdRegions::StdExpansionSharedPtr> &e) const { return e[0]->IsDeformed(); }
    const Array<OneD, NekDouble> &GetJac(const std::vector<StdRegions::StdExpansionSharedPtr> &e)
    {
        if (m_jac.num_elements() == 0 && !e.empty())
        {
            const size_t n = e[0]->GetJac().num_elements();
            m_jac          = Array<OneD, NekDouble>(n * e.size());
            for (size_t i = 0; i < e.size(); ++i)
                for (size_t q = 0; q < n; ++q) m_jac[i * n + q] = e[i]->GetJac()[q];
        }
        return m_jac;
    }
    // Array<TwoD>[ndf][nElmt(*nq)] (CoalescedGeomData.h:75)
    const Array<TwoD, const NekDouble> &GetDerivFactors(const std::vector<StdRegions::StdExpansionSharedPtr> &e)
    {
        if (m_df.num_elements() == 0 && !e.empty())
        {
            const int dim = e[0]->GetShapeDimension(), ndf = dim * dim;
            const size_t n = e[0]->GetJac().num_elements(), cols = n * e.size();
            Array<TwoD, NekDouble> df(ndf, cols);
            for (size_t i = 0; i < e.size(); ++i)
                for (int r = 0; r < ndf; ++r)
                    for (size_t q = 0; q < n; ++q) df[r][i * n + q] = e[i]->GetDerivFactors()[r * n + q];
            m_df = df;
        }
        return m_df;
    }

private:
    Array<OneD, NekDouble> m_jac;
    Array<TwoD, const NekDouble> m_df;
};
typedef std::shared_ptr<CoalescedGeomData> CoalescedGeomDataSharedPtr;

// Operator.h:45-55
#define OPERATOR_CREATE(cname)                                                          \
    static OperatorKey m_type;                                                          \
    static OperatorKey m_typeArr[];                                                     \
    friend struct MemoryManager<cname>;                                                 \
    static OperatorSharedPtr create(std::vector<StdRegions::StdExpansionSharedPtr> pCollExp, \
                                    std::shared_ptr<CoalescedGeomData> GeomData)        \
    {                                                                                   \
        return MemoryManager<cname>::AllocateSharedPtr(pCollExp, GeomData);             \
    }

// Operator.h:113-165
class Operator
{
public:
    Operator(std::vector<StdRegions::StdExpansionSharedPtr> pCollExp, CoalescedGeomDataSharedPtr GeomData)
        : m_isDeformed(GeomData->IsDeformed(pCollExp)), m_stdExp(pCollExp[0]), m_numElmt(pCollExp.size()),
          m_nqe(pCollExp[0]->GetTotPoints()), m_wspSize(0)
    {
    }
    virtual void operator()(const Array<OneD, const NekDouble> &input, Array<OneD, NekDouble> &output0,
                            Array<OneD, NekDouble> &output1, Array<OneD, NekDouble> &output2,
                            Array<OneD, NekDouble> &wsp, const StdRegions::ConstFactorMap &factors) = 0;
    virtual void operator()(int dir, const Array<OneD, const NekDouble> &input, Array<OneD, NekDouble> &output,
                            Array<OneD, NekDouble> &wsp) = 0;
    virtual ~Operator() {}
    unsigned int GetWspSize() { return m_wspSize; }
    unsigned int GetNumElmt() { return m_numElmt; }
    StdRegions::StdExpansionSharedPtr GetExpSharedPtr() { return m_stdExp; }

protected:
    bool m_isDeformed;
    StdRegions::StdExpansionSharedPtr m_stdExp;
    unsigned int m_numElmt, m_nqe, m_wspSize;
};
typedef std::shared_ptr<Operator> OperatorSharedPtr;
typedef std::tuple<LibUtilities::ShapeType, OperatorType, ImplementationType, ExpansionIsNodal> OperatorKey;

// NekFactory<OperatorKey, Operator, vector<StdExpansionSharedPtr>, CoalescedGeomDataSharedPtr>
class OperatorFactory
{
public:
    typedef OperatorSharedPtr (*CreatorFunction)(std::vector<StdRegions::StdExpansionSharedPtr>, CoalescedGeomDataSharedPtr);
    OperatorKey RegisterCreatorFunction(OperatorKey key, CreatorFunction f, std::string desc = "")
    {
        m_map[key] = std::make_pair(f, desc);
        return key;
    }
    OperatorSharedPtr CreateInstance(OperatorKey key, std::vector<StdRegions::StdExpansionSharedPtr> e,
                                     CoalescedGeomDataSharedPtr g)
    {
        auto it = m_map.find(key);
        if (it == m_map.end())
            NEKB200_ERROR("No such module: (" << LibUtilities::ShapeTypeMap[std::get<0>(key)] << ", "
                                              << OperatorTypeMap[std::get<1>(key)] << ", "
                                              << ImplementationTypeMap[std::get<2>(key)] << ")");
        return it->second.first(e, g);
    }
    bool ModuleExists(OperatorKey key) const { return m_map.count(key) != 0; }

private:
    std::map<OperatorKey, std::pair<CreatorFunction, std::string>> m_map;
};
inline OperatorFactory &GetOperatorFactory()
{
    static OperatorFactory f;
    return f;
}

// ---- session document stand-in + the XML subset the <COLLECTIONS> block needs (the reference reads a TinyXML
// document from LibUtilities::SessionReader; neither is available here)
} // namespace Collections
namespace LibUtilities
{
class SessionReader
{
public:
    static std::shared_ptr<SessionReader> CreateInstance(const std::string &xmlText)
    {
        return std::shared_ptr<SessionReader>(new SessionReader(xmlText));
    }
    const std::string &GetDocumentText() const { return m_text; }

private:
    explicit SessionReader(const std::string &t) : m_text(t) {}
    std::string m_text;
};
typedef std::shared_ptr<SessionReader> SessionReaderSharedPtr;
} // namespace LibUtilities
namespace Collections
{
namespace detail
{
struct XmlNode
{
    std::string tag;
    std::map<std::string, std::string> attr;
    std::vector<XmlNode> children;
    const XmlNode *FirstChild(const std::string &t) const
    {
        for (const XmlNode &c : children)
            if (c.tag == t) return &c;
        return nullptr;
    }
    const char *Attribute(const std::string &k) const
    {
        auto it = attr.find(k);
        return it == attr.end() ? nullptr : it->second.c_str();
    }
};
// elements, attributes (single or double quotes), self-closing tags, comments, declarations; text is ignored
inline XmlNode ParseXml(const std::string &s)
{
    XmlNode doc;
    std::vector<XmlNode *> stack{&doc};
    size_t i = 0;
    auto skipws = [&]() { while (i < s.size() && isspace((unsigned char)s[i])) ++i; };
    auto name = [&]() {
        const size_t b = i;
        while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == ':' || s[i] == '-' || s[i] == '.')) ++i;
        return s.substr(b, i - b);
    };
    while (i < s.size())
    {
        if (s[i] != '<') { ++i; continue; }
        if (s.compare(i, 4, "<!--") == 0)
        {
            const size_t e = s.find("-->", i);
            if (e == std::string::npos) NEKB200_ERROR("XML: unterminated comment");
            i = e + 3;
            continue;
        }
        if (s.compare(i, 2, "<?") == 0 || s.compare(i, 2, "<!") == 0)
        {
            const size_t e = s.find('>', i);
            if (e == std::string::npos) NEKB200_ERROR("XML: unterminated declaration");
            i = e + 1;
            continue;
        }
        if (s.compare(i, 2, "</") == 0)
        {
            i += 2;
            const std::string t = name();
            skipws();
            if (i >= s.size() || s[i] != '>' || stack.size() < 2 || stack.back()->tag != t) NEKB200_ERROR("XML: mismatched </" << t << ">");
            ++i;
            stack.pop_back();
            continue;
        }
        ++i;
        XmlNode node;
        node.tag = name();
        if (node.tag.empty()) NEKB200_ERROR("XML: empty tag name");
        bool closed = false;
        for (;;)
        {
            skipws();
            if (i >= s.size()) NEKB200_ERROR("XML: unterminated tag <" << node.tag);
            if (s[i] == '/') { closed = true; ++i; continue; }
            if (s[i] == '>') { ++i; break; }
            const std::string k = name();
            skipws();
            if (k.empty() || i >= s.size() || s[i] != '=') NEKB200_ERROR("XML: bad attribute in <" << node.tag << ">");
            ++i;
            skipws();
            if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) NEKB200_ERROR("XML: unquoted attribute " << k);
            const char q   = s[i++];
            const size_t e = s.find(q, i);
            if (e == std::string::npos) NEKB200_ERROR("XML: unterminated attribute " << k);
            node.attr[k] = s.substr(i, e - i);
            i            = e + 1;
        }
        // pointers into `children` stay valid only until the next push_back at that level: re-take back()
        stack.back()->children.push_back(node);
        if (!closed) stack.push_back(&stack.back()->children.back());
    }
    if (stack.size() != 1) NEKB200_ERROR("XML: unclosed <" << stack.back()->tag << ">");
    return doc;
}
inline std::string Lower(std::string v)
{
    for (char &c : v) c = (char)tolower((unsigned char)c);
    return v;
}
// ParseUtils::GenerateSeqVector (LibUtilities/BasicUtils/ParseUtils.cpp:108-121): "1-3,5" -> 1 2 3 5
inline bool GenerateSeqVector(const std::string &str, std::vector<unsigned int> &out)
{
    size_t i = 0;
    auto skipws = [&]() { while (i < str.size() && isspace((unsigned char)str[i])) ++i; };
    auto number = [&](unsigned int &v) {
        skipws();
        const size_t b = i;
        v              = 0;
        while (i < str.size() && isdigit((unsigned char)str[i])) v = v * 10 + (unsigned int)(str[i++] - '0');
        return i > b;
    };
    for (;;)
    {
        unsigned int a, b;
        if (!number(a)) return false;
        skipws();
        if (i < str.size() && str[i] == '-')
        {
            ++i;
            if (!number(b)) return false;
            for (unsigned int v = a; v <= b; ++v) out.push_back(v);
        }
        else out.push_back(a);
        skipws();
        if (i == str.size()) return true;
        if (str[i] != ',') return false;
        ++i;
    }
}
} // namespace detail

// CollectionOptimisation.cpp:52-316: which ImplementationType each (operator, shape, order) uses, from the
// constructor default and the session's <COLLECTIONS DEFAULT=".." MAXSIZE=".."><OPERATOR TYPE=".."><ELEMENT TYPE="H"
// ORDER="*" IMPTYPE="B200"/> block; defaults, lookup and error messages follow the reference.  The autotuner
// (SetWithTimings) times the reference's other implementations and is not mirrored.
class CollectionOptimisation
{
public:
    typedef std::pair<LibUtilities::ShapeType, int> ElmtOrder;
    // the unit tests' form: a null dummy session (TestHexCollection.cpp:3699-3704)
    CollectionOptimisation(void *pSession, ImplementationType defaultType = eB200)
    {
        if (pSession) NEKB200_ERROR("CollectionOptimisation: pass a LibUtilities::SessionReaderSharedPtr");
        Init(LibUtilities::SessionReaderSharedPtr(), defaultType);
    }
    CollectionOptimisation(LibUtilities::SessionReaderSharedPtr pSession, ImplementationType defaultType = eNoImpType)
    {
        Init(pSession, defaultType);
    }
    OperatorImpMap GetOperatorImpMap(StdRegions::StdExpansionSharedPtr pExp)
    {
        OperatorImpMap ret;
        const ElmtOrder searchKey(pExp->DetShapeType(), pExp->GetBasis(0)->GetNumModes()), defSearch(pExp->DetShapeType(), -1);
        for (auto &it : m_global)
        {
            auto it2 = it.second.find(searchKey);
            if (it2 == it.second.end()) it2 = it.second.find(defSearch);
            ret[it.first] = it2 == it.second.end() ? eNoCollection : it2->second;
        }
        return ret;
    }
    ImplementationType GetDefaultImplementationType() { return m_defaultType; }
    unsigned int GetMaxCollectionSize() { return m_maxCollSize; }
    bool IsUsingAutotuning() { return m_autotune; }
    bool SetByXml() { return m_setByXml; }

private:
    void Init(LibUtilities::SessionReaderSharedPtr pSession, ImplementationType defaultType)
    {
        using namespace LibUtilities;
        std::map<ElmtOrder, ImplementationType> defaults, defaultsPhysDeriv;
        m_setByXml = m_autotune = false;
        m_maxCollSize            = 0;
        m_defaultType            = defaultType == eNoImpType ? eIterPerExp : defaultType;
        std::map<std::string, ShapeType> elTypes;
        elTypes["S"] = static_cast<ShapeType>(NEKMF_SEG);
        elTypes["T"] = eTriangle; elTypes["Q"] = eQuadrilateral; elTypes["A"] = eTetrahedron;
        elTypes["P"] = ePyramid;  elTypes["R"] = ePrism;         elTypes["H"] = eHexahedron;
        for (auto &it2 : elTypes) defaults[ElmtOrder(it2.second, -1)] = defaultsPhysDeriv[ElmtOrder(it2.second, -1)] = m_defaultType;
        if (defaultType == eNoImpType)
            for (auto &it2 : elTypes)
            {
                for (int i = 1; i < 5; ++i) defaults[ElmtOrder(it2.second, i)] = eStdMat;
                defaultsPhysDeriv[ElmtOrder(it2.second, -1)] = eNoCollection;
                for (int i = 1; i < 3; ++i) defaultsPhysDeriv[ElmtOrder(it2.second, i)] = eSumFac;
            }
        std::map<std::string, OperatorType> opTypes;
        for (int i = 0; i < SIZE_OperatorType; ++i)
        {
            opTypes[OperatorTypeMap[i]]  = (OperatorType)i;
            m_global[(OperatorType)i] = (OperatorType)i == ePhysDeriv ? defaultsPhysDeriv : defaults;
        }
        std::map<std::string, ImplementationType> impTypes;
        for (int i = 0; i < SIZE_ImplementationType; ++i) impTypes[ImplementationTypeMap[i]] = (ImplementationType)i;
        if (!pSession) return; // dummy session: no file reader
        const detail::XmlNode doc    = detail::ParseXml(pSession->GetDocumentText());
        const detail::XmlNode *master = doc.FirstChild("NEKTAR");
        if (!master) NEKB200_ERROR("Unable to find NEKTAR tag in file.");
        const detail::XmlNode *xmlCol = master->FirstChild("COLLECTIONS");
        if (!xmlCol) return;
        const char *maxSize = xmlCol->Attribute("MAXSIZE");
        m_maxCollSize       = maxSize ? (unsigned int)atoi(maxSize) : 0;
        const char *defaultImpl = xmlCol->Attribute("DEFAULT");
        m_defaultType           = defaultType;
        if (defaultType == eNoImpType && defaultImpl)
        {
            const std::string collinfo(defaultImpl);
            m_autotune = detail::Lower(collinfo) == "auto";
            if (!m_autotune)
            {
                bool collectionFound = false;
                for (int i = 1; i < SIZE_ImplementationType; ++i)
                    if (detail::Lower(collinfo) == detail::Lower(ImplementationTypeMap[i]))
                    {
                        m_defaultType   = (ImplementationType)i;
                        collectionFound = true;
                        break;
                    }
                if (!collectionFound) NEKB200_ERROR("Unknown default collection scheme: " + collinfo);
                defaults.clear();
                for (auto &it2 : elTypes) defaults[ElmtOrder(it2.second, -1)] = m_defaultType;
                for (int i = 0; i < SIZE_OperatorType; ++i) m_global[(OperatorType)i] = defaults;
            }
        }
        for (const detail::XmlNode &elmt : xmlCol->children)
        {
            m_setByXml = true;
            if (detail::Lower(elmt.tag) != "operator") NEKB200_ERROR("Only OPERATOR tags are supported inside the COLLECTIONS tag.");
            const char *attr = elmt.Attribute("TYPE");
            if (!attr) NEKB200_ERROR("Missing TYPE in OPERATOR tag.");
            const std::string opType(attr);
            if (!opTypes.count(opType)) NEKB200_ERROR("Unknown OPERATOR type " + opType + ".");
            const OperatorType ot = opTypes[opType];
            for (const detail::XmlNode &elmt2 : elmt.children)
            {
                if (detail::Lower(elmt2.tag) != "element") NEKB200_ERROR("Only ELEMENT tags are supported inside the OPERATOR tag.");
                const char *a1 = elmt2.Attribute("TYPE");
                if (!a1) NEKB200_ERROR("Missing TYPE in ELEMENT tag.");
                const std::string elType(a1);
                auto it2 = elTypes.find(elType);
                if (it2 == elTypes.end()) NEKB200_ERROR("Unknown element type " + elType + " in ELEMENT tag");
                const char *a2 = elmt2.Attribute("IMPTYPE");
                if (!a2) NEKB200_ERROR("Missing IMPTYPE in ELEMENT tag.");
                const std::string impType(a2);
                if (!impTypes.count(impType)) NEKB200_ERROR("Unknown IMPTYPE type " + impType + ".");
                const char *a3 = elmt2.Attribute("ORDER");
                if (!a3) NEKB200_ERROR("Missing ORDER in ELEMENT tag.");
                const std::string order(a3);
                if (order == "*") m_global[ot][ElmtOrder(it2->second, -1)] = impTypes[impType];
                else
                {
                    std::vector<unsigned int> orders;
                    if (!detail::GenerateSeqVector(order, orders)) NEKB200_ERROR("Unable to interpret ORDER '" + order + "'.");
                    for (unsigned int o : orders) m_global[ot][ElmtOrder(it2->second, (int)o)] = impTypes[impType];
                }
            }
        }
    }
    std::map<OperatorType, std::map<ElmtOrder, ImplementationType>> m_global;
    ImplementationType m_defaultType;
    unsigned int m_maxCollSize;
    bool m_setByXml, m_autotune;
};

// Collection.h:53-110, Collection.cpp:46-87
class Collection
{
public:
    Collection(std::vector<StdRegions::StdExpansionSharedPtr> pCollExp, OperatorImpMap &impTypes)
        : m_geomData(std::make_shared<CoalescedGeomData>()), m_collExp(pCollExp), m_impTypes(impTypes)
    {
    }
    void Initialise(const OperatorType opType)
    {
        if (m_ops.count(opType)) return;
        if (m_collExp.empty()) return;
        OperatorKey key(m_collExp[0]->DetShapeType(), opType, m_impTypes[opType], false);
        m_ops[opType] = GetOperatorFactory().CreateInstance(key, m_collExp, m_geomData);
    }
    void ApplyOperator(const OperatorType &op, const Array<OneD, const NekDouble> &in, Array<OneD, NekDouble> &out,
                       const StdRegions::ConstFactorMap &factors = StdRegions::NullConstFactorMap)
    {
        Array<OneD, NekDouble> wsp(Op(op)->GetWspSize()), n1, n2;
        (*Op(op))(in, out, n1, n2, wsp, factors);
    }
    void ApplyOperator(const OperatorType &op, const Array<OneD, const NekDouble> &in, Array<OneD, NekDouble> &out0,
                       Array<OneD, NekDouble> &out1)
    {
        Array<OneD, NekDouble> wsp(Op(op)->GetWspSize()), n2;
        (*Op(op))(in, out0, out1, n2, wsp, StdRegions::NullConstFactorMap);
    }
    void ApplyOperator(const OperatorType &op, const Array<OneD, const NekDouble> &in, Array<OneD, NekDouble> &out0,
                       Array<OneD, NekDouble> &out1, Array<OneD, NekDouble> &out2,
                       const StdRegions::ConstFactorMap &factors = StdRegions::NullConstFactorMap)
    {
        Array<OneD, NekDouble> wsp(Op(op)->GetWspSize());
        (*Op(op))(in, out0, out1, out2, wsp, factors);
    }
    void ApplyOperator(const OperatorType &op, int dir, const Array<OneD, const NekDouble> &in, Array<OneD, NekDouble> &out)
    {
        Array<OneD, NekDouble> wsp(Op(op)->GetWspSize());
        (*Op(op))(dir, in, out, wsp);
    }
    bool HasOperator(const OperatorType &op) { return m_ops.count(op) != 0; }
    OperatorSharedPtr GetOpSharedPtr(const OperatorType &op) { return m_ops[op]; }
    CoalescedGeomDataSharedPtr GetGeomSharedPtr() { return m_geomData; }

protected:
    OperatorSharedPtr &Op(OperatorType op)
    {
        Initialise(op);
        return m_ops[op];
    }
    std::map<OperatorType, OperatorSharedPtr> m_ops;
    CoalescedGeomDataSharedPtr m_geomData;
    std::vector<StdRegions::StdExpansionSharedPtr> m_collExp;
    OperatorImpMap m_impTypes;

public:
    size_t GetNumElmt() const { return m_collExp.size(); }
    const StdRegions::StdExpansionSharedPtr &GetExp(size_t i) const { return m_collExp[i]; }
    const OperatorImpMap &GetImpTypes() const { return m_impTypes; }
};

} // namespace Collections

// ------------------------------------------------------------------------------------------ MultiRegions
namespace MultiRegions
{
// ExpList reduced to what drives the Collections: CreateCollections (ExpList.cpp:5005-5151) and the call sites that
// loop over the collections with their coefficient / quadrature offsets -- IProductWRTBase (:1262-1284), PhysDeriv
// (:1465-1504), BwdTrans (:1961-1989), GeneralMatrixOp for a Helmholtz key (:2359-2397).
class ExpList
{
public:
    ExpList(const std::vector<StdRegions::StdExpansionSharedPtr> &exp,
            LibUtilities::SessionReaderSharedPtr session = LibUtilities::SessionReaderSharedPtr())
        : m_exp(exp), m_session(session), m_ncoeffs(0), m_npoints(0)
    {
        for (auto &e : m_exp)
        {
            m_coeff_offset.push_back(m_ncoeffs);
            m_phys_offset.push_back(m_npoints);
            m_ncoeffs += e->GetNcoeffs();
            m_npoints += e->GetTotPoints();
        }
    }
    int GetNcoeffs() const { return m_ncoeffs; }
    int GetTotPoints() const { return m_npoints; }
    std::vector<Collections::Collection> &GetCollections() { return m_collections; }
    const std::vector<int> &GetCollCoeffOffset() const { return m_coll_coeff_offset; }
    const std::vector<int> &GetCollPhysOffset() const { return m_coll_phys_offset; }

    void CreateCollections(Collections::ImplementationType ImpType = Collections::eNoImpType)
    {
        // the reference iterates a map keyed by LibUtilities::ShapeType: Seg, Tri, Quad, Tet, Pyr, Prism, Hex
        struct ShapeLess
        {
            static int Rank(LibUtilities::ShapeType s)
            {
                switch ((int)s)
                {
                    case NEKMF_SEG: return 1;
                    case NEKMF_TRI: return 2;
                    case NEKMF_QUAD: return 3;
                    case NEKMF_TET: return 4;
                    case NEKMF_PYR: return 5;
                    case NEKMF_PRISM: return 6;
                    default: return 7;
                }
            }
            bool operator()(LibUtilities::ShapeType a, LibUtilities::ShapeType b) const { return Rank(a) < Rank(b); }
        };
        std::map<LibUtilities::ShapeType, std::vector<std::pair<StdRegions::StdExpansionSharedPtr, int>>, ShapeLess> collections;
        Collections::CollectionOptimisation colOpt(m_session, ImpType);
        const int collmax = colOpt.GetMaxCollectionSize() > 0 ? (int)colOpt.GetMaxCollectionSize() : 2 * (int)m_exp.size();
        m_collections.clear();
        m_coll_coeff_offset.clear();
        m_coll_phys_offset.clear();
        for (int i = 0; i < (int)m_exp.size(); ++i) collections[m_exp[i]->DetShapeType()].push_back(std::make_pair(m_exp[i], i));
        for (auto &it : collections)
        {
            Collections::OperatorImpMap impTypes = colOpt.GetOperatorImpMap(it.second[0].first);
            std::vector<StdRegions::StdExpansionSharedPtr> collExp;
            int prevCoeffOffset = m_coeff_offset[it.second[0].second];
            int prevPhysOffset  = m_phys_offset[it.second[0].second];
            m_coll_coeff_offset.push_back(prevCoeffOffset);
            m_coll_phys_offset.push_back(prevPhysOffset);
            collExp.push_back(it.second[0].first);
            int prevnCoeff    = it.second[0].first->GetNcoeffs();
            int prevnPhys     = it.second[0].first->GetTotPoints();
            bool prevDeformed = it.second[0].first->IsDeformed();
            int collcnt       = 1;
            for (size_t i = 1; i < it.second.size(); ++i)
            {
                const int nCoeffs     = it.second[i].first->GetNcoeffs();
                const int nPhys       = it.second[i].first->GetTotPoints();
                const bool Deformed   = it.second[i].first->IsDeformed();
                const int coeffOffset = m_coeff_offset[it.second[i].second];
                const int physOffset  = m_phys_offset[it.second[i].second];
                // next element different, not contiguous, or collmax reached: end the collection, start a new one
                if (prevCoeffOffset + nCoeffs != coeffOffset || prevnCoeff != nCoeffs || prevPhysOffset + nPhys != physOffset ||
                    prevDeformed != Deformed || prevnPhys != nPhys || collcnt >= collmax)
                {
                    m_collections.push_back(Collections::Collection(collExp, impTypes));
                    collExp.clear();
                    m_coll_coeff_offset.push_back(coeffOffset);
                    m_coll_phys_offset.push_back(physOffset);
                    collExp.push_back(it.second[i].first);
                    collcnt = 1;
                }
                else
                {
                    collExp.push_back(it.second[i].first);
                    collcnt++;
                }
                prevCoeffOffset = coeffOffset;
                prevPhysOffset  = physOffset;
                prevDeformed    = Deformed;
                prevnCoeff      = nCoeffs;
                prevnPhys       = nPhys;
            }
            m_collections.push_back(Collections::Collection(collExp, impTypes));
        }
    }
    void BwdTrans(const Array<OneD, const NekDouble> &inarray, Array<OneD, NekDouble> &outarray)
    {
        for (size_t i = 0; i < m_collections.size(); ++i)
        {
            Array<OneD, NekDouble> tmp = outarray + m_coll_phys_offset[i];
            m_collections[i].ApplyOperator(Collections::eBwdTrans, inarray + m_coll_coeff_offset[i], tmp);
        }
    }
    void IProductWRTBase(const Array<OneD, const NekDouble> &inarray, Array<OneD, NekDouble> &outarray)
    {
        for (size_t i = 0; i < m_collections.size(); ++i)
        {
            Array<OneD, NekDouble> tmp = outarray + m_coll_coeff_offset[i];
            m_collections[i].ApplyOperator(Collections::eIProductWRTBase, inarray + m_coll_phys_offset[i], tmp);
        }
    }
    void PhysDeriv(const Array<OneD, const NekDouble> &inarray, Array<OneD, NekDouble> &out_d0, Array<OneD, NekDouble> &out_d1,
                   Array<OneD, NekDouble> &out_d2)
    {
        for (size_t i = 0; i < m_collections.size(); ++i)
        {
            const size_t o = m_coll_phys_offset[i];
            Array<OneD, NekDouble> e0 = out_d0 + o, e1 = out_d1 + o, e2;
            if (m_collections[i].GetExp(0)->GetShapeDimension() == 3)
            {
                e2 = out_d2 + o;
                m_collections[i].ApplyOperator(Collections::ePhysDeriv, inarray + o, e0, e1, e2);
            }
            else m_collections[i].ApplyOperator(Collections::ePhysDeriv, inarray + o, e0, e1);
        }
    }
    void GeneralMatrixOp_Helmholtz(const Array<OneD, const NekDouble> &inarray, Array<OneD, NekDouble> &outarray,
                                   const StdRegions::ConstFactorMap &factors)
    {
        for (size_t i = 0; i < m_collections.size(); ++i)
        {
            Array<OneD, NekDouble> tmp = outarray + m_coll_coeff_offset[i];
            m_collections[i].ApplyOperator(Collections::eHelmholtz, inarray + m_coll_coeff_offset[i], tmp, factors);
        }
    }

private:
    std::vector<StdRegions::StdExpansionSharedPtr> m_exp;
    LibUtilities::SessionReaderSharedPtr m_session;
    int m_ncoeffs, m_npoints;
    std::vector<int> m_coeff_offset, m_phys_offset, m_coll_coeff_offset, m_coll_phys_offset;
    std::vector<Collections::Collection> m_collections;
};
} // namespace MultiRegions
} // namespace Nektar
