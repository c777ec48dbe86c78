"""TEST DOUBLES, not product code: Python restatements of the reference's selection logic around the operator
factory -- `CollectionOptimisation` (Collections/CollectionOptimisation.cpp:52-281: the session file's <COLLECTIONS>
block, defaults, (shape, order) lookup, error messages) and `ExpList::CreateCollections` with the four call sites
(MultiRegions/ExpList.cpp:5005-5151, 1262-1284, 1465-1504, 1961-1989, 2359-2397).  A real build uses the reference's
own classes unchanged; these exist so that the path from `<COLLECTIONS DEFAULT="B200">` to `nekmf_op_apply` can be
exercised end to end without Nektar++ (tests/test_host_logic.py on the CPU, tests/test_gpu_explist.py on the GPU).
The C++ twin is tests/cpp/NekStandIn.hpp."""
import numpy as np

from _util import nekmf

_nk = nekmf()
globals().update({k: v for k, v in vars(_nk).items() if not k.startswith("__")})  # the restated code uses nekmf's names unqualified


def GenerateSeqVector(text):
    """ParseUtils::GenerateSeqVector (LibUtilities/BasicUtils/ParseUtils.cpp:108-121): "1-3,5" -> [1, 2, 3, 5]."""
    out = []
    for item in text.split(","):
        item = item.strip()
        lo, sep, hi = item.partition("-")
        if sep:
            if not (lo.strip().isdigit() and hi.strip().isdigit()):
                raise NekError("cannot parse sequence '%s'" % text)
            out.extend(range(int(lo), int(hi) + 1))
        else:
            if not item.isdigit():
                raise NekError("cannot parse sequence '%s'" % text)
            out.append(int(item))
    return out


class CollectionOptimisation:
    """Collections::CollectionOptimisation (CollectionOptimisation.cpp:52-316): which ImplementationType each
    (operator, shape, order) uses, from the constructor default and the session's
    <COLLECTIONS DEFAULT="B200" MAXSIZE="..."><OPERATOR TYPE="Helmholtz"><ELEMENT TYPE="H" ORDER="*" IMPTYPE="B200"/>
    block.  `session` is None (the unit tests' dummy session), XML text, a path, or an xml.etree element whose root is
    <NEKTAR>; selection semantics, defaults and error messages follow the reference line by line.  The autotuner
    (SetWithTimings, DEFAULT="auto") needs the reference's other implementations to time against and is not mirrored:
    IsUsingAutotuning() reports the request, the caller decides."""

    _elTypes = {"S": eSegment, "T": eTriangle, "Q": eQuadrilateral, "A": eTetrahedron, "P": ePyramid, "R": ePrism,
                "H": eHexahedron}

    def __init__(self, session=None, defaultType=eNoImpType):
        self.m_setByXml, self.m_autotune, self.m_maxCollSize = False, False, 0
        self.m_defaultType = eIterPerExp if defaultType == eNoImpType else defaultType
        defaults = {(sh, -1): self.m_defaultType for sh in self._elTypes.values()}
        defaultsPhysDeriv = dict(defaults)
        if defaultType == eNoImpType:
            for sh in self._elTypes.values():
                for i in range(1, 5):
                    defaults[(sh, i)] = eStdMat
                defaultsPhysDeriv[(sh, -1)] = eNoCollection
                for i in range(1, 3):
                    defaultsPhysDeriv[(sh, i)] = eSumFac
        self.m_global = {op: dict(defaultsPhysDeriv if op == ePhysDeriv else defaults) for op in range(SIZE_OperatorType)}
        if session is None:
            return
        import xml.etree.ElementTree as ET
        if isinstance(session, str):
            root = ET.fromstring(session) if session.lstrip().startswith("<") else ET.parse(session).getroot()
        else:
            root = session
        if root.tag != "NEKTAR":
            raise NekError("Unable to find NEKTAR tag in file.")
        xmlCol = root.find("COLLECTIONS")
        if xmlCol is None:
            return
        self.m_maxCollSize = int(xmlCol.get("MAXSIZE", 0))
        defaultImpl = xmlCol.get("DEFAULT")
        self.m_defaultType = defaultType
        if defaultType == eNoImpType and defaultImpl:
            self.m_autotune = defaultImpl.lower() == "auto"
            if not self.m_autotune:
                names = [n.lower() for n in ImplementationTypeMap]
                if defaultImpl.lower() not in names[1:]:
                    raise NekError("Unknown default collection scheme: " + defaultImpl)
                self.m_defaultType = names.index(defaultImpl.lower())
                defaults = {(sh, -1): self.m_defaultType for sh in self._elTypes.values()}
                self.m_global = {op: dict(defaults) for op in range(SIZE_OperatorType)}
        for elmt in xmlCol:
            self.m_setByXml = True
            if elmt.tag.upper() != "OPERATOR":
                raise NekError("Only OPERATOR tags are supported inside the COLLECTIONS tag.")
            opType = elmt.get("TYPE")
            if opType is None:
                raise NekError("Missing TYPE in OPERATOR tag.")
            if opType not in OperatorTypeMap:
                raise NekError("Unknown OPERATOR type " + opType + ".")
            ot = OperatorTypeMap.index(opType)
            for elmt2 in elmt:
                if elmt2.tag.upper() != "ELEMENT":
                    raise NekError("Only ELEMENT tags are supported inside the OPERATOR tag.")
                elType = elmt2.get("TYPE")
                if elType is None:
                    raise NekError("Missing TYPE in ELEMENT tag.")
                if elType not in self._elTypes:
                    raise NekError("Unknown element type " + elType + " in ELEMENT tag")
                impType = elmt2.get("IMPTYPE")
                if impType is None:
                    raise NekError("Missing IMPTYPE in ELEMENT tag.")
                if impType not in ImplementationTypeMap:
                    raise NekError("Unknown IMPTYPE type " + impType + ".")
                order = elmt2.get("ORDER")
                if order is None:
                    raise NekError("Missing ORDER in ELEMENT tag.")
                imp, sh = ImplementationTypeMap.index(impType), self._elTypes[elType]
                if order == "*":
                    self.m_global[ot][(sh, -1)] = imp
                else:
                    for o in GenerateSeqVector(order):
                        self.m_global[ot][(sh, o)] = imp

    def GetOperatorImpMap(self, pExp):
        """(shape, number of modes in direction 0) first, then the shape's default, else eNoCollection
        (CollectionOptimisation.cpp:283-316)."""
        shape, nm0 = pExp.DetShapeType(), pExp.GetBasis(0).GetNumModes()
        return {op: table.get((shape, nm0), table.get((shape, -1), eNoCollection)) for op, table in self.m_global.items()}

    def GetDefaultImplementationType(self): return self.m_defaultType
    def GetMaxCollectionSize(self): return self.m_maxCollSize
    def IsUsingAutotuning(self): return self.m_autotune
    def SetByXml(self): return self.m_setByXml


class ExpList:
    """MultiRegions::ExpList reduced to what drives the Collections: CreateCollections (ExpList.cpp:5005-5151) and the
    four call sites that loop over the collections with coefficient / quadrature offsets -- IProductWRTBase
    (:1262-1284), PhysDeriv (:1465-1504), BwdTrans (:1961-1989), GeneralMatrixOp for Helmholtz (:2359-2397).

    `exps` lists the elements in mesh order as (StdExpansion, jac, df, deformed) with jac of 1 | nq values and df of
    ndf x (1 | nq) values (the element's GeomFactors).  Elements are grouped exactly as the reference does: per shape
    (in LibUtilities::ShapeType order), a collection ends where the next element is not contiguous in the coefficient /
    quadrature arrays, differs in nCoeffs / nPhys / deformed-ness, or the collection holds collmax = MAXSIZE (else
    2 x number of elements) members."""

    _shapeOrder = {eSegment: 1, eTriangle: 2, eQuadrilateral: 3, eTetrahedron: 4, ePyramid: 5, ePrism: 6, eHexahedron: 7}

    def __init__(self, exps, session=None):
        self.m_exp, self.m_session = list(exps), session
        self.m_coeff_offset, self.m_phys_offset = [], []
        nc = nq = 0
        for std, _, _, _ in self.m_exp:
            self.m_coeff_offset.append(nc)
            self.m_phys_offset.append(nq)
            nc += std.GetNcoeffs()
            nq += std.GetTotPoints()
        self.m_ncoeffs, self.m_npoints = nc, nq
        self.m_collections, self.m_coll_coeff_offset, self.m_coll_phys_offset = [], [], []

    def GetNcoeffs(self): return self.m_ncoeffs
    def GetTotPoints(self): return self.m_npoints

    def _group(self, collmax):
        """-> list of element-index lists, one per collection, in the reference's order"""
        byShape = {}
        for i, (std, _, _, _) in enumerate(self.m_exp):
            byShape.setdefault(std.DetShapeType(), []).append(i)
        groups = []
        for shape in sorted(byShape, key=lambda sh: self._shapeOrder[sh]):
            idx = byShape[shape]
            cur = [idx[0]]
            for prev, i in zip(idx, idx[1:]):
                sp, si = self.m_exp[prev][0], self.m_exp[i][0]
                split = (self.m_coeff_offset[prev] + si.GetNcoeffs() != self.m_coeff_offset[i] or
                         sp.GetNcoeffs() != si.GetNcoeffs() or
                         self.m_phys_offset[prev] + si.GetTotPoints() != self.m_phys_offset[i] or
                         bool(self.m_exp[prev][3]) != bool(self.m_exp[i][3]) or
                         sp.GetTotPoints() != si.GetTotPoints() or len(cur) >= collmax)
                if split:
                    groups.append(cur)
                    cur = [i]
                else:
                    cur.append(i)
            groups.append(cur)
        return groups

    def CreateCollections(self, ImpType=eNoImpType):
        colOpt = CollectionOptimisation(self.m_session, ImpType)
        collmax = colOpt.GetMaxCollectionSize() if colOpt.GetMaxCollectionSize() > 0 else 2 * len(self.m_exp)
        self.m_collections, self.m_coll_coeff_offset, self.m_coll_phys_offset = [], [], []
        impOfShape = {}  # ExpList.cpp:5043: one GetOperatorImpMap per SHAPE, from the first element of that shape
        for members in self._group(collmax):
            std, _, _, deformed = self.m_exp[members[0]]
            if std.DetShapeType() not in impOfShape:
                impOfShape[std.DetShapeType()] = colOpt.GetOperatorImpMap(std)
            jac = np.concatenate([np.atleast_1d(np.asarray(self.m_exp[i][1], dtype=np.float64)) for i in members])
            ndf = std.dim * std.coordim
            df = np.concatenate([np.asarray(self.m_exp[i][2], dtype=np.float64).reshape(ndf, -1) for i in members], axis=1)
            geom = CoalescedGeomData(np.ascontiguousarray(jac), np.ascontiguousarray(df).reshape(-1), bool(deformed))
            self.m_collections.append(Collection(std, len(members), geom, impOfShape[std.DetShapeType()]))
            self.m_coll_coeff_offset.append(self.m_coeff_offset[members[0]])
            self.m_coll_phys_offset.append(self.m_phys_offset[members[0]])
        return self

    def _spans(self):
        for c, co, po_ in zip(self.m_collections, self.m_coll_coeff_offset, self.m_coll_phys_offset):
            yield c, co, co + c.m_nElmt * c.m_stdExp.GetNcoeffs(), po_, po_ + c.m_nElmt * c.m_stdExp.GetTotPoints()

    def BwdTrans(self, inarray, outarray):
        for c, c0, c1, p0, p1 in self._spans():
            c.ApplyOperator(eBwdTrans, inarray[c0:c1], outarray[p0:p1])

    def IProductWRTBase(self, inarray, outarray):
        for c, c0, c1, p0, p1 in self._spans():
            c.ApplyOperator(eIProductWRTBase, inarray[p0:p1], outarray[c0:c1])

    def PhysDeriv(self, inarray, out_d0, out_d1=None, out_d2=None):
        for c, c0, c1, p0, p1 in self._spans():
            outs = [o[p0:p1] for o in (out_d0, out_d1, out_d2)[:c.m_stdExp.coordim] if o is not None]
            c.ApplyOperator(ePhysDeriv, inarray[p0:p1], *outs)

    def GeneralMatrixOp_Helmholtz(self, inarray, outarray, factors):
        """ExpList::GeneralMatrixOp for a Helmholtz matrix key (ExpList.cpp:2359-2397): local coefficients in and out"""
        for c, c0, c1, p0, p1 in self._spans():
            c.ApplyOperator(eHelmholtz, inarray[c0:c1], outarray[c0:c1], factors=factors)


# ----------------------------------------------------------------------------- AssemblyMap
