"""ctypes binding of libnekmf_b200.so (include/nekmf_b200.h) plus a thin host-side mirror of the
reference's Collections interface for the matrix-free path.

Names follow library/Collections (Operator.h:65-191, Collection.h:53-110):

    OperatorType / ImplementationType enums (with the new ``eB200`` before SIZE_ImplementationType)
    StdExpansion(shape, nummodes, numpoints)      <- StdRegions::StdExpansion: GetBasis(i) tables
    CoalescedGeomData(jac, df, deformed)          <- Collections::CoalescedGeomData (GetJac / GetDerivFactors)
    Collection(stdexp, nElmt, geom, impTypes)     <- Collections::Collection: Initialise / ApplyOperator
    AssemblyMap(localToGlobal, sign, nGlobal)     <- AssemblyMapCG: GlobalToLocal / Assemble
    HelmholtzCG                                   <- NekLinSysIterCG + GlobalLinSysIterativeFull mat-vec

Arrays may be numpy arrays (host: copies happen inside the C call, like a literal drop-in for
Array<OneD, NekDouble>) or torch CUDA tensors (device resident: no copies).  There is no CPU
compute path here: every operator call goes to the CUDA library and raises NekError if the
library or a GPU is missing.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnekmf_b200.so")

# LibUtilities::ShapeType subset, ABI numbering
eQuadrilateral, eTriangle, eHexahedron, ePrism, ePyramid, eTetrahedron, eSegment = 0, 1, 2, 3, 4, 5, 6
ShapeTypeMap = {0: "Quadrilateral", 1: "Triangle", 2: "Hexahedron", 3: "Prism", 4: "Pyramid", 5: "Tetrahedron", 6: "Segment"}
# Collections::OperatorType (Operator.h:65-73)
eBwdTrans, eHelmholtz, eIProductWRTBase, eIProductWRTDerivBase, ePhysDeriv, SIZE_OperatorType = 0, 1, 2, 3, 4, 5
OperatorTypeMap = ["BwdTrans", "Helmholtz", "IProductWRTBase", "IProductWRTDerivBase", "PhysDeriv"]
# Collections::ImplementationType (Operator.h:84-103) + the new entry
(eNoImpType, eNoCollection, eIterPerExp, eStdMat, eSumFac, eMatrixFree, eB200,
 SIZE_ImplementationType) = range(8)
ImplementationTypeMap = ["NoImplementationType", "NoCollection", "IterPerExp", "StdMat", "SumFac", "MatrixFree",
                         "B200"]
# basis / points types
eModified_A, eModified_B, eModified_C, eModifiedPyr_C = 0, 1, 2, 3
eGaussLobattoLegendre, eGaussRadauMAlpha1Beta0, eGaussRadauMAlpha2Beta0 = 0, 1, 2
HOST, DEVICE = 0, 1
ERR_NOCONVERGE = 6  # nekmf_status
eFactorLambda = "FactorLambda"

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

EXPORTS = [
    "nekmf_abi_version", "nekmf_last_error", "nekmf_device_count", "nekmf_set_device", "nekmf_launch_count",
    "nekmf_malloc_device", "nekmf_free_device", "nekmf_malloc_pinned", "nekmf_free_pinned", "nekmf_host_register",
    "nekmf_host_unregister", "nekmf_memcpy_h2d",
    "nekmf_memcpy_d2h", "nekmf_memset_device", "nekmf_sync", "nekmf_points", "nekmf_basis_rows", "nekmf_basis",
    "nekmf_op_create", "nekmf_op_set_geom", "nekmf_op_set_lambda", "nekmf_op_apply", "nekmf_op_set_stream",
    "nekmf_op_ncoeff", "nekmf_op_nphys", "nekmf_op_kernel_name", "nekmf_op_enable_timing", "nekmf_op_last_ms",
    "nekmf_op_destroy", "nekmf_map_create", "nekmf_map_global_to_local", "nekmf_map_assemble", "nekmf_map_destroy",
    "nekmf_comm_unique_id", "nekmf_comm_create", "nekmf_comm_transport", "nekmf_comm_destroy", "nekmf_exchange_create",
    "nekmf_exchange_add", "nekmf_exchange_destroy", "nekmf_cg_create", "nekmf_cg_solve", "nekmf_cg_matvec",
    "nekmf_cg_last_loop", "nekmf_cg_destroy", "nekmf_helmsolve_create", "nekmf_helmsolve", "nekmf_helmsolve_last_ms",
    "nekmf_helmsolve_last_phases",
    "nekmf_helmsolve_destroy", "nekmf_op_diagonal", "nekmf_cg_set_jacobi",
]


class NekError(RuntimeError):
    """ErrorUtil::NekError analogue (LibUtilities/BasicUtils/ErrorUtil.hpp:88-186)."""


_lib = None


def lib():
    """Load libnekmf_b200.so; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NekError("libnekmf_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C ithaca-sem_b200/csrc`")
        L = C.CDLL(os.environ.get("NEKMF_B200_LIB", LIB_PATH))  # the override is for A/B builds of the same library
        L.nekmf_last_error.restype = C.c_char_p
        L.nekmf_op_kernel_name.restype = C.c_char_p
        L.nekmf_op_kernel_name.argtypes = [_vp]
        L.nekmf_launch_count.restype = C.c_longlong
        L.nekmf_points.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
        L.nekmf_basis_rows.argtypes = [C.c_int, C.c_int]
        L.nekmf_basis.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        pp = C.POINTER(_dp)
        L.nekmf_op_create.argtypes = [C.c_int, C.c_int, _ip, _ip, _ip, _ip, pp, pp, pp, pp, pp, C.c_int, C.c_int,
                                      C.c_int, C.POINTER(_vp)]
        L.nekmf_op_set_geom.argtypes = [_vp, _vp, _vp, C.c_int]
        L.nekmf_op_set_lambda.argtypes = [_vp, C.c_double]
        L.nekmf_op_apply.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int]
        L.nekmf_op_set_stream.argtypes = [_vp, _vp]
        L.nekmf_op_ncoeff.argtypes = [_vp]
        L.nekmf_op_nphys.argtypes = [_vp]
        L.nekmf_op_enable_timing.argtypes = [_vp, C.c_int]
        L.nekmf_op_last_ms.argtypes = [_vp, C.POINTER(C.c_float)]
        L.nekmf_op_destroy.argtypes = [_vp]
        L.nekmf_map_create.argtypes = [C.c_int, C.c_int, _ip, _dp, C.POINTER(_vp)]
        L.nekmf_map_global_to_local.argtypes = [_vp, _vp, _vp, C.c_int, _vp]
        L.nekmf_map_assemble.argtypes = [_vp, _vp, _vp, C.c_int, _vp]
        L.nekmf_map_destroy.argtypes = [_vp]
        L.nekmf_comm_unique_id.argtypes = [C.c_char_p]
        L.nekmf_comm_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(_vp)]
        L.nekmf_comm_transport.argtypes = [_vp]
        L.nekmf_comm_destroy.argtypes = [_vp]
        L.nekmf_exchange_create.argtypes = [_vp, C.c_int, C.c_int, _ip, _ip, _ip, C.POINTER(_vp)]
        L.nekmf_exchange_add.argtypes = [_vp, _vp, _vp]
        L.nekmf_exchange_destroy.argtypes = [_vp]
        L.nekmf_cg_create.argtypes = [_vp, _vp, _vp, _vp, C.c_int, _dp, _dp, C.POINTER(_vp)]
        L.nekmf_cg_solve.argtypes = [_vp, _vp, _vp, C.c_int, C.c_double, C.c_int, _ip, _dp]
        L.nekmf_cg_matvec.argtypes = [_vp, _vp, _vp]
        L.nekmf_cg_last_loop.argtypes = [_vp, C.POINTER(C.c_float), _ip]
        L.nekmf_cg_destroy.argtypes = [_vp]
        L.nekmf_op_diagonal.argtypes = [_vp, _vp, C.c_int]
        L.nekmf_cg_set_jacobi.argtypes = [_vp]
        L.nekmf_helmsolve_create.argtypes = [_vp, _vp, _vp, C.POINTER(_vp)]
        L.nekmf_helmsolve.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_double, C.c_int, _ip, _dp]
        L.nekmf_helmsolve_last_ms.argtypes = [_vp, C.POINTER(C.c_float)]
        L.nekmf_helmsolve_last_phases.argtypes = [_vp, C.POINTER(C.c_float)]
        L.nekmf_helmsolve_destroy.argtypes = [_vp]
        L.nekmf_malloc_device.argtypes = [C.POINTER(_vp), C.c_size_t]
        L.nekmf_free_device.argtypes = [_vp]
        L.nekmf_malloc_pinned.argtypes = [C.POINTER(_vp), C.c_size_t]
        L.nekmf_free_pinned.argtypes = [_vp]
        L.nekmf_host_register.argtypes = [_vp, C.c_size_t]
        L.nekmf_host_unregister.argtypes = [_vp]
        L.nekmf_memcpy_h2d.argtypes = [_vp, _vp, C.c_size_t]
        L.nekmf_memcpy_d2h.argtypes = [_vp, _vp, C.c_size_t]
        L.nekmf_memset_device.argtypes = [_vp, C.c_int, C.c_size_t]
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise NekError("%s failed (status %d): %s" % (what, rc, lib().nekmf_last_error().decode()))


def device_count():
    return int(lib().nekmf_device_count())


def launch_count():
    return int(lib().nekmf_launch_count())


def _np_p(a):
    return a.ctypes.data_as(_dp)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    """(pointer, memkind, keepalive) of a numpy array or a torch tensor."""
    if x is None:
        return None, None, None
    if _is_torch(x):
        if x.dtype.is_floating_point and x.element_size() != 8:
            raise NekError("arrays must be float64")
        if not x.is_contiguous():
            raise NekError("arrays must be contiguous")
        return C.c_void_p(x.data_ptr()), (DEVICE if x.is_cuda else HOST), x
    a = x
    if not isinstance(a, np.ndarray) or a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
        raise NekError("arrays must be contiguous float64 numpy arrays or torch tensors")
    return C.c_void_p(a.ctypes.data), HOST, a


def _kind(*xs):
    kinds = set(_ptr(x)[1] for x in xs if x is not None)
    if len(kinds) != 1:
        raise NekError("all arrays of one call must live in the same memory space")
    return kinds.pop()


# ----------------------------------------------------------------------------- Foundations
def points(ptype, n):
    """PointsManager()[PointsKey(n, ptype)] -> (z, w, D) with D[k*n+i] = dh_k/dz(z_i)."""
    z, w, D = np.zeros(n), np.zeros(n), np.zeros(n * n)
    check(lib().nekmf_points(ptype, n, _np_p(z), _np_p(w), _np_p(D)), "nekmf_points")
    return z, w, D


class Basis:
    """LibUtilities::Basis: GetBdata/GetDbdata/GetD/GetZ/GetW/GetNumModes/GetNumPoints."""

    def __init__(self, btype, nm, ptype, nq):
        self.btype, self.nm, self.ptype, self.nq = btype, nm, ptype, nq
        self.Z, self.W, self.D = points(ptype, nq)
        self.rows = int(lib().nekmf_basis_rows(btype, nm))
        self.bdata, self.dbdata = np.zeros(self.rows * nq), np.zeros(self.rows * nq)
        check(lib().nekmf_basis(btype, nm, nq, _np_p(self.Z), _np_p(self.D), _np_p(self.bdata), _np_p(self.dbdata)),
              "nekmf_basis")

    def GetBdata(self): return self.bdata
    def GetDbdata(self): return self.dbdata
    def GetD(self): return self.D
    def GetZ(self): return self.Z
    def GetW(self): return self.W
    def GetNumModes(self): return self.nm
    def GetNumPoints(self): return self.nq
    def GetBasisType(self): return self.btype
    def GetPointsType(self): return self.ptype


def num_coeffs(shape, nm):
    """LibUtilities::StdXxxData::getNumberOfCoefficients (BasicUtils/ShapeType.hpp:111-337)."""
    return {eQuadrilateral: nm * nm, eTriangle: nm * (nm + 1) // 2, eHexahedron: nm ** 3,
            ePrism: nm * nm * (nm + 1) // 2, eTetrahedron: nm * (nm + 1) * (nm + 2) // 6,
            ePyramid: nm * (nm + 1) * (2 * nm + 1) // 6, eSegment: nm}[shape]


class StdExpansion:
    """The part of StdRegions::StdExpansion the Collections operators read: shape, the per-direction
    Basis objects with Nektar's default point distributions (SpatialDomains/MeshGraph.cpp:1609-1762:
    nq = nm+1 Gauss-Lobatto-Legendre in tensor directions, Gauss-Radau in collapsed ones)."""

    def __init__(self, shape, nummodes, numpoints=None, coordim=None):
        nm = nummodes
        nq0 = numpoints if numpoints is not None else nm + 1
        self.shape, self.nm = shape, nm
        self.dim = 1 if shape == eSegment else (2 if shape in (eQuadrilateral, eTriangle) else 3)
        # GetCoordim(): space dimension; differs from the element dimension only for segments (1..3)
        self.coordim = self.dim if coordim is None else int(coordim)
        bt = [eModified_A] * 3
        pt = [eGaussLobattoLegendre] * 3
        nq = [nq0] * 3
        if shape == eTriangle:
            bt[1], pt[1], nq[1] = eModified_B, eGaussRadauMAlpha1Beta0, nq0 - 1
        elif shape == ePrism:
            bt[2], pt[2], nq[2] = eModified_B, eGaussRadauMAlpha1Beta0, nq0 - 1
        elif shape == ePyramid:
            # nq2 = nq0 - 1 is what the MatrixFree operators require (Helmholtz.h:1519-1525)
            bt[2], pt[2], nq[2] = eModifiedPyr_C, eGaussRadauMAlpha2Beta0, nq0 - 1
        elif shape == eTetrahedron:
            bt[1], pt[1], nq[1] = eModified_B, eGaussRadauMAlpha1Beta0, nq0 - 1
            bt[2], pt[2], nq[2] = eModified_C, eGaussRadauMAlpha2Beta0, nq0 - 1
        elif shape not in (eQuadrilateral, eHexahedron, eSegment):
            raise NekError("shape %s not supported" % ShapeTypeMap.get(shape, shape))
        self.basis = [Basis(bt[d], nm, pt[d], nq[d]) for d in range(self.dim)]
        self.nq = nq[:self.dim]

    def GetBasis(self, d): return self.basis[d]
    def DetShapeType(self): return self.shape
    def GetNcoeffs(self): return num_coeffs(self.shape, self.nm)

    def GetTotPoints(self):
        n = 1
        for q in self.nq:
            n *= q
        return n


class CoalescedGeomData:
    """Collections::CoalescedGeomData (CoalescedGeomData.cpp:53-113, 251-313): jac [nElmt] | [nElmt*nq],
    derivative factors [dim*coordim][nElmt | nElmt*nq], deformed flag of the collection."""

    def __init__(self, jac, df, deformed):
        self.jac, self.df, self.deformed = jac, df, bool(deformed)

    def GetJac(self): return self.jac
    def GetDerivFactors(self): return self.df
    def IsDeformed(self): return self.deformed


# ----------------------------------------------------------------------------- Operator
class Operator:
    """Collections::Operator registered under (shape, opType, eB200, nodal=False).  __call__ mirrors
    Operator::operator()(input, output0, output1, output2, wsp, factors) (Operator.h:122-137)."""

    def __init__(self, stdexp, nElmt, geom, optype):
        self.stdexp, self.nElmt, self.optype = stdexp, int(nElmt), optype
        self.h = _vp()
        dim = stdexp.dim
        i3 = C.c_int * 3
        nm = i3(*([stdexp.nm] * dim + [1] * (3 - dim)))
        nq = i3(*(stdexp.nq + [1] * (3 - dim)))
        bt = i3(*([b.btype for b in stdexp.basis] + [0] * (3 - dim)))
        pt = i3(*([b.ptype for b in stdexp.basis] + [0] * (3 - dim)))

        def arr(name):
            a = (_dp * 3)()
            for d in range(dim):
                a[d] = _np_p(getattr(stdexp.basis[d], name))
            return a

        deformed = bool(geom.IsDeformed()) if geom is not None else False
        check(lib().nekmf_op_create(stdexp.shape, optype, nm, nq, bt, pt, arr("bdata"), arr("dbdata"), arr("D"),
                                    arr("Z"), arr("W"), self.nElmt, int(deformed), stdexp.coordim, C.byref(self.h)),
              "nekmf_op_create")
        self.m_isDeformed = deformed
        self.ncoeff = int(lib().nekmf_op_ncoeff(self.h))
        self.nphys = int(lib().nekmf_op_nphys(self.h))
        if geom is not None and (geom.GetJac() is not None or geom.GetDerivFactors() is not None):
            pj, kj, _ = _ptr(geom.GetJac())
            pd, kd, _ = _ptr(geom.GetDerivFactors())
            kind = _kind(geom.GetJac(), geom.GetDerivFactors())
            npt = self.nElmt * (self.nphys if deformed else 1)
            for a, n, nmq in ((geom.GetJac(), npt, "jac"), (geom.GetDerivFactors(), npt * dim * stdexp.coordim, "df")):
                if a is not None and (a.numel() if _is_torch(a) else a.size) != n:
                    raise NekError("CoalescedGeomData: %s has %d entries, expected %d" % (
                        nmq, a.numel() if _is_torch(a) else a.size, n))
            check(lib().nekmf_op_set_geom(self.h, pj, pd, kind), "nekmf_op_set_geom")

    @property
    def kernel_name(self):
        return lib().nekmf_op_kernel_name(self.h).decode()

    def GetNumElmt(self):
        return self.nElmt

    def SetLambda(self, lam):
        check(lib().nekmf_op_set_lambda(self.h, float(lam)), "nekmf_op_set_lambda")

    def diagonal(self, out=None):
        """diagonal of every elemental Helmholtz matrix [nElmt*ncoeff] (nekmf_op_diagonal); numpy unless `out` given"""
        if out is None:
            out = np.zeros(self.nElmt * self.ncoeff)
        p, kind, _ = _ptr(out)
        check(lib().nekmf_op_diagonal(self.h, p, kind), "nekmf_op_diagonal")
        return out

    def enable_timing(self, on=True):
        check(lib().nekmf_op_enable_timing(self.h, int(on)), "nekmf_op_enable_timing")

    def last_ms(self):
        ms = C.c_float(-1.0)
        check(lib().nekmf_op_last_ms(self.h, C.byref(ms)), "nekmf_op_last_ms")
        return ms.value

    def set_stream(self, stream_ptr):
        check(lib().nekmf_op_set_stream(self.h, C.c_void_p(stream_ptr)), "nekmf_op_set_stream")

    def _size_check(self, arrs, coeff, what):
        n = self.nElmt * (self.ncoeff if coeff else self.nphys)
        for a in arrs:
            sz = a.numel() if _is_torch(a) else a.size
            if sz < n:
                raise NekError("%s array too small: %d < %d" % (what, sz, n))

    def apply(self, ins, outs):
        """ins / outs: lists of arrays as the operator type requires."""
        cin = self.optype in (eBwdTrans, eHelmholtz)
        cout = self.optype not in (eBwdTrans, ePhysDeriv)
        self._size_check(ins, cin, "input")
        self._size_check(outs, cout, "output")
        kind = _kind(*(list(ins) + list(outs)))
        pi = [_ptr(a)[0] for a in ins] + [None] * (3 - len(ins))
        po = [_ptr(a)[0] for a in outs] + [None] * (3 - len(outs))
        check(lib().nekmf_op_apply(self.h, pi[0], pi[1], pi[2], po[0], po[1], po[2], kind), "nekmf_op_apply")

    def __call__(self, input, output0, output1=None, output2=None, wsp=None, factors=None):
        if isinstance(input, int):
            # Operator::operator()(dir, input, output, wsp): only PhysDeriv implements it, by computing
            # every direction and copying one (Collections/PhysDeriv.cpp:323-341)
            return self.apply_dir(input, output0, output1)
        if self.optype == eHelmholtz:
            if factors is None or eFactorLambda not in factors:
                raise NekError("Helmholtz: factors must contain eFactorLambda (Collections/Helmholtz.cpp:412-414)")
            self.SetLambda(factors[eFactorLambda])
        if self.optype == eIProductWRTDerivBase:
            # reference calling convention: (in0, in1, out, in2?) for 2-D / (in0, in1, in2, out) for 3-D
            # (Collections/IProductWRTDerivBase.cpp:285-330): the LAST array is the output
            arrs = [a for a in (input, output0, output1, output2) if a is not None]
            return self.apply(arrs[:-1], [arrs[-1]])
        outs = [a for a in (output0, output1, output2) if a is not None]
        return self.apply([input], outs)

    def apply_dir(self, dir, input, output):
        if self.optype != ePhysDeriv:
            raise NekError("%s: operator()(dir, ...) is not valid for this operator" % OperatorTypeMap[self.optype])
        dim = self.stdexp.coordim
        if not 0 <= dir < dim:
            raise NekError("PhysDeriv: direction %d out of range" % dir)
        n = self.nElmt * self.nphys
        if _is_torch(output):
            import torch
            tmp = [torch.empty(n, dtype=torch.float64, device=output.device) for _ in range(dim)]
        else:
            tmp = [np.empty(n) for _ in range(dim)]
        tmp[dir] = output
        self.apply([input], tmp)

    def __del__(self):
        try:
            if self.h:
                lib().nekmf_op_destroy(self.h)
                self.h = None
        except Exception:
            pass


def SetFixedImpType(defaultType):
    """Collections::SetFixedImpType (Operator.cpp:128-138)."""
    return {op: defaultType for op in range(SIZE_OperatorType)}


class Collection:
    """Collections::Collection (Collection.h:53-110, Collection.cpp:46-87): lazy Initialise(opType), then
    ApplyOperator.  Only eB200 is registered in this library; any other ImplementationType raises, as
    the reference's factory does for an unregistered key."""

    def __init__(self, stdexp, nElmt, geom, impTypes=None):
        self.m_stdExp, self.m_nElmt, self.m_geomData = stdexp, nElmt, geom
        self.m_impTypes = impTypes if impTypes is not None else SetFixedImpType(eB200)
        self.m_ops = {}

    def Initialise(self, opType):
        if opType in self.m_ops:
            return
        imp = self.m_impTypes[opType]
        if imp != eB200:
            raise NekError("no operator registered for key (%s, %s, %s)" % (
                ShapeTypeMap[self.m_stdExp.shape], OperatorTypeMap[opType], ImplementationTypeMap[imp]))
        self.m_ops[opType] = Operator(self.m_stdExp, self.m_nElmt, self.m_geomData, opType)

    def HasOperator(self, opType):
        return opType in self.m_ops

    def ApplyOperator(self, opType, *args, **kw):
        self.Initialise(opType)
        return self.m_ops[opType](*args, **kw)


# ----------------------------------------------------------------------------- ExpList (the callers of the path)
class AssemblyMap:
    """AssemblyMapCG local<->global (AssemblyMapCG.cpp:2853-2923)."""

    def __init__(self, localToGlobal, nGlobal, sign=None):
        l2g = np.ascontiguousarray(localToGlobal, dtype=np.int32)
        self.nLocal, self.nGlobal = int(l2g.size), int(nGlobal)
        self.h = _vp()
        sg = None
        if sign is not None:
            sg = np.ascontiguousarray(sign, dtype=np.float64)
        check(lib().nekmf_map_create(self.nLocal, self.nGlobal, l2g.ctypes.data_as(_ip),
                                     _np_p(sg) if sg is not None else None, C.byref(self.h)), "nekmf_map_create")

    def GlobalToLocal(self, glob, loc, stream=None):
        kind = _kind(glob, loc)
        check(lib().nekmf_map_global_to_local(self.h, _ptr(glob)[0], _ptr(loc)[0], kind, C.c_void_p(stream or 0)),
              "nekmf_map_global_to_local")

    def Assemble(self, loc, glob, stream=None):
        kind = _kind(glob, loc)
        check(lib().nekmf_map_assemble(self.h, _ptr(loc)[0], _ptr(glob)[0], kind, C.c_void_p(stream or 0)),
              "nekmf_map_assemble")

    def __del__(self):
        try:
            if self.h:
                lib().nekmf_map_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Comm:
    """One rank per GPU NCCL communicator; the unique id travels through torch.distributed."""

    def __init__(self, rank, nranks, unique_id):
        self.h = _vp()
        self.rank, self.nranks = rank, nranks
        check(lib().nekmf_comm_create(unique_id, rank, nranks, C.byref(self.h)), "nekmf_comm_create")

    @property
    def transport(self):
        """"p2p": kernels store into / spin on CUDA-IPC mapped peer windows over NVLink; "nccl": NCCL calls"""
        return "p2p" if lib().nekmf_comm_transport(self.h) else "nccl"

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        check(lib().nekmf_comm_unique_id(buf), "nekmf_comm_unique_id")
        return buf.raw

    @classmethod
    def from_torch_distributed(cls):
        import torch
        import torch.distributed as dist
        rank, n = dist.get_rank(), dist.get_world_size()
        ids = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return cls(rank, n, ids[0])

    def __del__(self):
        try:
            if self.h:
                lib().nekmf_comm_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Exchange:
    """Gs::Gather(gs_add) replacement for partition-interface DOFs."""

    def __init__(self, comm, peers, lists, nGlobal):
        self.h = _vp()
        self.comm = comm
        n = len(peers)
        offs = np.zeros(n + 1, dtype=np.int32)
        for i, l in enumerate(lists):
            offs[i + 1] = offs[i] + len(l)
        idx = np.ascontiguousarray(np.concatenate([np.asarray(l, dtype=np.int32) for l in lists])
                                   if n else np.zeros(0, dtype=np.int32))
        pr = np.ascontiguousarray(peers, dtype=np.int32)
        check(lib().nekmf_exchange_create(comm.h if comm is not None else None, int(nGlobal), n, pr.ctypes.data_as(_ip),
                                          offs.ctypes.data_as(_ip), idx.ctypes.data_as(_ip), C.byref(self.h)),
              "nekmf_exchange_create")

    def add(self, glob, stream=None):
        check(lib().nekmf_exchange_add(self.h, _ptr(glob)[0], C.c_void_p(stream or 0)), "nekmf_exchange_add")

    def __del__(self):
        try:
            if self.h:
                lib().nekmf_exchange_destroy(self.h)
                self.h = None
        except Exception:
            pass


class HelmholtzCG:
    """NekLinSysIterCG::DoConjugateGradient on A = Assemble o Helmholtz o GlobalToLocal."""

    def __init__(self, helm_op, amap, nDir, invdiag=None, exchange=None, comm=None, ownerMask=None):
        self.h = _vp()
        self.op, self.map, self.ex, self.comm = helm_op, amap, exchange, comm
        iv = np.ascontiguousarray(invdiag, dtype=np.float64) if invdiag is not None else None
        om = np.ascontiguousarray(ownerMask, dtype=np.float64) if ownerMask is not None else None
        check(lib().nekmf_cg_create(helm_op.h, amap.h, exchange.h if exchange is not None else None,
                                    comm.h if comm is not None else None, int(nDir),
                                    _np_p(iv) if iv is not None else None, _np_p(om) if om is not None else None,
                                    C.byref(self.h)), "nekmf_cg_create")

    def solve(self, rhs, x, tol=1e-9, maxiter=5000, raise_on_maxiter=True):
        """-> (m_totalIterations, final r.r).  Reaching the iteration cap raises NekError like the reference's
        NEKERROR(efatal, "Exceeded maximum number of iterations") (NekLinSysIterCG.cpp:190-203) unless
        raise_on_maxiter is False (fixed-iteration timing runs)."""
        its, eps = C.c_int(0), C.c_double(0.0)
        kind = _kind(rhs, x)
        rc = lib().nekmf_cg_solve(self.h, _ptr(rhs)[0], _ptr(x)[0], kind, float(tol), int(maxiter), C.byref(its),
                                  C.byref(eps))
        if rc == ERR_NOCONVERGE and not raise_on_maxiter:
            return its.value, eps.value
        check(rc, "nekmf_cg_solve")
        return its.value, eps.value

    def set_jacobi(self):
        """matrix-free diagonal preconditioner computed, assembled, exchanged and inverted on the device"""
        check(lib().nekmf_cg_set_jacobi(self.h), "nekmf_cg_set_jacobi")

    def matvec(self, w, s):
        check(lib().nekmf_cg_matvec(self.h, _ptr(w)[0], _ptr(s)[0]), "nekmf_cg_matvec")

    def last_loop(self):
        """-> (ms, iterations): device time of the iteration loop of the last solve (CUDA events)"""
        ms, n = C.c_float(-1.0), C.c_int(0)
        check(lib().nekmf_cg_last_loop(self.h, C.byref(ms), C.byref(n)), "nekmf_cg_last_loop")
        return ms.value, n.value

    def __del__(self):
        try:
            if self.h:
                lib().nekmf_cg_destroy(self.h)
                self.h = None
        except Exception:
            pass


class HelmSolver:
    """ContField::v_HelmSolve (MultiRegions/ContField.cpp:878-945) + the BwdTrans the solvers run next, as one
    device-resident chain: forcing at the quadrature points in, local coefficients (and physical values) out;
    host arrays cross PCIe once each way.  `cg` carries the Helmholtz operator, map, exchange, preconditioner."""

    def __init__(self, cg, iprod_op, bwd_op=None):
        self.h = _vp()
        self.cg, self.iprod, self.bwd = cg, iprod_op, bwd_op
        check(lib().nekmf_helmsolve_create(cg.h, iprod_op.h, bwd_op.h if bwd_op is not None else None, C.byref(self.h)),
              "nekmf_helmsolve_create")

    def HelmSolve(self, forcing, inout, phys_out=None, tol=1e-9, maxiter=5000, raise_on_maxiter=True):
        """inout: local coefficients with the Dirichlet values / initial guess on entry.  -> (iterations, final r.r)"""
        its, eps = C.c_int(0), C.c_double(0.0)
        arrs = [forcing, inout] + ([phys_out] if phys_out is not None else [])
        kind = _kind(*arrs)
        rc = lib().nekmf_helmsolve(self.h, _ptr(forcing)[0], _ptr(inout)[0],
                                   _ptr(phys_out)[0] if phys_out is not None else None, kind, float(tol), int(maxiter),
                                   C.byref(its), C.byref(eps))
        if not (rc == ERR_NOCONVERGE and not raise_on_maxiter):
            check(rc, "nekmf_helmsolve")
        return its.value, eps.value

    def last_ms(self):
        ms = C.c_float(-1.0)
        check(lib().nekmf_helmsolve_last_ms(self.h, C.byref(ms)), "nekmf_helmsolve_last_ms")
        return ms.value

    def last_phases(self):
        """-> [h2d, IProductWRTBase + lift + Assemble, CG, GlobalToLocal + BwdTrans, d2h] in ms of the last call"""
        ms = (C.c_float * 5)()
        check(lib().nekmf_helmsolve_last_phases(self.h, ms), "nekmf_helmsolve_last_phases")
        return [float(v) for v in ms]

    def __del__(self):
        try:
            if self.h:
                lib().nekmf_helmsolve_destroy(self.h)
                self.h = None
        except Exception:
            pass
