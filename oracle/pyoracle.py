"""ctypes bindings for the CPU checkers.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.

  Oracle / Elem : oracle/libmforacle.so  (plain-C restatement, oracle/mf_oracle.c)
  Ref           : oracle/_ref/libnekref_{scalar,avx2}.so  (the reference's own
                  MatrixFreeOps kernel headers + Polylib.cpp compiled in place; built by
                  oracle/Makefile where /root/reference exists, else used prebuilt)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
QUAD, TRI, HEX, PRISM, PYR, TET, SEG = 0, 1, 2, 3, 4, 5, 6
SHAPE_NAMES = {QUAD: "Quad", TRI: "Tri", HEX: "Hex", PRISM: "Prism", PYR: "Pyr", TET: "Tet", SEG: "Seg"}
OP_BWD, OP_HELM, OP_IPROD, OP_IPWDB, OP_PHYSDERIV = 0, 1, 2, 3, 4

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
ELOP = C.CFUNCTYPE(None, C.c_void_p, _dp, _dp)  # mfo_elop_fn: (ctx, in, out)


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def build(target="all"):
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(HERE, "libmforacle.so")
        if not os.path.exists(path):
            build("oracle")
        L = C.CDLL(path)
        L.mfo_create.restype = C.c_void_p
        L.mfo_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.mfo_destroy.argtypes = [C.c_void_p]
        for f in ("mfo_dim", "mfo_nmtot", "mfo_nqtot"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.mfo_set_coordim.argtypes = [C.c_void_p, C.c_int]
        for f in ("mfo_nq", "mfo_ptype", "mfo_btype", "mfo_brows"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        L.mfo_table.restype = _dp
        L.mfo_table.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.mfo_bwdtrans.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.mfo_iproduct.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp]
        L.mfo_physderiv.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.mfo_helmholtz.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_double, _dp, _dp]
        L.mfo_iproductwrtderivbase.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
        L.mfo_global_to_local.argtypes = [C.c_int, _ip, _dp, _dp, _dp]
        L.mfo_assemble.argtypes = [C.c_int, C.c_int, _ip, _dp, _dp, _dp]
        L.mfo_cg_helmholtz.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_int, C.c_int,
                                       C.c_int, _ip, _dp, _dp, _dp, _dp, C.c_double, C.c_int, _dp]
        L.mfo_cg_helmholtz.restype = C.c_int
        L.mfo_set_threads.argtypes = [C.c_int]
        L.mfo_chain_create.restype = C.c_void_p
        L.mfo_chain_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _ip, _dp, _dp]
        L.mfo_chain_destroy.argtypes = [C.c_void_p]
        L.mfo_chain_helmsolve.restype = C.c_int
        L.mfo_chain_helmsolve.argtypes = [C.c_void_p, ELOP, C.c_void_p, ELOP, C.c_void_p, ELOP, C.c_void_p, _dp, _dp, _dp,
                                          C.c_double, C.c_int, _dp]
        L.mfo_points.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
        L.mfo_basis.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        L.mfo_basis_rows.argtypes = [C.c_int, C.c_int]
        L.mfo_jacobfd.argtypes = [C.c_int, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double]
        _oracle = L
    return _oracle


def set_threads(n):
    oracle_lib().mfo_set_threads(int(n))


def max_threads():
    return int(oracle_lib().mfo_max_threads())


def points(ptype, n):
    L = oracle_lib()
    z, w, D = np.zeros(n), np.zeros(n), np.zeros(n * n)
    L.mfo_points(ptype, n, _p(z), _p(w), _p(D))
    return z, w, D


class Elem:
    """Tables + operators of one (shape, nm, nq0) expansion with Nektar's default points."""

    def __init__(self, shape, nm, nq0, coordim=None):
        self.L = oracle_lib()
        self.shape, self.nm, self.nq0 = shape, nm, nq0
        self.h = self.L.mfo_create(shape, nm, nq0)
        if not self.h:
            raise ValueError("unsupported shape")
        self.dim = self.L.mfo_dim(self.h)
        self.coordim = self.dim if coordim is None else int(coordim)  # segments: 1..3 space dimensions
        if self.coordim != self.dim:
            self.L.mfo_set_coordim(self.h, self.coordim)
        self.nmTot = self.L.mfo_nmtot(self.h)
        self.nqTot = self.L.mfo_nqtot(self.h)
        self.nq = [self.L.mfo_nq(self.h, d) for d in range(self.dim)]
        self.ptype = [self.L.mfo_ptype(self.h, d) for d in range(self.dim)]
        self.btype = [self.L.mfo_btype(self.h, d) for d in range(self.dim)]
        self.brows = [self.L.mfo_brows(self.h, d) for d in range(self.dim)]
        self.ndf = self.dim * self.coordim

        def tab(d, which, n):
            return np.ctypeslib.as_array(self.L.mfo_table(self.h, d, which), shape=(n,)).copy()

        self.bdata = [tab(d, 0, self.brows[d] * self.nq[d]) for d in range(self.dim)]
        self.dbdata = [tab(d, 1, self.brows[d] * self.nq[d]) for d in range(self.dim)]
        self.D = [tab(d, 2, self.nq[d] ** 2) for d in range(self.dim)]
        self.Z = [tab(d, 3, self.nq[d]) for d in range(self.dim)]
        self.W = [tab(d, 4, self.nq[d]) for d in range(self.dim)]

    def __del__(self):
        try:
            self.L.mfo_destroy(self.h)
        except Exception:
            pass

    def bwdtrans(self, nel, x):
        out = np.zeros(nel * self.nqTot)
        self.L.mfo_bwdtrans(self.h, nel, _p(x), _p(out))
        return out

    def iproduct(self, nel, deformed, jac, x):
        out = np.zeros(nel * self.nmTot)
        self.L.mfo_iproduct(self.h, nel, int(deformed), _p(jac), _p(x), _p(out))
        return out

    def physderiv(self, nel, deformed, df, x):
        outs = [np.zeros(nel * self.nqTot) for _ in range(self.coordim)]
        o1 = _p(outs[1]) if self.coordim >= 2 else None
        o2 = _p(outs[2]) if self.coordim == 3 else None
        self.L.mfo_physderiv(self.h, nel, int(deformed), _p(df), _p(x), _p(outs[0]), o1, o2)
        return outs

    def helmholtz(self, nel, deformed, jac, df, lam, x):
        out = np.zeros(nel * self.nmTot)
        self.L.mfo_helmholtz(self.h, nel, int(deformed), _p(jac), _p(df), float(lam), _p(x), _p(out))
        return out

    def iproductwrtderivbase(self, nel, deformed, jac, df, ins):
        out = np.zeros(nel * self.nmTot)
        i1 = _p(ins[1]) if self.coordim >= 2 else None
        i2 = _p(ins[2]) if self.coordim == 3 else None
        rc = self.L.mfo_iproductwrtderivbase(self.h, nel, int(deformed), _p(jac), _p(df), _p(ins[0]), i1, i2, _p(out))
        if rc != 0:
            raise NotImplementedError
        return out

    def cg(self, nel, deformed, jac, df, lam, nglobal, ndir, l2g, sign, invdiag, rhs, tol=1e-9, maxiter=5000):
        x = np.zeros(nglobal)
        eps = C.c_double(0.0)
        l2g = np.ascontiguousarray(l2g, dtype=np.int32)
        its = self.L.mfo_cg_helmholtz(self.h, nel, int(deformed), _p(jac), _p(df), float(lam), l2g.size, nglobal,
                                      ndir, l2g.ctypes.data_as(_ip), _p(sign), _p(invdiag), _p(rhs), _p(x),
                                      float(tol), int(maxiter), C.byref(eps))
        return x, its, eps.value


def global_to_local(l2g, sign, glob):
    L = oracle_lib()
    l2g = np.ascontiguousarray(l2g, dtype=np.int32)
    loc = np.zeros(l2g.size)
    L.mfo_global_to_local(l2g.size, l2g.ctypes.data_as(_ip), _p(sign), _p(glob), _p(loc))
    return loc


def assemble(l2g, sign, loc, nglobal):
    L = oracle_lib()
    l2g = np.ascontiguousarray(l2g, dtype=np.int32)
    glob = np.zeros(nglobal)
    L.mfo_assemble(l2g.size, nglobal, l2g.ctypes.data_as(_ip), _p(sign), _p(loc), _p(glob))
    return glob


class Ref:
    """The reference's own kernels (oracle/_ref).  variant: 'scalar' (default build, width 1) or 'avx2'."""

    def __init__(self, variant="scalar"):
        path = os.path.join(HERE, "_ref", "libnekref_%s.so" % variant)
        if not os.path.exists(path) and os.path.isdir("/root/reference/library"):
            build("ref")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = C.CDLL(path)
        pp = C.POINTER(_dp)
        L.nekref_create.restype = C.c_void_p
        L.nekref_create.argtypes = [C.c_int] * 5 + [pp] * 5 + [_ip] * 3 + [C.c_int, _dp, _dp]
        L.nekref_create2.restype = C.c_void_p
        L.nekref_create2.argtypes = [C.c_int] * 5 + [pp] * 5 + [_ip] * 3 + [C.c_int, _dp, _dp, C.c_int]
        L.nekref_run.argtypes = [C.c_void_p] + [_dp] * 6 + [C.c_double, C.c_int]
        L.nekref_destroy.argtypes = [C.c_void_p]
        for f in ("nekref_zwglj", "nekref_zwgrjm"):
            getattr(L, f).argtypes = [_dp, _dp, C.c_int, C.c_double, C.c_double]
        for f in ("nekref_Dglj", "nekref_Dgrjm"):
            getattr(L, f).argtypes = [_dp, _dp, C.c_int, C.c_double, C.c_double]
        L.nekref_jacobfd.argtypes = [C.c_int, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double]
        self.L = L
        self.width = L.nekref_width()
        self.variant = variant

    def max_threads(self):
        return int(self.L.nekref_max_threads())

    def points(self, ptype, n):
        z, w, D = np.zeros(n), np.zeros(n), np.zeros(n * n)
        if ptype == 0:
            self.L.nekref_zwglj(_p(z), _p(w), n, 0.0, 0.0)
            self.L.nekref_Dglj(_p(D), _p(z), n, 0.0, 0.0)
        else:
            a = float(ptype)
            self.L.nekref_zwgrjm(_p(z), _p(w), n, a, 0.0)
            self.L.nekref_Dgrjm(_p(D), _p(z), n, a, 0.0)
        return z, w, D

    def jacobfd(self, z, n, a, b):
        p, pd = np.zeros(z.size), np.zeros(z.size)
        self.L.nekref_jacobfd(z.size, _p(z), _p(p), _p(pd), n, a, b)
        return p, pd

    def operator(self, op, el, nel, deformed, jac, df):
        return RefOperator(self, op, el, nel, deformed, jac, df)


class RefOperator:
    def __init__(self, ref, op, el, nel, deformed, jac, df):
        self.ref, self.op, self.el, self.nel = ref, op, el, nel
        dim = el.dim

        def arr(lst):
            a = (_dp * 3)()
            for d in range(dim):
                a[d] = _p(lst[d])
            return a

        self._keep = (el.bdata, el.dbdata, el.D, el.Z, el.W, jac, df)
        blen = (C.c_int * 3)(*[el.bdata[d].size for d in range(dim)] + [0] * (3 - dim))
        nqd = (C.c_int * 3)(*el.nq + [1] * (3 - dim))
        pty = (C.c_int * 3)(*el.ptype + [0] * (3 - dim))
        self.h = ref.L.nekref_create2(op, el.shape, el.nm, el.nq0, int(deformed), arr(el.bdata), arr(el.dbdata),
                                      arr(el.D), arr(el.Z), arr(el.W), blen, nqd, pty, nel, _p(jac), _p(df), el.coordim)
        if not self.h:
            raise ValueError("reference operator not available")

    def __del__(self):
        try:
            self.ref.L.nekref_destroy(self.h)
        except Exception:
            pass

    def __call__(self, ins, lam=0.0, nthreads=1, outs=None):
        el, nel = self.el, self.nel
        if not isinstance(ins, (list, tuple)):
            ins = [ins]
        nout = {OP_BWD: el.nqTot, OP_HELM: el.nmTot, OP_IPROD: el.nmTot, OP_IPWDB: el.nmTot,
                OP_PHYSDERIV: el.nqTot}[self.op]
        nouts = el.coordim if self.op == OP_PHYSDERIV else 1
        if outs is None:
            outs = [np.zeros(nel * nout) for _ in range(nouts)]
        i = list(ins) + [None] * (3 - len(ins))
        o = list(outs) + [None] * (3 - len(outs))
        rc = self.ref.L.nekref_run(self.h, _p(i[0]), _p(i[1]), _p(i[2]), _p(o[0]), _p(o[1]), _p(o[2]), float(lam),
                                   int(nthreads))
        if rc != 0:
            raise NotImplementedError("reference kernel not instantiated for nm=%d nq=%d (rc=%d)"
                                      % (el.nm, el.nq0, rc))
        return outs if nouts > 1 else outs[0]


class Chain:
    """ContField::v_HelmSolve -> GlobalLinSysIterativeFull::v_Solve -> DoConjugateGradient -> BwdTrans on the CPU
    (mf_oracle.c: mfo_chain_helmsolve).  engine "oracle": this package's plain-C operators (the checker);
    engine a Ref instance: the reference's own kernels from oracle/_ref (the CPU baseline), elements and vector loops
    over `threads` OpenMP threads."""

    def __init__(self, el, nel, deformed, jac, df, lam, l2g, sign, nglobal, ndir, invdiag, engine="oracle", threads=1):
        self.L = oracle_lib()
        self.el, self.nel, self.threads = el, nel, int(threads)
        self.l2g = np.ascontiguousarray(l2g, dtype=np.int32)
        self.sign = None if sign is None else np.ascontiguousarray(sign, dtype=np.float64)
        self.invdiag = None if invdiag is None else np.ascontiguousarray(invdiag, dtype=np.float64)
        self.nphys = nel * el.nqTot
        self.h = self.L.mfo_chain_create(self.l2g.size, int(nglobal), int(ndir), self.nphys, self.l2g.ctypes.data_as(_ip),
                                         _p(self.sign), _p(self.invdiag))
        self._keep = (jac, df)
        if engine == "oracle":
            L, h, de = self.L, el.h, int(deformed)
            self._cb = (ELOP(lambda ctx, i, o: L.mfo_iproduct(h, nel, de, _p(jac), i, o)),
                        ELOP(lambda ctx, i, o: L.mfo_helmholtz(h, nel, de, _p(jac), _p(df), float(lam), i, o)),
                        ELOP(lambda ctx, i, o: L.mfo_bwdtrans(h, nel, i, o)))
            self.kind = "port"
        else:
            ops = [engine.operator(o, el, nel, deformed, jac, df) for o in (OP_IPROD, OP_HELM, OP_BWD)]
            R, nt = engine.L, self.threads
            self._ops = ops
            self._cb = tuple(ELOP(lambda ctx, i, o, hh=op.h: R.nekref_run(hh, i, None, None, o, None, None, float(lam), nt))
                             for op in ops)
            self.kind = "reference"

    def helmsolve(self, forcing, inout, phys_out=None, tol=1e-9, maxiter=5000):
        """-> (iterations, final r.r); iterations < 0: the loop counter reached maxiter (the reference's efatal)"""
        set_threads(self.threads)
        eps = C.c_double(0.0)
        its = self.L.mfo_chain_helmsolve(self.h, self._cb[0], None, self._cb[1], None, self._cb[2], None, _p(forcing),
                                         _p(inout), _p(phys_out), float(tol), int(maxiter), C.byref(eps))
        return its, eps.value

    def __del__(self):
        try:
            self.L.mfo_chain_destroy(self.h)
        except Exception:
            pass
