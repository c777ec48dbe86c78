"""GPU suite: the CUDA path, called through the C ABI (ctypes -> libnekmf_b200.so), against the CPU
oracle on the same seeded inputs and against the reference-generated golden vectors.

Tolerance: relative L2 and Linf <= 1e-12 (BASELINE.json north_star) for every operator.
"""
import os

import numpy as np
import pytest

import pyoracle as po
from _util import ROOT, nekmf, random_geometry, rel_errs

pytestmark = pytest.mark.gpu
TOL = 1e-12
SHAPES = {"Quad": po.QUAD, "Tri": po.TRI, "Hex": po.HEX, "Prism": po.PRISM, "Pyr": po.PYR, "Tet": po.TET}
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))


def _torch():
    import torch
    return torch


def check(got, want, what=""):
    l2, li = rel_errs(got, want)
    assert l2 < TOL and li < TOL, "%s: relL2=%.3e relLinf=%.3e" % (what, l2, li)


def run_all_ops(nk, shape, nm, nq0, nel, deformed, rng, lam=1.3, geometry=None):
    el = po.Elem(shape, nm, nq0)
    std = nk.StdExpansion(shape, nm, nq0)
    jac, df = geometry if geometry is not None else random_geometry(rng, el.dim, nel, el.nqTot, deformed)
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, deformed))
    x = rng.uniform(-1, 1, nel * el.nmTot)
    f = [rng.uniform(-1, 1, nel * el.nqTot) for _ in range(el.dim)]
    out = np.zeros(nel * el.nqTot)
    coll.ApplyOperator(nk.eBwdTrans, x, out)
    check(out, el.bwdtrans(nel, x), "BwdTrans")
    out = np.zeros(nel * el.nmTot)
    coll.ApplyOperator(nk.eIProductWRTBase, f[0], out)
    check(out, el.iproduct(nel, deformed, jac, f[0]), "IProductWRTBase")
    outs = [np.zeros(nel * el.nqTot) for _ in range(el.dim)]
    coll.ApplyOperator(nk.ePhysDeriv, f[0], *outs)
    check(np.concatenate(outs), np.concatenate(el.physderiv(nel, deformed, df, f[0])), "PhysDeriv")
    out = np.zeros(nel * el.nmTot)
    coll.ApplyOperator(nk.eHelmholtz, x, out, factors={nk.eFactorLambda: lam})
    check(out, el.helmholtz(nel, deformed, jac, df, lam, x), "Helmholtz")
    out = np.zeros(nel * el.nmTot)
    coll.ApplyOperator(nk.eIProductWRTDerivBase, *f, out)
    check(out, el.iproductwrtderivbase(nel, deformed, jac, df, f), "IProductWRTDerivBase")
    return coll


@pytest.mark.parametrize("deformed", [False, True])
@pytest.mark.parametrize("nm", list(range(2, 12)))
def test_hex_all_operators_default_quadrature(nm, deformed):
    """nm=2..11 (P=1..10), nq=nm+1: the TMA-fed compile-time kernels; 37 elements = ragged last batch."""
    nk = nekmf()
    coll = run_all_ops(nk, po.HEX, nm, nm + 1, 37, deformed, np.random.default_rng(100 + nm))
    assert "hex_" in coll.m_ops[nk.eBwdTrans].kernel_name  # pencil (hex_op_kernel) or register-slab (hex_slab_kernel)


@pytest.mark.parametrize("deformed", [False, True])
@pytest.mark.parametrize("shape,nm,nq0", [
    (po.HEX, 4, 6), (po.HEX, 5, 8), (po.HEX, 5, 10), (po.HEX, 4, 4),
    (po.QUAD, 2, 3), (po.QUAD, 4, 5), (po.QUAD, 6, 7), (po.QUAD, 8, 9), (po.QUAD, 4, 8), (po.QUAD, 11, 12),
    (po.TRI, 2, 3), (po.TRI, 4, 5), (po.TRI, 6, 7), (po.TRI, 8, 9), (po.TRI, 4, 8),
    (po.PRISM, 2, 3), (po.PRISM, 4, 5), (po.PRISM, 7, 8), (po.PRISM, 8, 9), (po.PRISM, 4, 7),
    (po.TET, 2, 3), (po.TET, 4, 5), (po.TET, 7, 8), (po.TET, 8, 9), (po.TET, 4, 7),
    (po.PYR, 2, 3), (po.PYR, 3, 4), (po.PYR, 4, 5), (po.PYR, 5, 6), (po.PYR, 7, 8), (po.PYR, 8, 9), (po.PYR, 4, 7),
])
def test_all_shapes_runtime_kernels(shape, nm, nq0, deformed):
    """collapsed-coordinate shapes (CORRECT terms of eModified_A), quads, over-integration (nq up to 2 nm)."""
    run_all_ops(nekmf(), shape, nm, nq0, 13, deformed, np.random.default_rng(7 * nm + nq0 + shape))


# default policy of dense_helm.cu (dense_wanted): first nm at which regular Helmholtz takes the DMMA kernel
DENSE_FROM = {"Tri": 3, "Tet": 2, "Pyr": 2}


@pytest.mark.parametrize("nel", [1, 15, 16, 17, 63, 64, 65, 300])
@pytest.mark.parametrize("shape,nm", [("Tri", 2), ("Tri", 5), ("Tri", 7), ("Tri", 9),
                                      ("Tet", 2), ("Tet", 3), ("Tet", 4), ("Tet", 5), ("Tet", 6), ("Tet", 7), ("Tet", 8),
                                      ("Tet", 9), ("Pyr", 2), ("Pyr", 3), ("Pyr", 4), ("Pyr", 5), ("Pyr", 6), ("Pyr", 7)])
def test_dense_dmma_helmholtz(shape, nm, nel, monkeypatch):
    """Regular Tri / Tet / Pyr Helmholtz as a batched DMMA GEMM with the reference-element matrices (dense_helm.cu),
    forced on at every order it is instantiated for: ragged element tiles (16 per warp, 64 per CTA), row counts that
    are not a multiple of 8, lambda = 0, a changed lambda, a second set_geom, device arrays offset by one double."""
    monkeypatch.setenv("NEKMF_DENSE", "1")
    torch = _torch()
    nk = nekmf()
    rng = np.random.default_rng(nm * 131 + nel + len(shape))
    el = po.Elem(SHAPES[shape], nm, nm + 1)
    std = nk.StdExpansion(SHAPES[shape], nm, nm + 1)
    jac, df = random_geometry(rng, el.dim, nel, el.nqTot, False)
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, False))
    x = rng.uniform(-1, 1, nel * el.nmTot)
    for lam in (1.3, 0.0, 37.5):
        out = np.zeros(nel * el.nmTot)
        coll.ApplyOperator(nk.eHelmholtz, x, out, factors={nk.eFactorLambda: lam})
        check(out, el.helmholtz(nel, False, jac, df, lam, x), "Helmholtz(dense, lambda=%g)" % lam)
    assert "dense_helm_kernel" in coll.m_ops[nk.eHelmholtz].kernel_name, coll.m_ops[nk.eHelmholtz].kernel_name
    want = el.helmholtz(nel, False, jac, df, 1.0, x)
    xd = torch.zeros(x.size + 1, dtype=torch.float64, device="cuda")
    xd[1:] = torch.from_numpy(x).cuda()
    yd = torch.zeros(x.size + 1, dtype=torch.float64, device="cuda")
    coll.ApplyOperator(nk.eHelmholtz, xd[1:], yd[1:], factors={nk.eFactorLambda: 1.0})
    torch.cuda.synchronize()
    check(yd[1:].cpu().numpy(), want, "Helmholtz(dense, device arrays offset by one double)")
    # the quadrature-space kernel it replaces must agree to rounding (NEKMF_DENSE=0 is read at set_geom)
    monkeypatch.setenv("NEKMF_DENSE", "0")
    coll0 = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, False))
    out0 = np.zeros(nel * el.nmTot)
    coll0.ApplyOperator(nk.eHelmholtz, x, out0, factors={nk.eFactorLambda: 1.0})
    assert "dense_helm_kernel" not in coll0.m_ops[nk.eHelmholtz].kernel_name
    check(out0, want, "Helmholtz(quadrature space)")
    check(yd[1:].cpu().numpy(), out0, "dense vs quadrature-space kernel")


@pytest.mark.parametrize("nel", [1, 2, 3, 7, 8, 9, 100])
@pytest.mark.parametrize("nm", [2, 3, 4, 5, 6, 7, 8])
def test_prism_extruded_dmma_helmholtz(nm, nel, monkeypatch):
    """Regular prisms whose segment direction is orthogonal to the triangle plane (G01 = G12 = 0: extruded meshes,
    prisms cut from boxes) take prism_helm_kernel (dense_helm.cu): the segment's generalised eigen-decomposition
    turns the element into nm triangle Helmholtz problems for the DMMA GEMM.  The geometry is an extruded element
    rotated by a random orthogonal matrix, so all nine derivative factors are non-zero."""
    monkeypatch.delenv("NEKMF_DENSE", raising=False)
    nk = nekmf()
    rng = np.random.default_rng(nm * 53 + nel)
    el = po.Elem(po.PRISM, nm, nm + 1)
    std = nk.StdExpansion(po.PRISM, nm, nm + 1)
    jac = rng.uniform(0.5, 1.5, nel)
    df = np.zeros((3, 3, nel))  # [c][d]
    for e in range(nel):
        a = np.zeros((3, 3))
        a[np.ix_([0, 2], [0, 2])] = rng.uniform(-0.3, 0.3, (2, 2)) + 1.5 * np.eye(2)
        a[1, 1] = rng.uniform(0.8, 2.0)
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        df[:, :, e] = q @ a
    df = np.ascontiguousarray(df.reshape(9, nel)).reshape(-1)
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, False))
    x = rng.uniform(-1, 1, nel * el.nmTot)
    for lam in (1.3, 0.0, 37.5):
        out = np.zeros(nel * el.nmTot)
        coll.ApplyOperator(nk.eHelmholtz, x, out, factors={nk.eFactorLambda: lam})
        check(out, el.helmholtz(nel, False, jac, df, lam, x), "Helmholtz(prism extruded, lambda=%g)" % lam)
    assert "prism_helm_kernel" in coll.m_ops[nk.eHelmholtz].kernel_name, coll.m_ops[nk.eHelmholtz].kernel_name
    monkeypatch.setenv("NEKMF_DENSE", "0")
    coll0 = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, False))
    out0 = np.zeros(nel * el.nmTot)
    coll0.ApplyOperator(nk.eHelmholtz, x, out0, factors={nk.eFactorLambda: 37.5})
    assert "shape_op_kernel" in coll0.m_ops[nk.eHelmholtz].kernel_name
    check(out, out0, "prism DMMA vs quadrature-space kernel")


@pytest.mark.parametrize("nel", [1, 2, 3, 9, 100])
@pytest.mark.parametrize("nm", [2, 3, 4, 5, 6, 7, 8])
def test_prism_general_dmma_helmholtz(nm, nel, monkeypatch):
    """General regular prisms (G01, G12 != 0) take prism_gen_kernel (dense_helm.cu): eight triangle-matrix terms in the
    segment eigen-basis, mixed matrices recovered from probes (the numpy restatement of the same formulation is
    tests/test_oracle.py::test_prism_general_kronecker_formulation_against_oracle)."""
    monkeypatch.delenv("NEKMF_DENSE", raising=False)
    monkeypatch.setenv("NEKMF_PRISM_GENERAL", "1")  # every order (the default policy takes it at nm 3..5)
    nk = nekmf()
    rng = np.random.default_rng(nm * 71 + nel)
    el = po.Elem(po.PRISM, nm, nm + 1)
    std = nk.StdExpansion(po.PRISM, nm, nm + 1)
    jac, df = random_geometry(rng, 3, nel, el.nqTot, False)
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, False))
    x = rng.uniform(-1, 1, nel * el.nmTot)
    for lam in (1.3, 0.0, 37.5):
        out = np.zeros(nel * el.nmTot)
        coll.ApplyOperator(nk.eHelmholtz, x, out, factors={nk.eFactorLambda: lam})
        check(out, el.helmholtz(nel, False, jac, df, lam, x), "Helmholtz(prism general, lambda=%g)" % lam)
    assert "prism_gen_kernel" in coll.m_ops[nk.eHelmholtz].kernel_name, coll.m_ops[nk.eHelmholtz].kernel_name
    monkeypatch.setenv("NEKMF_PRISM_GENERAL", "0")
    coll0 = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, False))
    out0 = np.zeros(nel * el.nmTot)
    coll0.ApplyOperator(nk.eHelmholtz, x, out0, factors={nk.eFactorLambda: 37.5})
    assert "shape_op_kernel" in coll0.m_ops[nk.eHelmholtz].kernel_name
    check(out, out0, "prism general DMMA vs quadrature-space kernel")


def golden_cases():
    return sorted(set(k.rsplit("_", 1)[0] for k in GOLD.files if k.endswith("_x") and not k.startswith("Seg")))


@pytest.mark.parametrize("key", golden_cases())
def test_against_reference_golden_vectors(key):
    """outputs of the reference's own kernels (oracle/_ref at fixture-generation time)"""
    nk = nekmf()
    name, nm, nq0, geo = key.split("_")
    shape, nm, nq0, deformed = SHAPES[name], int(nm), int(nq0), geo == "def"
    g = lambda s: GOLD[key + "_" + s]
    std = nk.StdExpansion(shape, nm, nq0)
    ncoef, nq = std.GetNcoeffs(), std.GetTotPoints()
    nel = g("x").size // ncoef
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(g("jac"), g("df"), deformed))
    out = np.zeros(nel * nq)
    coll.ApplyOperator(nk.eBwdTrans, g("x"), out)
    check(out, g("bwd"), "BwdTrans")
    out = np.zeros(nel * ncoef)
    coll.ApplyOperator(nk.eIProductWRTBase, g("f0"), out)
    check(out, g("iprod"), "IProductWRTBase")
    outs = [np.zeros(nel * nq) for _ in range(std.dim)]
    coll.ApplyOperator(nk.ePhysDeriv, g("f0"), *outs)
    check(np.concatenate(outs), np.concatenate([g("pd%d" % d) for d in range(std.dim)]), "PhysDeriv")
    out = np.zeros(nel * ncoef)
    coll.ApplyOperator(nk.eHelmholtz, g("x"), out, factors={nk.eFactorLambda: 1.5})
    check(out, g("helm"), "Helmholtz")
    out = np.zeros(nel * ncoef)
    coll.ApplyOperator(nk.eIProductWRTDerivBase, *[g("f%d" % d) for d in range(std.dim)], out)
    check(out, g("ipwdb"), "IProductWRTDerivBase")


def box_geometry(nel, hx, hy, hz):
    jac = np.full(nel, hx * hy * hz / 8.0)
    df = np.zeros((9, nel))
    df[0], df[4], df[8] = 2.0 / hx, 2.0 / hy, 2.0 / hz
    return jac, df.reshape(-1).copy()


@pytest.mark.parametrize("nm", [2, 3, 4, 5, 6, 7, 8, 9, 10, 11])
@pytest.mark.parametrize("nel", [1, 31, 32, 33, 1000])
def test_hex_helmholtz_coefficient_space_kernel(nm, nel):
    """axis-aligned boxes (diagonal Laplacian metric) take the Kronecker coefficient-space kernel;
    per-element box sizes vary so the per-element scalars are exercised."""
    nk = nekmf()
    rng = np.random.default_rng(nm * 1000 + nel)
    el = po.Elem(po.HEX, nm, nm + 1)
    std = nk.StdExpansion(nk.eHexahedron, nm, nm + 1)
    h = rng.uniform(0.05, 2.0, (3, nel))
    jac = h[0] * h[1] * h[2] / 8.0
    df = np.zeros((9, nel))
    df[0], df[4], df[8] = 2.0 / h[0], 2.0 / h[1], 2.0 / h[2]
    df = df.reshape(-1).copy()
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, False))
    x = rng.uniform(-1, 1, nel * el.nmTot)
    for lam in (0.0, 1.0, 37.5):
        out = np.zeros(nel * el.nmTot)
        coll.ApplyOperator(nk.eHelmholtz, x, out, factors={nk.eFactorLambda: lam})
        check(out, el.helmholtz(nel, False, jac, df, lam, x), "Helmholtz(kron)")
    assert "kron" in coll.m_ops[nk.eHelmholtz].kernel_name
    # a sheared (non-diagonal metric) collection
    jac2, df2 = random_geometry(rng, 3, nel, el.nqTot, False)
    coll2 = nk.Collection(std, nel, nk.CoalescedGeomData(jac2, df2, False))
    out = np.zeros(nel * el.nmTot)
    coll2.ApplyOperator(nk.eHelmholtz, x, out, factors={nk.eFactorLambda: 1.0})
    check(out, el.helmholtz(nel, False, jac2, df2, 1.0, x), "Helmholtz(full metric)")
    # a sheared / rotated affine collection takes the full-metric coefficient-space kernels (hex_kron_full.cuh up to
    # nm = 6, hex_kron_fullrows.cuh at nm = 7..9; at nm = 10 the quadrature-space kernel measured faster)
    kn = coll2.m_ops[nk.eHelmholtz].kernel_name
    assert ("kronfull" in kn) == (nm <= 9) and ("kronfullrows" in kn) == (7 <= nm <= 9), kn
    for lam in (0.0, 37.5):
        coll2.ApplyOperator(nk.eHelmholtz, x, out, factors={nk.eFactorLambda: lam})
        check(out, el.helmholtz(nel, False, jac2, df2, lam, x), "Helmholtz(full metric, lambda %g)" % lam)


@pytest.mark.parametrize("nm", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("nel", [1, 31, 32, 33, 4097])
def test_quad_helmholtz_coefficient_space_kernel(nm, nel):
    """Axis-aligned regular quads (diagonal Laplacian metric) take quad_kron.cu: one lane per element, padded
    per-element TMA copies for even nm, one bulk copy per batch for odd nm, ragged last batch, device arrays that
    are only 8-byte aligned; rotated / sheared elements take the full-metric variant of the same kernel."""
    torch = _torch()
    nk = nekmf()
    rng = np.random.default_rng(nm * 977 + nel)
    el = po.Elem(po.QUAD, nm, nm + 1)
    std = nk.StdExpansion(nk.eQuadrilateral, nm, nm + 1)
    h = rng.uniform(0.05, 2.0, (2, nel))
    jac = h[0] * h[1] / 4.0
    df = np.zeros((4, nel))
    df[0], df[3] = 2.0 / h[0], 2.0 / h[1]
    df = df.reshape(-1).copy()
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, False))
    x = rng.uniform(-1, 1, nel * el.nmTot)
    for lam in (0.0, 1.0, 37.5):
        out = np.zeros(nel * el.nmTot)
        coll.ApplyOperator(nk.eHelmholtz, x, out, factors={nk.eFactorLambda: lam})
        check(out, el.helmholtz(nel, False, jac, df, lam, x), "Helmholtz(quad kron)")
    assert "quad_helm_kron" in coll.m_ops[nk.eHelmholtz].kernel_name
    # device-resident arrays: 16-byte aligned (TMA) and offset by one double (plain loads / stores)
    want = el.helmholtz(nel, False, jac, df, 1.0, x)
    for off in (0, 1):
        xd = torch.zeros(x.size + off, dtype=torch.float64, device="cuda")
        xd[off:] = torch.from_numpy(x).cuda()
        yd = torch.zeros(x.size + off, dtype=torch.float64, device="cuda")
        coll.ApplyOperator(nk.eHelmholtz, xd[off:], yd[off:], factors={nk.eFactorLambda: 1.0})
        torch.cuda.synchronize()
        check(yd[off:].cpu().numpy(), want, "Helmholtz(quad kron, device, offset %d)" % off)
    jac2, df2 = random_geometry(rng, 2, nel, el.nqTot, False)
    coll2 = nk.Collection(std, nel, nk.CoalescedGeomData(jac2, df2, False))
    out = np.zeros(nel * el.nmTot)
    coll2.ApplyOperator(nk.eHelmholtz, x, out, factors={nk.eFactorLambda: 1.0})
    check(out, el.helmholtz(nel, False, jac2, df2, 1.0, x), "Helmholtz(quad kron, full metric)")
    assert "full metric" in coll2.m_ops[nk.eHelmholtz].kernel_name
    coll2.ApplyOperator(nk.eHelmholtz, x, out, factors={nk.eFactorLambda: 0.0})
    check(out, el.helmholtz(nel, False, jac2, df2, 0.0, x), "Laplacian(quad kron, full metric)")


@pytest.mark.parametrize("nel", [0, 1, 2, 3, 5, 8, 9])
def test_edge_element_counts(nel):
    """empty, single-element and ragged collections (the reference pads to the SIMD width instead)."""
    nk = nekmf()
    rng = np.random.default_rng(nel)
    if nel == 0:
        std = nk.StdExpansion(nk.eHexahedron, 5, 6)
        coll = nk.Collection(std, 0, nk.CoalescedGeomData(np.zeros(0), np.zeros(0), False))
        coll.ApplyOperator(nk.eHelmholtz, np.zeros(1), np.zeros(1), factors={nk.eFactorLambda: 1.0})
        return
    for shape in (po.HEX, po.TET, po.QUAD):
        run_all_ops(nk, shape, 5, 6, nel, True, rng)
        run_all_ops(nk, shape, 5, 6, nel, False, rng)


def test_device_resident_and_misaligned_arrays():
    """torch CUDA tensors are used in place (no copies); pointers that are only 8-byte aligned (slices of
    a larger field, as ExpList passes `in + offset`) take the non-TMA load path and must agree."""
    torch = _torch()
    nk = nekmf()
    rng = np.random.default_rng(42)
    nel, nm = 203, 5
    el = po.Elem(po.HEX, nm, nm + 1)
    std = nk.StdExpansion(nk.eHexahedron, nm, nm + 1)
    for deformed in (False, True):
        jac, df = random_geometry(rng, 3, nel, el.nqTot, deformed)
        coll = nk.Collection(std, nel, nk.CoalescedGeomData(torch.tensor(jac, device="cuda"),
                                                            torch.tensor(df, device="cuda"), deformed))
        x = rng.uniform(-1, 1, nel * el.nmTot)
        want = el.helmholtz(nel, deformed, jac, df, 0.9, x)
        for off in (0, 1):
            xin = torch.zeros(x.size + 3, dtype=torch.float64, device="cuda")
            xin[off:off + x.size] = torch.tensor(x, device="cuda")
            out = torch.zeros(x.size + 3, dtype=torch.float64, device="cuda")
            coll.ApplyOperator(nk.eHelmholtz, xin[off:off + x.size], out[off:off + x.size],
                               factors={nk.eFactorLambda: 0.9})
            torch.cuda.synchronize()
            check(out[off:off + x.size].cpu().numpy(), want, "Helmholtz device off=%d" % off)
            assert float(out[off + x.size:].abs().max()) == 0.0 and (off == 0 or float(out[0]) == 0.0)
        # box geometry -> coefficient-space kernel, also with misaligned pointers
        if not deformed:
            jb, dfb = box_geometry(nel, 0.3, 0.2, 0.5)
            collb = nk.Collection(std, nel, nk.CoalescedGeomData(jb, dfb, False))
            wantb = el.helmholtz(nel, False, jb, dfb, 0.9, x)
            for off in (0, 1):
                xin = torch.zeros(x.size + 3, dtype=torch.float64, device="cuda")
                xin[off:off + x.size] = torch.tensor(x, device="cuda")
                out = torch.zeros(x.size + 3, dtype=torch.float64, device="cuda")
                collb.ApplyOperator(nk.eHelmholtz, xin[off:off + x.size], out[off:off + x.size],
                                    factors={nk.eFactorLambda: 0.9})
                torch.cuda.synchronize()
                check(out[off:off + x.size].cpu().numpy(), wantb, "Helmholtz kron off=%d" % off)


def test_physderiv_direction_overload_and_errors():
    nk = nekmf()
    rng = np.random.default_rng(9)
    nel = 6
    el = po.Elem(po.HEX, 4, 5)
    std = nk.StdExpansion(nk.eHexahedron, 4, 5)
    jac, df = random_geometry(rng, 3, nel, el.nqTot, True)
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, True))
    f = rng.uniform(-1, 1, nel * el.nqTot)
    want = el.physderiv(nel, True, df, f)
    for d in range(3):
        out = np.zeros(nel * el.nqTot)
        coll.ApplyOperator(nk.ePhysDeriv, d, f, out)
        check(out, want[d], "PhysDeriv dir %d" % d)
    # operator()(dir, ...) is invalid for every other operator (reference: NEKERROR efatal)
    with pytest.raises(nk.NekError):
        coll.ApplyOperator(nk.eBwdTrans, 0, f, np.zeros(nel * el.nqTot))
    # Helmholtz without eFactorLambda
    with pytest.raises(nk.NekError):
        coll.ApplyOperator(nk.eHelmholtz, np.zeros(nel * el.nmTot), np.zeros(nel * el.nmTot))
    # geometry missing
    coll2 = nk.Collection(std, nel, nk.CoalescedGeomData(None, None, True))
    with pytest.raises(nk.NekError):
        coll2.ApplyOperator(nk.eIProductWRTBase, f, np.zeros(nel * el.nmTot))
    # unregistered implementation type
    coll3 = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, True), nk.SetFixedImpType(nk.eStdMat))
    with pytest.raises(nk.NekError):
        coll3.Initialise(nk.eBwdTrans)


def test_full_size_properties_hex_p4():
    """BASELINE.json size (64^3 hex, P=4): size-independent properties instead of a CPU recompute:
    linearity, symmetry x^T A y = y^T A x, positivity, BwdTrans/IProduct adjointness, and agreement of
    the coefficient-space kernel with the quadrature-space kernel on the same box mesh."""
    torch = _torch()
    nk = nekmf()
    nel, nm, nq = 64 ** 3, 5, 6
    std = nk.StdExpansion(nk.eHexahedron, nm, nq)
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.rand(nel * 125, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    y = torch.rand(nel * 125, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    h = 1.0 / 64
    jac = torch.full((nel,), (h / 2) ** 3, dtype=torch.float64, device="cuda")
    df = torch.zeros((9, nel), dtype=torch.float64, device="cuda")
    df[0] = df[4] = df[8] = 2 / h
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df.reshape(-1), False))
    # the same collection through the quadrature-space kernel (coefficient-space kernels switched off at creation)
    os.environ["NEKMF_HEX_KRON"] = "0"
    try:
        coll_q = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df.reshape(-1), False))
        coll_q.Initialise(nk.eHelmholtz)
    finally:
        del os.environ["NEKMF_HEX_KRON"]
    # ... and, with a tiny shear on the LAST element only, through the full-metric coefficient-space kernel
    df2 = df.clone()
    df2[1, -1] = 1e-300
    coll_f = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df2.reshape(-1), False))
    lam = {nk.eFactorLambda: 1.0}
    Ax, Ay, Axy, Aq = (torch.empty_like(x) for _ in range(4))
    coll.ApplyOperator(nk.eHelmholtz, x, Ax, factors=lam)
    coll.ApplyOperator(nk.eHelmholtz, y, Ay, factors=lam)
    coll.ApplyOperator(nk.eHelmholtz, (2.0 * x - 3.0 * y).contiguous(), Axy, factors=lam)
    coll_q.ApplyOperator(nk.eHelmholtz, x, Aq, factors=lam)
    Af = torch.empty_like(x)
    coll_f.ApplyOperator(nk.eHelmholtz, x, Af, factors=lam)
    torch.cuda.synchronize()
    assert "hex_helm_kron_kernel" in coll.m_ops[nk.eHelmholtz].kernel_name
    assert "kron" not in coll_q.m_ops[nk.eHelmholtz].kernel_name
    assert "kronfull" in coll_f.m_ops[nk.eHelmholtz].kernel_name
    assert float(torch.linalg.vector_norm(Af - Aq)) < 1e-12 * float(torch.linalg.vector_norm(Aq))
    nrm = float(torch.linalg.vector_norm(Ax))
    assert float(torch.linalg.vector_norm(Axy - (2.0 * Ax - 3.0 * Ay))) < 1e-12 * nrm * 5
    assert float(torch.linalg.vector_norm(Ax - Aq)) < 1e-12 * nrm
    assert float((Ax - Aq).abs().max()) < 1e-12 * float(Aq.abs().max())
    xAy, yAx = float(torch.dot(x, Ay)), float(torch.dot(y, Ax))
    assert abs(xAy - yAx) < 1e-12 * max(abs(xAy), float(torch.dot(x, Ax)))
    assert float(torch.dot(x, Ax)) > 0
    # <B u, f>_{Jw} = <u, IProduct(f)>
    fphys = torch.rand(nel * 216, dtype=torch.float64, device="cuda", generator=g)
    Bu, Itf = torch.empty_like(fphys), torch.empty_like(x)
    coll.ApplyOperator(nk.eBwdTrans, x, Bu)
    coll.ApplyOperator(nk.eIProductWRTBase, fphys, Itf)
    torch.cuda.synchronize()
    z, w, _ = nk.points(nk.eGaussLobattoLegendre, nq)
    wt = torch.tensor(w, device="cuda")
    w3 = (wt[:, None, None] * wt[None, :, None] * wt[None, None, :]).reshape(-1) * (h / 2) ** 3
    lhs = float(torch.sum(Bu.reshape(nel, 216) * fphys.reshape(nel, 216) * w3[None, :]))
    rhs = float(torch.dot(x, Itf))
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)


def test_cpp_collections_mirror():
    """the adapter source a maintainer drops into library/Collections (integration/B200Operators.cpp: the five
    *_B200 operator classes over the C ABI, registered in the operator factory) compiled against the stand-in headers
    of tests/cpp/ and driven by tests/cpp/TestCollectionB200.cpp, the analogue of
    library/UnitTests/Collections/TestHexCollection.cpp"""
    import subprocess
    exe = os.path.join(ROOT, "tests", "cpp", "TestCollectionB200")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASSED" in r.stdout


@pytest.mark.parametrize("case", ["hex_regular_diag", "hex_regular_sheared", "hex_deformed", "tet_deformed", "quad_regular",
                                  "tet_regular", "prism_extruded"])
def test_host_array_pipeline_many_chunks(case):
    """NEKMF_HOST applies are cut into element chunks (2 MB ramping to 32 MB) over a 3-stream H2D/kernel/D2H pipeline
    (abi.cu): collections large enough for several chunks (ragged last one) must give exactly the
    device-array result for every operator (incl. 3-input / 3-output ones), and match the oracle."""
    torch = _torch()
    nk = nekmf()
    rng = np.random.default_rng(hash(case) % 1000)
    shape, nm, nq0, nel, deformed = {
        "hex_regular_diag": (po.HEX, 5, 6, 20001, False), "hex_regular_sheared": (po.HEX, 5, 6, 20001, False),
        "hex_deformed": (po.HEX, 4, 5, 25003, True), "tet_deformed": (po.TET, 5, 6, 40001, True),
        "quad_regular": (po.QUAD, 6, 7, 90001, False), "tet_regular": (po.TET, 6, 7, 30001, False),
        "prism_extruded": (po.PRISM, 5, 6, 30001, False)}[case]
    el = po.Elem(shape, nm, nq0)
    jac, df = random_geometry(rng, el.dim, nel, el.nqTot, deformed)
    if case == "hex_regular_diag":
        df = df.reshape(9, nel).copy()
        for n in range(9):
            if n not in (0, 4, 8):
                df[n] = 0.0
        df = df.reshape(-1)
    if case == "prism_extruded":  # segment direction orthogonal to the triangle plane: G01 = G12 = 0
        df = df.reshape(9, nel).copy()
        for n in (1, 3, 5, 7):
            df[n] = 0.0
        df = df.reshape(-1)
    std = nk.StdExpansion(shape, nm, nq0)
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, deformed))
    x = rng.uniform(-1, 1, nel * el.nmTot)
    f = [rng.uniform(-1, 1, nel * el.nqTot) for _ in range(el.dim)]
    dev = torch.device("cuda", 0)
    tx = torch.from_numpy(x).to(dev)
    tf = [torch.from_numpy(a).to(dev) for a in f]

    def both(op, ins_h, ins_d, nout, out_len, **kw):
        oh = [np.zeros(out_len) for _ in range(nout)]
        od = [torch.zeros(out_len, dtype=torch.float64, device=dev) for _ in range(nout)]
        if op == nk.eIProductWRTDerivBase:
            coll.ApplyOperator(op, *ins_h, oh[0])
            coll.ApplyOperator(op, *ins_d, od[0])
        else:
            coll.ApplyOperator(op, ins_h[0], *oh, **kw)
            coll.ApplyOperator(op, ins_d[0], *od, **kw)
        torch.cuda.synchronize()
        for a, b in zip(oh, od):
            assert np.array_equal(a, b.cpu().numpy()), "host-array and device-array results differ (op %d)" % op
        return oh

    both(nk.eBwdTrans, [x], [tx], 1, nel * el.nqTot)
    both(nk.eIProductWRTBase, [f[0]], [tf[0]], 1, nel * el.nmTot)
    both(nk.ePhysDeriv, [f[0]], [tf[0]], el.dim, nel * el.nqTot)
    if case == "hex_regular_diag":
        assert "kron" in coll.m_ops[nk.eHelmholtz].kernel_name if nk.eHelmholtz in coll.m_ops else True
    h = both(nk.eHelmholtz, [x], [tx], 1, nel * el.nmTot, factors={nk.eFactorLambda: 0.7})[0]
    if case == "tet_regular":
        assert "dense_helm_kernel" in coll.m_ops[nk.eHelmholtz].kernel_name
    if case == "prism_extruded":
        assert "prism_helm_kernel" in coll.m_ops[nk.eHelmholtz].kernel_name
    both(nk.eIProductWRTDerivBase, f, tf, 1, nel * el.nmTot)
    # oracle on the last 300 elements (covers the ragged last chunk and its geometry offsets)
    ns = 300
    e0 = nel - ns
    gs = el.nqTot if deformed else 1
    jac_s = jac[e0 * gs:]
    df_s = np.ascontiguousarray(df.reshape(el.dim * el.dim, -1)[:, e0 * gs:]).reshape(-1)
    check(h[e0 * el.nmTot:], el.helmholtz(ns, deformed, jac_s, df_s, 0.7, x[e0 * el.nmTot:]), "Helmholtz tail")


@pytest.mark.parametrize("deformed", [False, True])
@pytest.mark.parametrize("nm", list(range(2, 10)))
@pytest.mark.parametrize("shape", ["Quad", "Tri", "Prism", "Tet"])
def test_shape_fast_kernels(shape, nm, deformed):
    """compile-time sized Quad/Tri/Prism/Tet kernels (shape_kernels.cuh), nm=2..9 with the default quadrature:
    several batches per CTA plus a ragged last batch; singular vertex/edge (CORRECT) terms folded into the
    sum-factorisation intermediates must reproduce the reference's separate correction loops."""
    nk = nekmf()
    nel = {"Quad": 331, "Tri": 331, "Prism": 67, "Tet": 67}[shape]
    coll = run_all_ops(nk, SHAPES[shape], nm, nm + 1, nel, deformed, np.random.default_rng(31 * nm + len(shape)))
    for op in (nk.eBwdTrans, nk.eIProductWRTBase, nk.ePhysDeriv, nk.eHelmholtz):
        # regular quads up to nm = 8 take the coefficient-space Helmholtz kernel (quad_kron.cu, full-metric variant
        # for this random geometry)
        want = "shape_op_kernel"
        if shape == "Quad" and nm <= 8:
            if op == nk.eHelmholtz:
                want = "quad_helm_kron" if not deformed else "shape_op_kernel"
            elif nm <= 7 and not (op == nk.ePhysDeriv and deformed):
                want = "quad_lane_kernel"  # one lane per element (quad_lane.cu)
        if shape == "Tri" and nm <= 7 and op != nk.eHelmholtz and not (op == nk.ePhysDeriv and deformed):
            want = "tri_lane_kernel"  # one lane per element (tri_lane.cu)
        if op == nk.eHelmholtz and not deformed and ((shape == "Tet" and nm >= DENSE_FROM["Tet"]) or
                                                      (shape == "Tri" and nm >= DENSE_FROM["Tri"])):
            want = "dense_helm_kernel"  # DMMA coefficient-space kernel (dense_helm.cu)
        if op == nk.eHelmholtz and not deformed and shape == "Prism" and 3 <= nm <= 5:
            want = "prism_gen_kernel"  # general regular prisms: eight-term DMMA kernel (dense_helm.cu), default at nm 3..5
        if op == nk.eHelmholtz and not deformed and shape == "Prism" and nm in (6, 7):
            want = "prism_helm_dmma_kernel"  # general regular prisms: fused quadrature-space kernel on tensor tiles
        if op in (nk.eBwdTrans, nk.eIProductWRTBase):
            # tensor-core kernels where they measured faster (prism_dmma.cu, tet_dmma.cu)
            if shape == "Prism" and 5 <= nm <= 7:
                want = "prism_dmma_kernel"
            if shape == "Tet" and nm in (5, 6, 7):
                want = "tet_bwd_gemm_kernel" if op == nk.eBwdTrans else ("tet_dmma_kernel" if (nm == 6 and deformed) else "tet_ip_gemm_kernel")
        assert want in coll.m_ops[op].kernel_name, coll.m_ops[op].kernel_name
    # IProductWRTDerivBase: lane kernels for regular quads up to nm = 5 / triangles up to nm = 6, otherwise the compile-time
    # sized kernel (chain-rule stage + the transposed-derivative / IProduct half of the fused Helmholtz kernel)
    lane = not deformed and ((shape == "Quad" and nm <= 5) or (shape == "Tri" and nm <= 6))
    name = coll.m_ops[nk.eIProductWRTDerivBase].kernel_name
    assert ("_lane_kernel<ipwdb" if lane else "shape_op_kernel") in name, name


@pytest.mark.parametrize("zero_copy", ["0", "1"])
def test_host_array_pinned_and_pageable_repeated(zero_copy, monkeypatch):
    """Repeated NEKMF_HOST applies.  NEKMF_HOST_ZEROCOPY=1: page-locked arrays are read and written by the
    kernels directly over PCIe (abi.cu); otherwise, and always for pageable arrays, the staged copy pipeline
    runs.  New input VALUES, a new lambda and alternating array kinds must all be honoured and both routes
    must agree bit for bit."""
    monkeypatch.setenv("NEKMF_HOST_ZEROCOPY", zero_copy)
    torch = _torch()
    nk = nekmf()
    rng = np.random.default_rng(5)
    nm, nel = 5, 30011
    el = po.Elem(po.HEX, nm, nm + 1)
    jac, df = box_geometry(nel, 0.1, 0.2, 0.3)
    std = nk.StdExpansion(po.HEX, nm, nm + 1)
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, False))
    xh = torch.empty(nel * el.nmTot, dtype=torch.float64).pin_memory()
    yh = torch.empty(nel * el.nmTot, dtype=torch.float64).pin_memory()
    ns = 200  # oracle on the last elements (ragged tail of the chunk schedule)
    js, dfs = jac[-ns:], np.ascontiguousarray(df.reshape(9, -1)[:, -ns:]).reshape(-1)
    for it, lam in enumerate([1.0, 1.0, 1.0, 2.5, 2.5, 2.5]):
        x = rng.uniform(-1, 1, nel * el.nmTot)
        xh.copy_(torch.from_numpy(x))
        yh.zero_()
        coll.ApplyOperator(nk.eHelmholtz, xh, yh, factors={nk.eFactorLambda: lam})
        check(yh.numpy()[-ns * el.nmTot:], el.helmholtz(ns, False, js, dfs, lam, x[-ns * el.nmTot:]), "replay %d" % it)
        check(yh.numpy()[:ns * el.nmTot], el.helmholtz(ns, False, jac[:ns], np.ascontiguousarray(
            df.reshape(9, -1)[:, :ns]).reshape(-1), lam, x[:ns * el.nmTot]), "replay head %d" % it)
    # pageable numpy arrays in between, then the pinned pair again
    x = rng.uniform(-1, 1, nel * el.nmTot)
    y = np.zeros(nel * el.nmTot)
    for _ in range(3):
        coll.ApplyOperator(nk.eHelmholtz, x, y, factors={nk.eFactorLambda: 2.5})
    check(y[-ns * el.nmTot:], el.helmholtz(ns, False, js, dfs, 2.5, x[-ns * el.nmTot:]), "pageable")
    xh.copy_(torch.from_numpy(x))
    for _ in range(3):
        yh.zero_()
        coll.ApplyOperator(nk.eHelmholtz, xh, yh, factors={nk.eFactorLambda: 2.5})
        assert np.array_equal(yh.numpy(), y)


@pytest.mark.parametrize("deformed", [False, True])
@pytest.mark.parametrize("coordim", [1, 2, 3])
@pytest.mark.parametrize("nm,nq0", [(2, 3), (4, 5), (7, 8), (11, 12), (5, 8)])
def test_segment_operators(nm, nq0, coordim, deformed):
    """1-D elements embedded in 1..3 space dimensions (boundary / trace expansions): every operator the reference
    registers for eSegment, against the oracle and (where stored) the reference's golden vectors."""
    nk = nekmf()
    rng = np.random.default_rng(100 * nm + 10 * coordim + deformed)
    nel = 1001
    el = po.Elem(po.SEG, nm, nq0, coordim=coordim)
    std = nk.StdExpansion(nk.eSegment, nm, nq0, coordim=coordim)
    assert std.GetNcoeffs() == nm and std.GetTotPoints() == nq0
    npt = nel * (nq0 if deformed else 1)
    jac = rng.uniform(0.5, 1.5, npt)
    df = rng.uniform(-1.5, 1.5, coordim * npt)
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, deformed))
    x = rng.uniform(-1, 1, nel * nm)
    f = [rng.uniform(-1, 1, nel * nq0) for _ in range(coordim)]
    out = np.zeros(nel * nq0)
    coll.ApplyOperator(nk.eBwdTrans, x, out)
    check(out, el.bwdtrans(nel, x), "BwdTransSeg")
    out = np.zeros(nel * nm)
    coll.ApplyOperator(nk.eIProductWRTBase, f[0], out)
    check(out, el.iproduct(nel, deformed, jac, f[0]), "IProductSeg")
    outs = [np.zeros(nel * nq0) for _ in range(coordim)]
    coll.ApplyOperator(nk.ePhysDeriv, f[0], *outs)
    check(np.concatenate(outs), np.concatenate(el.physderiv(nel, deformed, df, f[0])), "PhysDerivSeg")
    if coordim == 3 and not deformed:
        # the reference multiplies the third input by df[1] here (IProductWRTDerivBase.h:321-323): refused
        with pytest.raises(nk.NekError, match="not supported"):
            coll.Initialise(nk.eIProductWRTDerivBase)
    else:
        out = np.zeros(nel * nm)
        coll.ApplyOperator(nk.eIProductWRTDerivBase, *f, out)
        check(out, el.iproductwrtderivbase(nel, deformed, jac, df, f), "IProductWRTDerivBaseSeg")
    with pytest.raises(nk.NekError):
        coll.Initialise(nk.eHelmholtz)  # no (eSegment, eHelmholtz) operator in the reference either


@pytest.mark.parametrize("geometry", ["box_cut", "general_affine", "deformed"])
def test_config4_mixed_mesh_against_oracle(geometry):
    """BASELINE configs[3] on the ACTUAL synthetic mixed mesh of bench_configs.py (3^3 cubes: 18 hexahedra, 14 prisms,
    12 tetrahedra, P=6): box-cut (axis-aligned affine elements, extruded prisms), under a global rotation + shear
    (general affine: full metric, non-extruded prisms) and warped (per-point factors); every collection against the
    oracle, all five operators' worth of geometry exercised through Helmholtz + IProductWRTBase + PhysDeriv"""
    import bench_configs as bc
    torch = _torch()
    nk = nekmf()
    nm, n, lam = 7, 3, 1.0
    shp = {"Hex": (nk.eHexahedron, po.HEX), "Prism": (nk.ePrism, po.PRISM), "Tet": (nk.eTetrahedron, po.TET)}
    std_of = {k: nk.StdExpansion(v[0], nm) for k, v in shp.items()}
    deformed = geometry == "deformed"
    if geometry == "box_cut":
        mesh = bc.mixed_mesh(n, 1.0 / n)
    elif geometry == "general_affine":
        mesh = bc.mixed_mesh(n, 1.0 / n, bc.GENERAL_MAP)
    else:
        mesh = {k: (v[0], v[1].cpu().numpy(), v[2].cpu().numpy()) for k, v in
                bc.mixed_mesh_deformed(n, 1.0 / n, std_of, torch, torch.device("cuda"), bc.GENERAL_MAP).items()}
    rng = np.random.default_rng(5)
    kernels = set()
    for name, (nshape, pshape) in shp.items():
        nel, jac, df = mesh[name]
        el = po.Elem(pshape, nm, nm + 1)
        geom = nk.CoalescedGeomData(jac, df, deformed)
        x = rng.uniform(-1, 1, nel * el.nmTot)
        f = rng.uniform(-1, 1, nel * el.nqTot)
        coll = nk.Collection(std_of[name], nel, geom)
        out = np.zeros(nel * el.nmTot)
        coll.ApplyOperator(nk.eHelmholtz, x, out, factors={nk.eFactorLambda: lam})
        assert max(rel_errs(out, el.helmholtz(nel, deformed, jac, df, lam, x))) < 1e-12, (name, geometry)
        kernels.add(coll.m_ops[nk.eHelmholtz].kernel_name)
        out2 = np.zeros(nel * el.nmTot)
        coll.ApplyOperator(nk.eIProductWRTBase, f, out2)
        assert max(rel_errs(out2, el.iproduct(nel, deformed, jac, f))) < 1e-12
        d = [np.zeros(nel * el.nqTot) for _ in range(3)]
        coll.ApplyOperator(nk.ePhysDeriv, f, *d)
        for g, w in zip(d, el.physderiv(nel, deformed, df, f)):
            assert max(rel_errs(g, w)) < 1e-12
    assert len(kernels) == 3


@pytest.mark.parametrize("nm,nel", [(7, 1), (7, 2), (7, 37), (7, 4096 + 3), (8, 37), (9, 1), (9, 38), (10, 37), (11, 2),
                                    (11, 1001)])
@pytest.mark.parametrize("deformed", [False, True])
def test_hex_dmma_bwd_iprod(nm, nel, deformed, monkeypatch):
    """the tensor-core (DMMA m8n8k4, two contractions chained in registers) BwdTrans / IProductWRTBase at nm = 7..11,
    nq = nm + 1 (hex_dmma.cu): against the oracle and against the DFMA kernel it replaces (NEKMF_HEX_DMMA=0), odd
    element counts (the single-element tail), caller arrays that are only 8-byte aligned"""
    torch = _torch()
    nk = nekmf()
    rng = np.random.default_rng(nel)
    nq = nm + 1
    el = po.Elem(po.HEX, nm, nq)
    std = nk.StdExpansion(nk.eHexahedron, nm, nq)
    jac, df = random_geometry(rng, 3, nel, el.nqTot, deformed)
    geom = nk.CoalescedGeomData(jac, df, deformed)
    c = rng.uniform(-1, 1, nel * el.nmTot)
    f = rng.uniform(-1, 1, nel * el.nqTot)
    want_b, want_i = el.bwdtrans(nel, c), el.iproduct(nel, deformed, jac, f)
    res = {}
    for mode in ("all", "0"):
        monkeypatch.setenv("NEKMF_HEX_DMMA", mode)
        bwd, ipr = nk.Operator(std, nel, geom, nk.eBwdTrans), nk.Operator(std, nel, geom, nk.eIProductWRTBase)
        assert ("hex_dmma_kernel" in bwd.kernel_name) == (mode == "all"), bwd.kernel_name
        assert ("hex_dmma_kernel" in ipr.kernel_name) == (mode == "all"), ipr.kernel_name
        ob, oi = np.zeros(nel * el.nqTot), np.zeros(nel * el.nmTot)
        bwd.apply([c], [ob])
        ipr.apply([f], [oi])
        assert max(rel_errs(ob, want_b)) < 1e-12 and max(rel_errs(oi, want_i)) < 1e-12
        # device arrays at an odd offset (8-byte aligned only)
        cd = torch.zeros(c.size + 1, dtype=torch.float64, device="cuda")
        cd[1:] = torch.tensor(c, device="cuda")
        od = torch.zeros(ob.size + 1, dtype=torch.float64, device="cuda")
        bwd.apply([cd[1:]], [od[1:]])
        fd = torch.zeros(f.size + 1, dtype=torch.float64, device="cuda")
        fd[1:] = torch.tensor(f, device="cuda")
        oid = torch.zeros(oi.size + 1, dtype=torch.float64, device="cuda")
        ipr.apply([fd[1:]], [oid[1:]])
        torch.cuda.synchronize()
        assert np.array_equal(od[1:].cpu().numpy(), ob) and np.array_equal(oid[1:].cpu().numpy(), oi)
        res[mode] = (ob, oi)
    assert max(rel_errs(res["all"][0], res["0"][0])) < 1e-13 and max(rel_errs(res["all"][1], res["0"][1])) < 1e-13


@pytest.mark.parametrize("nel", [1, 2, 37, 4096 + 3])
def test_hex_dmma_physderiv_nm7(nel, monkeypatch):
    """regular PhysDeriv at nm = 7, nq = 8 on FP64 tensor-core tiles (hex_dmma.cu: xi_0 with the lane's own values as the
    A operand, xi_1 from fragment loads, xi_2 as DFMA in the owning lane): against the oracle and against the pencil kernel
    it replaces (NEKMF_HEX_DMMA=0), odd element counts, caller arrays that are only 8-byte aligned"""
    torch = _torch()
    nk = nekmf()
    rng = np.random.default_rng(100 + nel)
    nm, nq = 7, 8
    el = po.Elem(po.HEX, nm, nq)
    std = nk.StdExpansion(nk.eHexahedron, nm, nq)
    jac, df = random_geometry(rng, 3, nel, el.nqTot, False)
    geom = nk.CoalescedGeomData(jac, df, False)
    f = rng.uniform(-1, 1, nel * el.nqTot)
    want = el.physderiv(nel, False, df, f)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("NEKMF_HEX_DMMA", mode)
        pd = nk.Operator(std, nel, geom, nk.ePhysDeriv)
        assert ("hex_dmma_pd_kernel" in pd.kernel_name) == (mode == "1"), pd.kernel_name
        outs = [np.zeros(nel * el.nqTot) for _ in range(3)]
        pd.apply([f], outs)
        for g, w in zip(outs, want):
            assert max(rel_errs(g, w)) < 1e-12
        fd = torch.zeros(f.size + 1, dtype=torch.float64, device="cuda")
        fd[1:] = torch.tensor(f, device="cuda")
        od = [torch.zeros(f.size + 1, dtype=torch.float64, device="cuda") for _ in range(3)]
        pd.apply([fd[1:]], [o[1:] for o in od])
        torch.cuda.synchronize()
        for o, g in zip(od, outs):
            assert np.array_equal(o[1:].cpu().numpy(), g)
        res[mode] = outs
    for a, b in zip(res["1"], res["0"]):
        assert max(rel_errs(a, b)) < 1e-13


@pytest.mark.parametrize("nm,nel", [(3, 5), (4, 1), (4, 38), (5, 37), (6, 2), (6, 1001), (7, 1), (7, 2), (7, 37), (7, 4099)])
@pytest.mark.parametrize("deformed", [False, True])
def test_prism_dmma_bwd_iprod(nm, nel, deformed, monkeypatch):
    """BwdTrans / IProductWRTBase on prisms at nm = 3..7 on FP64 tensor-core tiles (prism_dmma.cu: two chained DMMA
    passes per quadrature plane, the collapsed xi_2 contraction and the singular-edge correction in the owning lane):
    against the oracle and against the pencil kernels they replace (NEKMF_PRISM_DMMA=0), odd element counts, caller
    arrays that are only 8-byte aligned"""
    torch = _torch()
    nk = nekmf()
    rng = np.random.default_rng(7 * nm + nel)
    el = po.Elem(po.PRISM, nm, nm + 1)
    std = nk.StdExpansion(nk.ePrism, nm)
    jac, df = random_geometry(rng, 3, nel, el.nqTot, deformed)
    geom = nk.CoalescedGeomData(jac, df, deformed)
    c = rng.uniform(-1, 1, nel * el.nmTot)
    f = rng.uniform(-1, 1, nel * el.nqTot)
    want_b, want_i = el.bwdtrans(nel, c), el.iproduct(nel, deformed, jac, f)
    res = {}
    for mode in ("all", "0"):
        monkeypatch.setenv("NEKMF_PRISM_DMMA", mode)
        bwd, ipr = nk.Operator(std, nel, geom, nk.eBwdTrans), nk.Operator(std, nel, geom, nk.eIProductWRTBase)
        assert ("prism_dmma_kernel" in bwd.kernel_name) == (mode == "all"), bwd.kernel_name
        assert ("prism_dmma_kernel" in ipr.kernel_name) == (mode == "all"), ipr.kernel_name
        ob, oi = np.zeros(nel * el.nqTot), np.zeros(nel * el.nmTot)
        bwd.apply([c], [ob])
        ipr.apply([f], [oi])
        assert max(rel_errs(ob, want_b)) < 1e-12 and max(rel_errs(oi, want_i)) < 1e-12
        # device arrays at an odd offset (8-byte aligned only)
        cd = torch.zeros(c.size + 1, dtype=torch.float64, device="cuda")
        cd[1:] = torch.tensor(c, device="cuda")
        od = torch.zeros(ob.size + 1, dtype=torch.float64, device="cuda")
        bwd.apply([cd[1:]], [od[1:]])
        fd = torch.zeros(f.size + 1, dtype=torch.float64, device="cuda")
        fd[1:] = torch.tensor(f, device="cuda")
        oid = torch.zeros(oi.size + 1, dtype=torch.float64, device="cuda")
        ipr.apply([fd[1:]], [oid[1:]])
        torch.cuda.synchronize()
        assert np.array_equal(od[1:].cpu().numpy(), ob) and np.array_equal(oid[1:].cpu().numpy(), oi)
        res[mode] = (ob, oi)
    assert max(rel_errs(res["all"][0], res["0"][0])) < 1e-13 and max(rel_errs(res["all"][1], res["0"][1])) < 1e-13


@pytest.mark.parametrize("nm,nel", [(3, 5), (4, 1), (4, 38), (5, 37), (5, 2), (6, 2), (6, 1001), (7, 1), (7, 2), (7, 37), (7, 4099)])
@pytest.mark.parametrize("deformed", [False, True])
def test_tet_dmma_bwd_iprod(nm, nel, deformed, monkeypatch):
    """BwdTrans / IProductWRTBase on tetrahedra at nm = 3..7 (tet_dmma.cu: the xi_0 contraction on FP64 tensor-core tiles,
    the two collapsed contractions by one lane per mode pair, top-vertex / bottom-vertex / singular-edge corrections by
    two extra lanes): against the oracle and against the pencil kernels they replace (NEKMF_TET_DMMA=0), odd element
    counts, caller arrays that are only 8-byte aligned"""
    torch = _torch()
    nk = nekmf()
    rng = np.random.default_rng(11 * nm + nel)
    el = po.Elem(po.TET, nm, nm + 1)
    std = nk.StdExpansion(nk.eTetrahedron, nm)
    jac, df = random_geometry(rng, 3, nel, el.nqTot, deformed)
    geom = nk.CoalescedGeomData(jac, df, deformed)
    c = rng.uniform(-1, 1, nel * el.nmTot)
    f = rng.uniform(-1, 1, nel * el.nqTot)
    want_b, want_i = el.bwdtrans(nel, c), el.iproduct(nel, deformed, jac, f)
    res = {}
    monkeypatch.setenv("NEKMF_TET_GEMM", "0")  # the GEMM kernel of tet_gemm.cu has its own test
    for mode in ("all", "0"):
        monkeypatch.setenv("NEKMF_TET_DMMA", mode)
        bwd, ipr = nk.Operator(std, nel, geom, nk.eBwdTrans), nk.Operator(std, nel, geom, nk.eIProductWRTBase)
        assert ("tet_dmma_kernel" in bwd.kernel_name) == (mode == "all"), bwd.kernel_name
        assert ("tet_dmma_kernel" in ipr.kernel_name) == (mode == "all"), ipr.kernel_name
        ob, oi = np.zeros(nel * el.nqTot), np.zeros(nel * el.nmTot)
        bwd.apply([c], [ob])
        ipr.apply([f], [oi])
        assert max(rel_errs(ob, want_b)) < 1e-12 and max(rel_errs(oi, want_i)) < 1e-12
        # device arrays at an odd offset (8-byte aligned only)
        cd = torch.zeros(c.size + 1, dtype=torch.float64, device="cuda")
        cd[1:] = torch.tensor(c, device="cuda")
        od = torch.zeros(ob.size + 1, dtype=torch.float64, device="cuda")
        bwd.apply([cd[1:]], [od[1:]])
        fd = torch.zeros(f.size + 1, dtype=torch.float64, device="cuda")
        fd[1:] = torch.tensor(f, device="cuda")
        oid = torch.zeros(oi.size + 1, dtype=torch.float64, device="cuda")
        ipr.apply([fd[1:]], [oid[1:]])
        torch.cuda.synchronize()
        assert np.array_equal(od[1:].cpu().numpy(), ob) and np.array_equal(oid[1:].cpu().numpy(), oi)
        res[mode] = (ob, oi)
    assert max(rel_errs(res["all"][0], res["0"][0])) < 1e-13 and max(rel_errs(res["all"][1], res["0"][1])) < 1e-13


@pytest.mark.parametrize("nm,nel", [(3, 5), (4, 1), (4, 38), (5, 37), (6, 2), (6, 1001), (7, 1), (7, 2), (7, 37), (7, 4099)])
@pytest.mark.parametrize("deformed", [False, True])
def test_pyr_dmma_bwd_iprod(nm, nel, deformed, monkeypatch):
    """BwdTrans / IProductWRTBase on pyramids at nm = 3..7 on FP64 tensor-core tiles (the PYR variant of prism_dmma.cu:
    mode lines of length nm - max(p, q), top-vertex correction through the entries (0,1), (1,0), (1,1)): against the oracle
    and against the runtime-sized kernel they replace (NEKMF_PYR_DMMA=0), odd element counts, caller arrays that are only
    8-byte aligned"""
    torch = _torch()
    nk = nekmf()
    rng = np.random.default_rng(13 * nm + nel)
    el = po.Elem(po.PYR, nm, nm + 1)
    std = nk.StdExpansion(nk.ePyramid, nm)
    jac, df = random_geometry(rng, 3, nel, el.nqTot, deformed)
    geom = nk.CoalescedGeomData(jac, df, deformed)
    c = rng.uniform(-1, 1, nel * el.nmTot)
    f = rng.uniform(-1, 1, nel * el.nqTot)
    want_b, want_i = el.bwdtrans(nel, c), el.iproduct(nel, deformed, jac, f)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("NEKMF_PYR_DMMA", mode)
        bwd, ipr = nk.Operator(std, nel, geom, nk.eBwdTrans), nk.Operator(std, nel, geom, nk.eIProductWRTBase)
        assert ("pyr_dmma_kernel" in bwd.kernel_name) == (mode == "1"), bwd.kernel_name
        assert ("pyr_dmma_kernel" in ipr.kernel_name) == (mode == "1"), ipr.kernel_name
        ob, oi = np.zeros(nel * el.nqTot), np.zeros(nel * el.nmTot)
        bwd.apply([c], [ob])
        ipr.apply([f], [oi])
        assert max(rel_errs(ob, want_b)) < 1e-12 and max(rel_errs(oi, want_i)) < 1e-12
        cd = torch.zeros(c.size + 1, dtype=torch.float64, device="cuda")
        cd[1:] = torch.tensor(c, device="cuda")
        od = torch.zeros(ob.size + 1, dtype=torch.float64, device="cuda")
        bwd.apply([cd[1:]], [od[1:]])
        fd = torch.zeros(f.size + 1, dtype=torch.float64, device="cuda")
        fd[1:] = torch.tensor(f, device="cuda")
        oid = torch.zeros(oi.size + 1, dtype=torch.float64, device="cuda")
        ipr.apply([fd[1:]], [oid[1:]])
        torch.cuda.synchronize()
        assert np.array_equal(od[1:].cpu().numpy(), ob) and np.array_equal(oid[1:].cpu().numpy(), oi)
        res[mode] = (ob, oi)
    assert max(rel_errs(res["1"][0], res["0"][0])) < 1e-13 and max(rel_errs(res["1"][1], res["0"][1])) < 1e-13


@pytest.mark.parametrize("deformed", [False, True])
@pytest.mark.parametrize("nm", list(range(2, 10)))
def test_pyr_physderiv_shape_kernel(nm, deformed):
    """pyramid PhysDeriv runs the compile-time sized quadrature-space kernel (shape_kernels.cuh, prism tensor structure
    with both base directions collapsing towards the apex, PhysDerivKernels.hpp:505-527): several batches plus a ragged
    last one, against the oracle; the direction overload too"""
    nk = nekmf()
    rng = np.random.default_rng(17 * nm + deformed)
    nel = 67
    el = po.Elem(po.PYR, nm, nm + 1)
    std = nk.StdExpansion(nk.ePyramid, nm)
    jac, df = random_geometry(rng, 3, nel, el.nqTot, deformed)
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, deformed))
    f = rng.uniform(-1, 1, nel * el.nqTot)
    outs = [np.zeros(nel * el.nqTot) for _ in range(3)]
    coll.ApplyOperator(nk.ePhysDeriv, f, *outs)
    assert "shape_op_kernel<Pyr,physderiv" in coll.m_ops[nk.ePhysDeriv].kernel_name, coll.m_ops[nk.ePhysDeriv].kernel_name
    for g, w in zip(outs, el.physderiv(nel, deformed, df, f)):
        assert max(rel_errs(g, w)) < 1e-12


@pytest.mark.parametrize("deformed", [False, True])
@pytest.mark.parametrize("nm", list(range(2, 10)))
@pytest.mark.parametrize("base", [False, True])
def test_pyr_shape_kernels(nm, deformed, base, monkeypatch):
    """pyramids in the compile-time sized family (shape_kernels.cuh, nm = 2..9 with the default quadrature: mode lines of
    length nm - max(p, q) from a first-mode table, two tensor contractions as for prisms, top-vertex term through the
    entries (0,1), (1,0), (1,1); Laplacian metric / chain rule with both base directions collapsing towards the apex):
    all five operators against the oracle, several batches plus a ragged last one.  base=True switches the tensor-core
    kernels layered on top off (NEKMF_PYR_DMMA=0, NEKMF_DENSE=0) so that every operator runs this family at every order"""
    if base:
        monkeypatch.setenv("NEKMF_PYR_DMMA", "0")
        monkeypatch.setenv("NEKMF_DENSE", "0")
    else:
        monkeypatch.delenv("NEKMF_PYR_DMMA", raising=False)
        monkeypatch.delenv("NEKMF_DENSE", raising=False)
    nk = nekmf()
    coll = run_all_ops(nk, po.PYR, nm, nm + 1, 67, deformed, np.random.default_rng(41 * nm + deformed))
    for op in (nk.eBwdTrans, nk.eIProductWRTBase, nk.ePhysDeriv, nk.eHelmholtz, nk.eIProductWRTDerivBase):
        want = "shape_op_kernel<Pyr"
        if not base:
            if op in (nk.eBwdTrans, nk.eIProductWRTBase) and (nm == 7 or (nm == 5 and (op == nk.eBwdTrans or deformed))):
                want = "pyr_dmma_kernel"  # tensor-core tiles (prism_dmma.cu) where they measured faster
            if op == nk.eHelmholtz and not deformed and nm <= 6:
                want = "dense_helm_kernel"  # DMMA coefficient-space kernel (dense_helm.cu; the pencil kernel wins from nm = 7)
        assert want in coll.m_ops[op].kernel_name, coll.m_ops[op].kernel_name


@pytest.mark.parametrize("nm,nel", [(5, 1), (5, 37), (6, 8), (6, 1001), (7, 1), (7, 7), (7, 8), (7, 9), (7, 37), (7, 4099)])
def test_tet_gemm_bwdtrans(nm, nel, monkeypatch):
    """BwdTrans on tetrahedra at nm = 5..7 as FP64 tensor-core GEMMs over eight elements (tet_gemm.cu: the two collapsed
    contractions pre-combined into per-p tables with the vertex / edge corrections folded in on the host, last contraction
    in the owning lane): against the oracle and against the pencil kernel (NEKMF_TET_GEMM=0, NEKMF_TET_DMMA=0), ragged
    last batches, caller arrays that are only 8-byte aligned"""
    torch = _torch()
    nk = nekmf()
    rng = np.random.default_rng(19 * nm + nel)
    el = po.Elem(po.TET, nm, nm + 1)
    std = nk.StdExpansion(nk.eTetrahedron, nm)
    jac, df = random_geometry(rng, 3, nel, el.nqTot, False)
    geom = nk.CoalescedGeomData(jac, df, False)
    c = rng.uniform(-1, 1, nel * el.nmTot)
    want = el.bwdtrans(nel, c)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("NEKMF_TET_GEMM", mode)
        monkeypatch.setenv("NEKMF_TET_DMMA", "0")
        bwd = nk.Operator(std, nel, geom, nk.eBwdTrans)
        assert ("tet_bwd_gemm_kernel" in bwd.kernel_name) == (mode == "1"), bwd.kernel_name
        ob = np.zeros(nel * el.nqTot)
        bwd.apply([c], [ob])
        assert max(rel_errs(ob, want)) < 1e-12
        cd = torch.zeros(c.size + 1, dtype=torch.float64, device="cuda")
        cd[1:] = torch.tensor(c, device="cuda")
        od = torch.zeros(ob.size + 1, dtype=torch.float64, device="cuda")
        bwd.apply([cd[1:]], [od[1:]])
        torch.cuda.synchronize()
        assert np.array_equal(od[1:].cpu().numpy(), ob)
        res[mode] = ob
    assert max(rel_errs(res["1"], res["0"])) < 1e-13


@pytest.mark.parametrize("nm,nel", [(5, 1), (5, 37), (6, 8), (6, 1001), (7, 1), (7, 7), (7, 8), (7, 9), (7, 37), (7, 4099)])
@pytest.mark.parametrize("deformed", [False, True])
def test_tet_gemm_iproduct(nm, nel, deformed, monkeypatch):
    """IProductWRTBase on tetrahedra at nm = 5..7 as FP64 tensor-core GEMMs over eight elements (tet_gemm.cu: input lines
    read straight from global memory and contracted with A_p in the loading lane, the two collapsed contractions as one
    table per p with the vertex / edge corrections as extra rows): against the oracle and against the pencil kernel,
    ragged last batches, caller arrays that are only 8-byte aligned"""
    torch = _torch()
    nk = nekmf()
    rng = np.random.default_rng(23 * nm + nel)
    el = po.Elem(po.TET, nm, nm + 1)
    std = nk.StdExpansion(nk.eTetrahedron, nm)
    jac, df = random_geometry(rng, 3, nel, el.nqTot, deformed)
    geom = nk.CoalescedGeomData(jac, df, deformed)
    f = rng.uniform(-1, 1, nel * el.nqTot)
    want = el.iproduct(nel, deformed, jac, f)
    res = {}
    for mode in ("all", "0"):
        monkeypatch.setenv("NEKMF_TET_GEMM", mode)
        monkeypatch.setenv("NEKMF_TET_DMMA", "0")
        ipr = nk.Operator(std, nel, geom, nk.eIProductWRTBase)
        assert ("tet_ip_gemm_kernel" in ipr.kernel_name) == (mode == "all"), ipr.kernel_name
        oi = np.zeros(nel * el.nmTot)
        ipr.apply([f], [oi])
        assert max(rel_errs(oi, want)) < 1e-12
        fd = torch.zeros(f.size + 1, dtype=torch.float64, device="cuda")
        fd[1:] = torch.tensor(f, device="cuda")
        oid = torch.zeros(oi.size + 1, dtype=torch.float64, device="cuda")
        ipr.apply([fd[1:]], [oid[1:]])
        torch.cuda.synchronize()
        assert np.array_equal(oid[1:].cpu().numpy(), oi)
        res[mode] = oi
    assert max(rel_errs(res["all"], res["0"])) < 1e-13


@pytest.mark.parametrize("nel", [1, 2, 3, 9, 100, 1001])
@pytest.mark.parametrize("nm", [5, 6, 7])
def test_prism_fused_dmma_helmholtz(nm, nel, monkeypatch):
    """General regular prisms at nm = 5..7: the whole Helmholtz chain fused in quadrature space on FP64 tensor-core tiles
    (prism_helm_dmma.cu: BwdTrans tiles, xi_0 / xi_1 derivatives and their transposes as tiles, xi_2 in the owning lane,
    IProduct tiles fed from the accumulator fragments): against the oracle for several lambda, against the quadrature-space
    pencil kernel, device arrays offset by one double"""
    monkeypatch.delenv("NEKMF_DENSE", raising=False)
    monkeypatch.delenv("NEKMF_PRISM_GENERAL", raising=False)
    monkeypatch.setenv("NEKMF_PRISM_FUSED", "1")
    torch = _torch()
    nk = nekmf()
    rng = np.random.default_rng(nm * 73 + nel)
    el = po.Elem(po.PRISM, nm, nm + 1)
    std = nk.StdExpansion(po.PRISM, nm, nm + 1)
    jac, df = random_geometry(rng, 3, nel, el.nqTot, False)
    coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, False))
    x = rng.uniform(-1, 1, nel * el.nmTot)
    for lam in (1.3, 0.0, 37.5):
        out = np.zeros(nel * el.nmTot)
        coll.ApplyOperator(nk.eHelmholtz, x, out, factors={nk.eFactorLambda: lam})
        check(out, el.helmholtz(nel, False, jac, df, lam, x), "Helmholtz(prism fused, lambda=%g)" % lam)
    assert "prism_helm_dmma_kernel" in coll.m_ops[nk.eHelmholtz].kernel_name, coll.m_ops[nk.eHelmholtz].kernel_name
    xd = torch.zeros(x.size + 1, dtype=torch.float64, device="cuda")
    xd[1:] = torch.from_numpy(x).cuda()
    yd = torch.zeros(x.size + 1, dtype=torch.float64, device="cuda")
    coll.ApplyOperator(nk.eHelmholtz, xd[1:], yd[1:], factors={nk.eFactorLambda: 37.5})
    torch.cuda.synchronize()
    assert np.array_equal(yd[1:].cpu().numpy(), out)
    monkeypatch.setenv("NEKMF_PRISM_FUSED", "0")
    monkeypatch.setenv("NEKMF_PRISM_GENERAL", "0")
    coll0 = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, False))
    out0 = np.zeros(nel * el.nmTot)
    coll0.ApplyOperator(nk.eHelmholtz, x, out0, factors={nk.eFactorLambda: 37.5})
    assert "shape_op_kernel" in coll0.m_ops[nk.eHelmholtz].kernel_name
    check(out, out0, "prism fused DMMA vs quadrature-space kernel")
