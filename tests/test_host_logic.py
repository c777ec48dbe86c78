"""CPU suite, part 2: the library loads and exports its ABI, the host-side tables agree with the
oracle, the library refuses to compute without a GPU, and the mesh / partition / sharded-CG host logic
is right (world_size 2 over gloo)."""
import os
import re
import socket

import numpy as np
import pytest

import pyoracle as po
from _util import ROOT, load_pkg_module, nekmf, rel_errs
import _nek_standin as si


def test_library_exports_every_declared_symbol():
    nk = nekmf()
    hdr = open(os.path.join(ROOT, "include", "nekmf_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(nekmf_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 40
    L = nk.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(nk.EXPORTS) == declared
    assert L.nekmf_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    nk = nekmf()
    if nk.device_count() > 0:
        pytest.skip("a GPU is present")
    std = nk.StdExpansion(nk.eHexahedron, 5, 6)
    coll = nk.Collection(std, 4, nk.CoalescedGeomData(np.ones(4), np.ones(36), False))
    with pytest.raises(nk.NekError, match="no CUDA device"):
        coll.Initialise(nk.eHelmholtz)
    with pytest.raises(nk.NekError, match="no CUDA device"):
        nk.AssemblyMap(np.zeros(4, dtype=np.int32), 1)


def test_argument_validation_needs_no_gpu():
    nk = nekmf()
    with pytest.raises(nk.NekError, match="out of range"):
        nk.AssemblyMap(np.array([0, 5], dtype=np.int32), 2)
    std = nk.StdExpansion(nk.eHexahedron, 5, 6)
    assert std.GetNcoeffs() == 125 and std.GetTotPoints() == 216
    assert nk.StdExpansion(nk.eTetrahedron, 7).GetNcoeffs() == 84
    assert nk.StdExpansion(nk.ePrism, 7).GetNcoeffs() == 196
    assert nk.StdExpansion(nk.eTetrahedron, 7).nq == [8, 7, 7]
    assert nk.StdExpansion(nk.ePrism, 7).nq == [8, 8, 7]
    assert nk.StdExpansion(nk.ePyramid, 7).GetNcoeffs() == 140 and nk.StdExpansion(nk.ePyramid, 7).nq == [8, 8, 7]
    with pytest.raises(nk.NekError):
        nk.StdExpansion(99, 4)
    assert nk.ImplementationTypeMap[nk.eB200] == "B200" and nk.SIZE_ImplementationType == nk.eB200 + 1


@pytest.mark.parametrize("ptype", [0, 1, 2])
def test_product_points_match_oracle(ptype):
    nk = nekmf()
    for n in range(2, 16):
        z, w, D = nk.points(ptype, n)
        zo, wo, Do = po.points(ptype, n)
        assert np.abs(z - zo).max() < 1e-14
        assert np.abs(w - wo).max() < 1e-14 * np.abs(wo).max() * 10
        assert np.abs(D - Do).max() < 1e-13 * np.abs(Do).max()


@pytest.mark.parametrize("shape", [po.QUAD, po.TRI, po.HEX, po.PRISM, po.PYR, po.TET])
def test_product_basis_tables_match_oracle(shape):
    nk = nekmf()
    for nm in range(2, 12):
        e, s = po.Elem(shape, nm, nm + 1), nk.StdExpansion(shape, nm)
        assert s.GetNcoeffs() == e.nmTot and s.GetTotPoints() == e.nqTot
        for d in range(e.dim):
            assert s.basis[d].rows == e.brows[d]
            assert max(rel_errs(s.basis[d].bdata, e.bdata[d])) < 1e-13
            assert max(rel_errs(s.basis[d].dbdata, e.dbdata[d])) < 1e-12


def test_structured_mesh_numbering():
    mesh_mod = load_pkg_module("mesh")
    nm = 4
    m = mesh_mod.StructuredHexMesh(3, 2, 4, nm)
    G = (3 * 3 + 1) * (2 * 3 + 1) * (4 * 3 + 1)
    assert m.nGlobal == G and m.nLocal == 24 * nm ** 3
    interior = (3 * 3 - 1) * (2 * 3 - 1) * (4 * 3 - 1)
    assert m.nDir == G - interior
    l2g = m.localToGlobal
    assert l2g.min() == 0 and l2g.max() == G - 1 and np.unique(l2g).size == G
    # element-interior modes (p,q,r >= 2) are never shared; vertex modes are shared by up to 8 elements
    counts = np.bincount(l2g, minlength=G)
    assert counts.max() == 8 and (counts == 1).sum() >= 24 * (nm - 2) ** 3
    # slabs: the two copies of the interface plane list the same DOFs in the same order
    a = mesh_mod.StructuredHexMesh(3, 2, 4, nm, slab=(0, 2))
    b = mesh_mod.StructuredHexMesh(3, 2, 4, nm, slab=(1, 2))
    assert a.peers == [1] and b.peers == [0]
    assert a.interface_lists[0].size == b.interface_lists[0].size == (3 * 3 + 1) * (2 * 3 + 1)
    assert a.nElmt + b.nElmt == m.nElmt
    # ownership: every DOF of the global problem is owned exactly once
    assert int(a.ownerMask.sum() + b.ownerMask.sum()) == G


def test_serial_cg_solves_helmholtz():
    """oracle CG (NekLinSysIterCG restatement) on a 3x3x3 hex mesh, P=4: converges and the solution has
    the expected spectral accuracy; the sharded reference with one rank is the same algorithm."""
    import _sharded_ref as sr
    mesh_mod = load_pkg_module("mesh")
    nk = nekmf()
    nm, lam = 5, 1.0
    mesh = mesh_mod.StructuredHexMesh(3, 3, 3, nm)
    el = po.Elem(po.HEX, nm, nm + 1)
    jac, df = mesh.geometry()
    rhs, u_exact = sr.helmholtz_rhs(None, mesh, el, jac, lam)
    diag = mesh.helmholtz_diagonal(nk.StdExpansion(nk.eHexahedron, nm).basis[0], lam)
    invdiag = 1.0 / diag[mesh.nDir:]
    # different summation orders in the dot products: compare fully converged solutions
    x, its, eps = el.cg(mesh.nElmt, False, jac, df, lam, mesh.nGlobal, mesh.nDir, mesh.localToGlobal, None,
                        invdiag, rhs, tol=1e-13)
    x2, its2, eps2 = sr.sharded_cg(None, mesh, el, jac, df, lam, rhs, invdiag, tol=1e-13)
    assert abs(its - its2) <= 1 and np.abs(x - x2).max() < 1e-11 * np.abs(x).max()
    uq = el.bwdtrans(mesh.nElmt, po.global_to_local(mesh.localToGlobal, None, x))
    assert np.abs(uq - u_exact).max() < 2e-4
    # the matrix-free diagonal equals the diagonal of the assembled operator
    e0 = np.zeros(mesh.nGlobal)
    probe = [mesh.nDir, mesh.nDir + 7, mesh.nGlobal - 1]
    for g in probe:
        e0[:] = 0
        e0[g] = 1
        col = po.assemble(mesh.localToGlobal, None,
                          el.helmholtz(mesh.nElmt, False, jac, df, lam, po.global_to_local(mesh.localToGlobal, None, e0)),
                          mesh.nGlobal)
        assert abs(col[g] - diag[g]) < 1e-12 * abs(diag[g])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, nm, lam, out, part=None):
    import torch.distributed as dist
    import _sharded_ref as sr
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh_mod = load_pkg_module("mesh")
    nk = nekmf()
    mesh = (mesh_mod.StructuredHexMesh(3, 2, 4, nm, slab=(rank, world)) if part is None else
            mesh_mod.StructuredHexMesh(3, 2, 4, nm, part=part, rank=rank))
    el = po.Elem(po.HEX, nm, nm + 1)
    jac, df = mesh.geometry()
    rhs, _ = sr.helmholtz_rhs(dist, mesh, el, jac, lam)
    diag = mesh.helmholtz_diagonal(nk.StdExpansion(nk.eHexahedron, nm).basis[0], lam)
    sr.exchange_add(dist, diag, mesh.peers, mesh.interface_lists)
    x, its, eps = sr.sharded_cg(dist, mesh, el, jac, df, lam, rhs, 1.0 / diag[mesh.nDir:], tol=1e-13)
    # return the solution on the global lattice for comparison
    np.save(os.path.join(out, "x_%d.npy" % rank), x[mesh.lattice_ids])
    np.save(os.path.join(out, "meta_%d.npy" % rank), np.array([its, eps, mesh.gz0, mesh.gz1, mesh.gy0, mesh.gy1, mesh.gx0,
                                                                mesh.gx1]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,part", [(2, None), (4, (1, 2, 2))])
def test_sharded_cg_gloo(tmp_path, world, part):
    """two ranks in z-slabs / four ranks in a 1 x 2 x 2 box partition (the middle line of DOFs is held by all four),
    gloo: interface exchange + masked dots + all-reduce reproduce the serial solve (iteration count, solution)."""
    import torch.multiprocessing as mp
    import _sharded_ref as sr
    nm, lam = 4, 1.0
    mp.spawn(_gloo_worker, args=(world, _free_port(), nm, lam, str(tmp_path), part), nprocs=world, join=True)
    mesh_mod = load_pkg_module("mesh")
    nk = nekmf()
    mesh = mesh_mod.StructuredHexMesh(3, 2, 4, nm)
    el = po.Elem(po.HEX, nm, nm + 1)
    jac, df = mesh.geometry()
    rhs, _ = sr.helmholtz_rhs(None, mesh, el, jac, lam)
    diag = mesh.helmholtz_diagonal(nk.StdExpansion(nk.eHexahedron, nm).basis[0], lam)
    x, its, eps = sr.sharded_cg(None, mesh, el, jac, df, lam, rhs, 1.0 / diag[mesh.nDir:], tol=1e-13)
    xs = x[mesh.lattice_ids]
    for r in range(world):
        xr = np.load(os.path.join(str(tmp_path), "x_%d.npy" % r))
        meta = np.load(os.path.join(str(tmp_path), "meta_%d.npy" % r))
        assert abs(int(meta[0]) - its) <= 2
        z0, z1, y0, y1, x0, x1 = (int(v) for v in meta[2:8])
        assert np.abs(xr - xs[z0:z1 + 1, y0:y1 + 1, x0:x1 + 1]).max() < 1e-10 * np.abs(xs).max()


def test_cpp_collections_mirror_builds_and_refuses_without_gpu():
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")], check=True)
    nk = nekmf()
    if nk.device_count() > 0:
        pytest.skip("a GPU is present: the real run is in the gpu suite")
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "TestCollectionB200")], capture_output=True, text=True)
    assert r.returncode == 77 and "no CUDA device" in r.stdout


def test_cpp_collection_optimisation_mirror():
    """the C++ host side's CollectionOptimisation (session <COLLECTIONS> block, defaults, lookup, error messages):
    tests/cpp/TestCollectionOptimisation.cpp, needs no GPU"""
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")], check=True)
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "TestCollectionOptimisation")], capture_output=True, text=True)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout + r.stderr


def test_config1_quad_helmholtz_solve_oracle():
    """BASELINE configs[0] on the CPU: 2-D Helmholtz on a structured quad mesh at P=5 (nm=6, nq=7), the
    reference's Helmholtz2D_modal set-up (lambda=1, u = sin(pi x) sin(pi y), homogeneous Dirichlet).  The
    MatrixFree-path CG solve converges to the analytic solution with spectral accuracy, and the matrix-free
    operator equals the StdMat-style dense elemental matrix applied element by element."""
    mesh_mod = load_pkg_module("mesh")
    nk = nekmf()
    nm, lam = 6, 1.0
    mesh = mesh_mod.StructuredQuadMesh(4, 4, nm)
    el = po.Elem(po.QUAD, nm, nm + 1)
    jac, df = mesh.geometry()
    X, Y = mesh.quad_coords(el.Z[0])
    u = np.sin(np.pi * X) * np.sin(np.pi * Y)
    rhs = po.assemble(mesh.localToGlobal, None, -el.iproduct(mesh.nElmt, False, jac, -(lam + 2 * np.pi ** 2) * u),
                      mesh.nGlobal)
    rhs[:mesh.nDir] = 0.0
    diag = mesh.helmholtz_diagonal(nk.StdExpansion(nk.eQuadrilateral, nm).basis[0], lam)
    x, its, eps = el.cg(mesh.nElmt, False, jac, df, lam, mesh.nGlobal, mesh.nDir, mesh.localToGlobal, None,
                        1.0 / diag[mesh.nDir:], rhs, tol=1e-12)
    uq = el.bwdtrans(mesh.nElmt, po.global_to_local(mesh.localToGlobal, None, x))
    assert np.abs(uq - u).max() < 5e-6 and its < 200
    # StdMat comparator: dense elemental matrix from unit vectors (all elements are congruent)
    n = el.nmTot
    A = np.zeros((n, n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = 1.0
        A[:, j] = el.helmholtz(1, False, jac[:1], df.reshape(4, -1)[:, :1].reshape(-1).copy(), lam, e)
    loc = np.random.default_rng(0).uniform(-1, 1, mesh.nLocal)
    mf = el.helmholtz(mesh.nElmt, False, jac, df, lam, loc)
    sm = (loc.reshape(mesh.nElmt, n) @ A.T).reshape(-1)
    assert max(rel_errs(mf, sm)) < 1e-12


def test_collection_optimisation_mirror(tmp_path):
    """Collections::CollectionOptimisation (CollectionOptimisation.cpp:52-316) as mirrored in nekmf.py: constructor
    defaults, the <COLLECTIONS> block of a session file (DEFAULT, MAXSIZE, OPERATOR / ELEMENT with ORDER "*" or a
    sequence), the (shape, order) -> shape default -> eNoCollection lookup, and the reference's error messages.  This
    is how ExpList selects eB200; needs no GPU."""
    nk = nekmf()
    hex5, tet2, tet6 = nk.StdExpansion(nk.eHexahedron, 5, 6), nk.StdExpansion(nk.eTetrahedron, 2, 3), nk.StdExpansion(nk.eTetrahedron, 6, 7)
    # the unit tests' form: dummy session + explicit type (TestHexCollection.cpp:3699-3704)
    opt = si.CollectionOptimisation(None, nk.eB200)
    assert opt.GetOperatorImpMap(hex5) == nk.SetFixedImpType(nk.eB200)
    assert not opt.SetByXml() and not opt.IsUsingAutotuning() and opt.GetMaxCollectionSize() == 0
    # no session, no type: IterPerExp, StdMat for orders 1..4, PhysDeriv NoCollection / SumFac for orders 1, 2
    opt = si.CollectionOptimisation()
    assert opt.GetDefaultImplementationType() == nk.eIterPerExp
    low, high = opt.GetOperatorImpMap(tet2), opt.GetOperatorImpMap(tet6)
    assert low[nk.eBwdTrans] == nk.eStdMat and low[nk.ePhysDeriv] == nk.eSumFac
    assert high[nk.eHelmholtz] == nk.eIterPerExp and high[nk.ePhysDeriv] == nk.eNoCollection
    # session file: case-insensitive DEFAULT, per-operator overrides
    xml = """<NEKTAR><COLLECTIONS DEFAULT="b200" MAXSIZE="64">
               <OPERATOR TYPE="Helmholtz"><ELEMENT TYPE="H" ORDER="2-4,7" IMPTYPE="MatrixFree"/>
                                          <ELEMENT TYPE="A" ORDER="*" IMPTYPE="StdMat"/></OPERATOR>
             </COLLECTIONS></NEKTAR>"""
    path = tmp_path / "session.xml"
    path.write_text(xml)
    for src in (xml, str(path)):
        opt = si.CollectionOptimisation(src)
        assert opt.GetDefaultImplementationType() == nk.eB200 and opt.GetMaxCollectionSize() == 64 and opt.SetByXml()
        assert opt.GetOperatorImpMap(hex5)[nk.eHelmholtz] == nk.eB200            # order 5 is not in 2-4,7
        assert opt.GetOperatorImpMap(nk.StdExpansion(nk.eHexahedron, 7, 8))[nk.eHelmholtz] == nk.eMatrixFree
        assert opt.GetOperatorImpMap(nk.StdExpansion(nk.eHexahedron, 3, 4))[nk.eBwdTrans] == nk.eB200
        assert opt.GetOperatorImpMap(tet6)[nk.eHelmholtz] == nk.eStdMat and opt.GetOperatorImpMap(tet6)[nk.eBwdTrans] == nk.eB200
    # an explicit constructor type wins over DEFAULT (CollectionOptimisation.cpp:150-153)
    assert si.CollectionOptimisation(xml, nk.eMatrixFree).GetOperatorImpMap(hex5)[nk.eBwdTrans] == nk.eMatrixFree
    assert si.CollectionOptimisation('<NEKTAR><COLLECTIONS DEFAULT="auto"/></NEKTAR>').IsUsingAutotuning()
    assert si.GenerateSeqVector("1-3, 5") == [1, 2, 3, 5]
    for bad, msg in (('<FOO/>', "Unable to find NEKTAR tag"),
                     ('<NEKTAR><COLLECTIONS DEFAULT="Fast"/></NEKTAR>', "Unknown default collection scheme: Fast"),
                     ('<NEKTAR><COLLECTIONS><THING/></COLLECTIONS></NEKTAR>', "Only OPERATOR tags"),
                     ('<NEKTAR><COLLECTIONS><OPERATOR/></COLLECTIONS></NEKTAR>', "Missing TYPE in OPERATOR tag"),
                     ('<NEKTAR><COLLECTIONS><OPERATOR TYPE="Mass"/></COLLECTIONS></NEKTAR>', "Unknown OPERATOR type Mass"),
                     ('<NEKTAR><COLLECTIONS><OPERATOR TYPE="BwdTrans"><ELEMENT TYPE="X" ORDER="*" IMPTYPE="B200"/></OPERATOR>'
                      '</COLLECTIONS></NEKTAR>', "Unknown element type X"),
                     ('<NEKTAR><COLLECTIONS><OPERATOR TYPE="BwdTrans"><ELEMENT TYPE="H" ORDER="*" IMPTYPE="Cuda"/></OPERATOR>'
                      '</COLLECTIONS></NEKTAR>', "Unknown IMPTYPE type Cuda"),
                     ('<NEKTAR><COLLECTIONS><OPERATOR TYPE="BwdTrans"><ELEMENT TYPE="H" IMPTYPE="B200"/></OPERATOR>'
                      '</COLLECTIONS></NEKTAR>', "Missing ORDER in ELEMENT tag")):
        with pytest.raises(nk.NekError, match=msg):
            si.CollectionOptimisation(bad)
    # a Collection built from the map refuses implementation types that are not registered in this library
    coll = nk.Collection(tet6, 3, nk.CoalescedGeomData(np.ones(3), np.ones(27), False),
                         si.CollectionOptimisation(xml).GetOperatorImpMap(tet6))
    with pytest.raises(nk.NekError, match="no operator registered for key"):
        coll.Initialise(nk.eHelmholtz)


def test_explist_create_collections_mirror(monkeypatch):
    """MultiRegions::ExpList::CreateCollections (ExpList.cpp:5005-5151) as mirrored in nekmf.py: one pass per shape in
    LibUtilities::ShapeType order; a collection ends at a gap in the coefficient / quadrature arrays, a change of
    nCoeffs / nPhys / deformed-ness or at collmax members; the call sites hand each collection its slice of the
    ExpList arrays.  Needs no GPU: ApplyOperator is replaced by a recorder."""
    nk = nekmf()
    hex4, hex5, tet4 = nk.StdExpansion(nk.eHexahedron, 4), nk.StdExpansion(nk.eHexahedron, 5), nk.StdExpansion(nk.eTetrahedron, 4)

    def elem(std, deformed, tag):
        npt = std.GetTotPoints() if deformed else 1
        return (std, np.full(npt, float(tag)), np.full(std.dim * std.dim * npt, 10.0 + tag), deformed)

    mesh = [elem(hex4, False, 0), elem(hex4, False, 1), elem(hex4, False, 2), elem(hex4, True, 3), elem(hex4, True, 4),
            elem(tet4, False, 5), elem(tet4, False, 6), elem(hex5, False, 7), elem(hex4, False, 8), elem(hex4, False, 9)]
    exp = si.ExpList(mesh).CreateCollections(nk.eB200)
    members = [(c.m_stdExp.DetShapeType(), c.m_stdExp.nm, c.m_nElmt, bool(c.m_geomData.IsDeformed())) for c in exp.m_collections]
    # tetrahedra come before hexahedra (ShapeType order); the tets at positions 5, 6 break the contiguity of the hexes
    assert members == [(nk.eTetrahedron, 4, 2, False), (nk.eHexahedron, 4, 3, False), (nk.eHexahedron, 4, 2, True),
                       (nk.eHexahedron, 5, 1, False), (nk.eHexahedron, 4, 2, False)]
    nc4, nq4, nct, nqt = hex4.GetNcoeffs(), hex4.GetTotPoints(), tet4.GetNcoeffs(), tet4.GetTotPoints()
    assert exp.m_coll_coeff_offset == [5 * nc4, 0, 3 * nc4, 5 * nc4 + 2 * nct, 5 * nc4 + 2 * nct + hex5.GetNcoeffs()]
    assert exp.m_coll_phys_offset == [5 * nq4, 0, 3 * nq4, 5 * nq4 + 2 * nqt, 5 * nq4 + 2 * nqt + hex5.GetTotPoints()]
    # coalesced geometry in the reference layout: jac [nElmt(*nq)], df [ndf][nElmt(*nq)]
    g = exp.m_collections[1].m_geomData
    assert np.array_equal(g.GetJac(), [0.0, 1.0, 2.0]) and np.array_equal(np.asarray(g.GetDerivFactors()).reshape(9, 3)[4], [10.0, 11.0, 12.0])
    g = exp.m_collections[2].m_geomData
    assert g.GetJac().size == 2 * nq4 and np.array_equal(np.asarray(g.GetDerivFactors()).reshape(9, 2 * nq4)[0, nq4 - 1:nq4 + 1], [13.0, 14.0])
    # MAXSIZE from the session caps the members of a collection
    capped = si.ExpList(mesh, '<NEKTAR><COLLECTIONS DEFAULT="B200" MAXSIZE="2"/></NEKTAR>').CreateCollections()
    assert [c.m_nElmt for c in capped.m_collections] == [2, 2, 1, 2, 1, 2]
    assert all(c.m_impTypes[nk.eHelmholtz] == nk.eB200 for c in capped.m_collections)
    # call sites: every collection gets its own slice of the ExpList arrays
    calls = []
    monkeypatch.setattr(nk.Collection, "ApplyOperator", lambda self, op, *a, **kw: calls.append((op, [x.size for x in a], [x[0] for x in a], kw)))
    coeffs, phys = np.arange(exp.GetNcoeffs(), dtype=np.float64), np.arange(exp.GetTotPoints(), dtype=np.float64)
    exp.BwdTrans(coeffs, phys)
    assert [c[1] for c in calls] == [[2 * nct, 2 * nqt], [3 * nc4, 3 * nq4], [2 * nc4, 2 * nq4], [hex5.GetNcoeffs(), hex5.GetTotPoints()], [2 * nc4, 2 * nq4]]
    assert [c[2] for c in calls] == [[float(a), float(b)] for a, b in zip(exp.m_coll_coeff_offset, exp.m_coll_phys_offset)]
    calls.clear()
    d = [np.zeros(exp.GetTotPoints()) for _ in range(3)]
    exp.PhysDeriv(phys, *d)
    exp.IProductWRTBase(phys, coeffs)
    exp.GeneralMatrixOp_Helmholtz(coeffs, np.zeros_like(coeffs), {nk.eFactorLambda: 2.0})
    assert [c[0] for c in calls] == [nk.ePhysDeriv] * 5 + [nk.eIProductWRTBase] * 5 + [nk.eHelmholtz] * 5
    assert all(len(c[1]) == 4 for c in calls[:5]) and calls[-1][3] == {"factors": {nk.eFactorLambda: 2.0}}
    assert calls[5][1] == [2 * nqt, 2 * nct] and calls[10][1] == [2 * nct, 2 * nct]


def test_interface_from_universal_maps():
    """the gslib set-up restated (mesh.interface_from_universal_maps): against the analytic z-slab interfaces (the
    whole shared lattice plane in lattice order, the lower rank owning it), on a 2 x 2 x 2 box partition (7 neighbours,
    edge DOFs on 4 ranks, the centre DOF on 8) where the pairwise exchange-add must reproduce the unpartitioned
    assembly, and on a hand-made case with a DOF shared by three ranks"""
    mesh_mod = load_pkg_module("mesh")
    R = 3
    meshes = [mesh_mod.StructuredHexMesh(3, 2, 5, 4, slab=(r, R)) for r in range(R)]
    for r, m in enumerate(meshes):
        want_peers = [q for q in (r - 1, r + 1) if 0 <= q < R]
        assert m.peers == want_peers
        for q, lst in zip(m.peers, m.interface_lists):
            plane = m.lattice_ids[0 if q < r else -1].reshape(-1)
            assert np.array_equal(lst, plane)
        mask = np.ones((m.Gzl, m.Gyl, m.Gxl))
        if r > 0:
            mask[0] = 0.0
        want = np.empty(m.nGlobal)
        want[m.lattice_ids.reshape(-1)] = mask.reshape(-1)
        assert np.array_equal(m.ownerMask, want)
    # ---- box partition: exchange-add of the rank-local assemblies == the unpartitioned assembly
    nm = 3
    full = mesh_mod.StructuredHexMesh(4, 4, 4, nm)
    rng = np.random.default_rng(3)
    loc_full = rng.uniform(-1, 1, full.nLocal)
    want = po.assemble(full.localToGlobal, None, loc_full, full.nGlobal)
    boxes = [mesh_mod.StructuredHexMesh(4, 4, 4, nm, part=(2, 2, 2), rank=r) for r in range(8)]
    assert all(len(b.peers) == 7 for b in boxes)
    glob = []
    for b in boxes:
        mine = loc_full.reshape(full.nElmt, nm ** 3)[b.element_ids].reshape(-1)
        glob.append(po.assemble(b.localToGlobal, None, mine, b.nGlobal))
    sent = [[g[l].copy() for l in b.interface_lists] for b, g in zip(boxes, glob)]
    for r, b in enumerate(boxes):
        for k, q in enumerate(b.peers):
            back = boxes[q].peers.index(r)
            assert np.array_equal(b.universal[b.interface_lists[k]], boxes[q].universal[boxes[q].interface_lists[back]])
            glob[r][b.interface_lists[k]] += sent[q][back]
    owned = 0
    for b, g in zip(boxes, glob):
        ids = full.lattice_ids[b.gz0:b.gz1 + 1, b.gy0:b.gy1 + 1, b.gx0:b.gx1 + 1]
        assert np.abs(g[b.lattice_ids] - want[ids]).max() < 1e-13
        owned += int(b.ownerMask.sum())
    assert owned == full.nGlobal
    multiplicity = np.zeros(full.nGlobal + 1, dtype=int)
    for b in boxes:
        np.add.at(multiplicity, b.universal, 1)
    assert multiplicity.max() == 8 and (multiplicity == 4).sum() > 0
    # three ranks around a shared point: ids 7 (all three), 8 (ranks 0, 1), 9 (ranks 1, 2), 0 = not exchanged
    maps = [np.array([5, 7, 8, 0]), np.array([8, 9, 7, 0, 11]), np.array([0, 9, 12, 7])]
    vals = [np.array([1.0, 2.0, 3.0, 4.0]), np.array([10.0, 20.0, 30.0, 40.0, 50.0]), np.array([100.0, 200.0, 300.0, 400.0])]
    setups = [mesh_mod.interface_from_universal_maps(maps, r) for r in range(3)]
    assert setups[0][0] == [1, 2] and setups[1][0] == [0, 2] and setups[2][0] == [0, 1]
    assert [l.tolist() for l in setups[1][1]] == [[2, 0], [2, 1]]  # ascending universal id: 7 then 8; 7 then 9
    sent = [[vals[r][l].copy() for l in setups[r][1]] for r in range(3)]
    out = [v.copy() for v in vals]
    for r in range(3):
        for k, p in enumerate(setups[r][0]):
            out[r][setups[r][1][k]] += sent[p][setups[p][0].index(r)]
    total = {}
    for r in range(3):
        for u, v in zip(maps[r], vals[r]):
            total[u] = total.get(u, 0.0) + v
    for r in range(3):
        for i, u in enumerate(maps[r]):
            assert out[r][i] == (total[u] if u != 0 else vals[r][i]), (r, i)
    counted = {}
    for r in range(3):
        for u, w in zip(maps[r], setups[r][2]):
            if u != 0:
                counted[u] = counted.get(u, 0.0) + w
    assert all(v == 1.0 for v in counted.values())
    with pytest.raises(ValueError):
        mesh_mod.interface_from_universal_maps([np.array([3, 3]), np.array([3])], 0)


def _gloo_universal_worker(rank, world, port, outdir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh_mod = load_pkg_module("mesh")
    m = mesh_mod.StructuredHexMesh(2, 2, 4, 3, slab=(rank, world))
    peers, lists, owner = mesh_mod.interface_from_universal_map(dist, m.universal)
    ok = peers == m.peers and all(np.array_equal(a, b) for a, b in zip(lists, m.interface_lists))
    np.save(os.path.join(outdir, "uok_%d.npy" % rank), np.array([1.0 if ok else 0.0, owner[m.nDir:].sum()]))
    dist.destroy_process_group()


def test_interface_from_universal_map_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_gloo_universal_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    res = [np.load(os.path.join(str(tmp_path), "uok_%d.npy" % r)) for r in range(2)]
    assert all(r[0] == 1.0 for r in res)
    # every interior DOF of the full mesh is owned exactly once
    full = load_pkg_module("mesh").StructuredHexMesh(2, 2, 4, 3)
    assert res[0][1] + res[1][1] == full.nGlobal - full.nDir


def test_flop_model_counts_the_reference_loops():
    """tools/flop_model.py (the FP64 fractions of the sweep tables): the closed-form pass lengths against a literal count of the
    multiply-adds in the reference's BwdTrans loop nests (BwdTransKernels.hpp:78-484), every shape, nm = 2..9"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from flop_model import algorithmic_flops

    def count(shape, nm):
        nq = nm + 1
        if shape == "Hex":
            return 2 * (nq * nm ** 3 + nq ** 2 * nm ** 2 + nq ** 3 * nm)
        if shape == "Quad":
            return 2 * (nq * nm ** 2 + nq ** 2 * nm)
        if shape == "Tri":
            nq0, nq1 = nq, nm
            n = sum(nm - p for p in range(nm)) * nq1          # q-contraction per (j, p)
            return 2 * (n + nq1 * nq0 * nm)                      # p-contraction per (j, i)
        nq0, nq1, nq2 = nq, (nm if shape == "Tet" else nq), nm
        if shape == "Prism":
            s1 = nq2 * sum(nm * (nm - p) for p in range(nm))
            s2 = nq2 * nq1 * nm * nm
        elif shape == "Pyr":
            s1 = nq2 * sum(nm - max(p, q) for p in range(nm) for q in range(nm))
            s2 = nq2 * nq1 * nm * nm
        else:
            s1 = nq2 * sum(nm - p - q for p in range(nm) for q in range(nm - p))
            s2 = nq2 * nq1 * sum(nm - p for p in range(nm))
        return 2 * (s1 + s2 + nq2 * nq1 * nq0 * nm)

    for shape in ("Hex", "Quad", "Tri", "Prism", "Pyr", "Tet"):
        for nm in range(2, 10):
            assert algorithmic_flops("BwdTrans", shape, nm, nm + 1) == count(shape, nm), (shape, nm)
            # the composite operators are built from the same pass count
            sf = count(shape, nm)
            assert algorithmic_flops("Helmholtz", shape, nm, nm + 1) > (5 if shape not in ("Quad", "Tri") else 4) * sf
