"""GPU suite: the ExpList call sites (MultiRegions/ExpList.cpp:1262-1284 BwdTrans, 1465-1504 IProductWRTBase,
1961-1989 PhysDeriv, 2359-2397 GeneralMatrixOp) driving SEVERAL collections of one field through the eB200 operators
with `array + offset` slices -- mixed shapes, a regular / deformed split, an odd coefficient count that leaves the
following collections 8-byte (not 16-byte) aligned -- against the CPU oracle, collection by collection.  The grouping
logic is the restated CreateCollections of tests/_nek_standin.py; everything below ApplyOperator is the product."""
import numpy as np
import pytest

import pyoracle as po
import _nek_standin as si
from _util import nekmf, random_geometry, rel_errs

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _mesh(nk, rng):
    hex4, hex5 = nk.StdExpansion(nk.eHexahedron, 4), nk.StdExpansion(nk.eHexahedron, 5)
    tet4, pri4 = nk.StdExpansion(nk.eTetrahedron, 4), nk.StdExpansion(nk.ePrism, 4)

    def elem(std, deformed):
        jac, df = random_geometry(rng, 3, 1, std.GetTotPoints(), deformed)
        return (std, jac, df, deformed)

    # hexes: 3 regular | 2 deformed | (2 tets, 3 prisms in between) | 1 of another order (125 coefficients: odd) |
    # 3 regular again, now at an odd offset
    return ([elem(hex4, False) for _ in range(3)] + [elem(hex4, True) for _ in range(2)] +
            [elem(tet4, False) for _ in range(2)] + [elem(pri4, True) for _ in range(3)] + [elem(hex5, False)] +
            [elem(hex4, False) for _ in range(3)])


def _oracle_elem(c):
    std = c.m_stdExp
    shape = {si.eHexahedron: po.HEX, si.eTetrahedron: po.TET, si.ePrism: po.PRISM}[std.DetShapeType()]
    return po.Elem(shape, std.nm, std.nq[0])


@pytest.mark.parametrize("memory", ["host", "device"])
def test_explist_call_sites_multi_collection(memory):
    import torch
    nk = nekmf()
    rng = np.random.default_rng(2024)
    exp = si.ExpList(_mesh(nk, rng)).CreateCollections(nk.eB200)
    assert len(exp.m_collections) == 6
    # at least one collection starts at an odd (8-byte aligned only) offset in both arrays
    assert any(o % 2 for o in exp.m_coll_coeff_offset)
    nC, nP, lam = exp.GetNcoeffs(), exp.GetTotPoints(), 0.8
    coeffs, phys = rng.uniform(-1, 1, nC), rng.uniform(-1, 1, nP)

    def arr(a):
        return torch.tensor(a, device="cuda") if memory == "device" else a.copy()

    def back(a):
        if memory == "device":
            torch.cuda.synchronize()
            return a.cpu().numpy()
        return a

    bwd, ipr, helm = arr(np.zeros(nP)), arr(np.zeros(nC)), arr(np.zeros(nC))
    d = [arr(np.zeros(nP)) for _ in range(3)]
    exp.BwdTrans(arr(coeffs), bwd)
    exp.IProductWRTBase(arr(phys), ipr)
    exp.PhysDeriv(arr(phys), *d)
    exp.GeneralMatrixOp_Helmholtz(arr(coeffs), helm, {nk.eFactorLambda: lam})
    bwd, ipr, helm, d = back(bwd), back(ipr), back(helm), [back(x) for x in d]
    for c, c0, c1, p0, p1 in exp._spans():
        el, n, g = _oracle_elem(c), c.m_nElmt, c.m_geomData
        jac, df, deformed = g.GetJac(), g.GetDerivFactors(), g.IsDeformed()
        assert max(rel_errs(bwd[p0:p1], el.bwdtrans(n, coeffs[c0:c1]))) < TOL
        assert max(rel_errs(ipr[c0:c1], el.iproduct(n, deformed, jac, phys[p0:p1]))) < TOL
        for got, want in zip(d, el.physderiv(n, deformed, df, phys[p0:p1])):
            assert max(rel_errs(got[p0:p1], want)) < TOL
        assert max(rel_errs(helm[c0:c1], el.helmholtz(n, deformed, jac, df, lam, coeffs[c0:c1]))) < TOL
    kernels = sorted({c.m_ops[nk.eHelmholtz].kernel_name for c in exp.m_collections})
    assert len(kernels) >= 3, kernels  # regular hex, deformed hex, tet, prism kernels all took part
