"""CPU suite: an INDEPENDENT statement of what the operators compute, from the mathematics rather than from the
reference's (or this repository's) sum-factorised loops -- the role the LocalRegions routines play in the
reference's own unit tests (library/UnitTests/Collections/TestHexCollection.cpp compares the MatrixFree collection
with HexExp::v_HelmholtzMatrixOp_MatFree, LocalRegions/HexExp.cpp:1953-2012, and StdExpansion3D::PhysDeriv,
StdRegions/StdExpansion3D.cpp:347-405).

  * the 1-D basis tables (Modified_A / Modified_B; Modified_C re-packs B) against the closed-form definitions
    coded in LibUtilities/Foundations/Basis.cpp:392-512 (Karniadakis & Sherwin), with the Jacobi polynomials and
    their ANALYTIC derivatives from scipy -- no collocation differentiation matrix involved;
  * the elemental operators of tensor-product shapes as dense matrices assembled point by point:
        BwdTrans   u(x_q)      = sum_m phi_m(x_q) c_m
        IProduct   I_m         = sum_q w_q J_q phi_m(x_q) f(x_q)
        PhysDeriv  du/dx_c     = sum_d df[c*dim+d] d u / d xi_d
        Helmholtz  H_mn        = sum_q w_q J_q [ lambda phi_m phi_n + grad phi_m . grad phi_n ]
    with phi_m(xi) = prod_d A_{m_d}(xi_d) and grad = df^T grad_xi, for regular and deformed geometry.
The oracle (and through the GPU parity tests the CUDA path) must reproduce these to 1e-12."""
import numpy as np
import pytest
from scipy.special import eval_jacobi

import pyoracle as po
from _util import random_geometry, rel_errs


def _jac(n, a, b, z):
    return eval_jacobi(n, a, b, z) if n >= 0 else np.zeros_like(z)


def _djac(n, a, b, z):
    """d/dz P_n^{a,b}(z) = (n + a + b + 1)/2 P_{n-1}^{a+1,b+1}(z)"""
    return 0.5 * (n + a + b + 1) * _jac(n - 1, a + 1, b + 1, z) if n >= 1 else np.zeros_like(z)


def modified_a(nm, z):
    """Basis.cpp:392-421: A_0 = (1-z)/2, A_1 = (1+z)/2, A_p = (1-z)/2 (1+z)/2 P^{1,1}_{p-2}; returns (A, dA/dz)"""
    A, dA = np.zeros((nm, z.size)), np.zeros((nm, z.size))
    A[0], dA[0] = 0.5 * (1 - z), -0.5
    A[1], dA[1] = 0.5 * (1 + z), 0.5
    bub, dbub = 0.25 * (1 - z) * (1 + z), -0.5 * z
    for p in range(2, nm):
        A[p] = bub * _jac(p - 2, 1, 1, z)
        dA[p] = dbub * _jac(p - 2, 1, 1, z) + bub * _djac(p - 2, 1, 1, z)
    return A, dA


def modified_b(nm, z):
    """Basis.cpp:423-511, rows packed p-major with q running fastest (q = 0 .. nm-1-p):
    p = 0: A_q;  p = 1: (1-z)/2 then (1-z)/2 (1+z)/2 P^{1,1}_{q-2} for q >= 2 (row q = 1 is skipped by the packing:
    nm - 1 rows);  p >= 2: ((1-z)/2)^p, then ((1-z)/2)^p (1+z)/2 P^{2p-1,1}_{q-1}."""
    rows, drows = [], []
    A, dA = modified_a(nm, z)
    rows += list(A)
    drows += list(dA)
    om, op_ = 0.5 * (1 - z), 0.5 * (1 + z)
    if nm > 1:
        rows.append(om)
        drows.append(np.full_like(z, -0.5))
        for q in range(2, nm):
            rows.append(A[q])
            drows.append(dA[q])
    for p in range(2, nm):
        omp, domp = om ** p, -0.5 * p * om ** (p - 1)
        rows.append(omp)
        drows.append(domp)
        for q in range(1, nm - p):
            P, dP = _jac(q - 1, 2 * p - 1, 1, z), _djac(q - 1, 2 * p - 1, 1, z)
            rows.append(omp * op_ * P)
            drows.append(domp * op_ * P + omp * 0.5 * P + omp * op_ * dP)
    return np.array(rows), np.array(drows)


@pytest.mark.parametrize("nm", [2, 3, 5, 8, 11])
def test_modified_a_tables_match_closed_form(nm):
    el = po.Elem(po.HEX, nm, nm + 1)
    z = el.Z[0]
    A, dA = modified_a(nm, z)
    assert max(rel_errs(el.bdata[0].reshape(nm, -1), A)) < 1e-13
    # dbdata is the COLLOCATION derivative D b (Basis.cpp:417-420): exact for these polynomials up to rounding
    assert max(rel_errs(el.dbdata[0].reshape(nm, -1), dA)) < 1e-11


@pytest.mark.parametrize("nm", [2, 3, 5, 8])
def test_modified_b_and_c_tables_match_closed_form(nm):
    el = po.Elem(po.TET, nm, nm + 1)          # direction 1: Modified_B on Gauss-Radau (alpha=1), 2: Modified_C (alpha=2)
    B, dB = modified_b(nm, el.Z[1])
    assert B.shape[0] == nm * (nm + 1) // 2
    assert max(rel_errs(el.bdata[1].reshape(B.shape[0], -1), B)) < 1e-13
    assert max(rel_errs(el.dbdata[1].reshape(B.shape[0], -1), dB)) < 1e-10
    # Modified_C(p,q,r) = Modified_B(p+q, r): for every p the B rows from row-block p on (Basis.cpp:548-559)
    B2, dB2 = modified_b(nm, el.Z[2])
    blk = np.cumsum([0] + [nm - p for p in range(nm)])
    Crows = np.concatenate([B2[blk[p]:] for p in range(nm)])
    dCrows = np.concatenate([dB2[blk[p]:] for p in range(nm)])
    assert Crows.shape[0] == nm * (nm + 1) * (nm + 2) // 6
    assert max(rel_errs(el.bdata[2].reshape(Crows.shape[0], -1), Crows)) < 1e-13
    assert max(rel_errs(el.dbdata[2].reshape(Crows.shape[0], -1), dCrows)) < 1e-10


def _tensor_modes(dim, nm, nq0):
    """phi[m, q], dphi[d][m, q] of the tensor-product expansion at the tensor quadrature grid; mode index
    m = p + nm (q + nm r), point index likewise (first direction fastest) -- StdHexExp / StdQuadExp ordering"""
    el = po.Elem(po.HEX if dim == 3 else po.QUAD, nm, nq0)
    z, w = el.Z[0], el.W[0]
    A, dA = modified_a(nm, z)
    if dim == 2:
        phi = np.einsum("qj,pi->qpji", A, A).reshape(nm * nm, -1)
        d0 = np.einsum("qj,pi->qpji", A, dA).reshape(nm * nm, -1)
        d1 = np.einsum("qj,pi->qpji", dA, A).reshape(nm * nm, -1)
        W = np.einsum("j,i->ji", w, w).reshape(-1)
        return el, phi, [d0, d1], W
    phi = np.einsum("rk,qj,pi->rqpkji", A, A, A).reshape(nm ** 3, -1)
    d0 = np.einsum("rk,qj,pi->rqpkji", A, A, dA).reshape(nm ** 3, -1)
    d1 = np.einsum("rk,qj,pi->rqpkji", A, dA, A).reshape(nm ** 3, -1)
    d2 = np.einsum("rk,qj,pi->rqpkji", dA, A, A).reshape(nm ** 3, -1)
    W = np.einsum("k,j,i->kji", w, w, w).reshape(-1)
    return el, phi, [d0, d1, d2], W


@pytest.mark.parametrize("dim,nm,nq0,deformed", [(3, 3, 4, False), (3, 5, 6, False), (3, 4, 5, True), (3, 4, 6, True),
                                                 (2, 6, 7, False), (2, 5, 6, True), (2, 4, 7, True)])
def test_operators_equal_dense_pointwise_statement(dim, nm, nq0, deformed):
    rng = np.random.default_rng(100 * dim + nm)
    el, phi, dphi, W = _tensor_modes(dim, nm, nq0)
    nel, lam, nM, nQ = 3, 0.7, el.nmTot, el.nqTot
    assert phi.shape == (nM, nQ)
    jac, df = random_geometry(rng, dim, nel, nQ, deformed)
    J = jac.reshape(nel, -1)                                   # [nel, 1 | nQ]
    DF = df.reshape(dim * dim, nel, -1)                        # DF[c*dim+d][e][pt] = d xi_d / d x_c
    c = rng.uniform(-1, 1, (nel, nM))
    f = rng.uniform(-1, 1, (nel, nQ))
    # BwdTrans / IProductWRTBase
    assert max(rel_errs(el.bwdtrans(nel, c.reshape(-1)), (c @ phi).reshape(-1))) < 1e-12
    assert max(rel_errs(el.iproduct(nel, deformed, jac, f.reshape(-1)), ((f * W * J) @ phi.T).reshape(-1))) < 1e-12
    # PhysDeriv of a field that IS in the polynomial space (the collocation derivative is then exact)
    u = c @ phi
    du_xi = [c @ d for d in dphi]
    want = [sum(DF[cc * dim + d] * du_xi[d] for d in range(dim)) for cc in range(dim)]
    got = el.physderiv(nel, deformed, df, u.reshape(-1))
    for g, wv in zip(got, want):
        assert max(rel_errs(g, wv.reshape(-1))) < 1e-11
    # Helmholtz as dense elemental matrices
    want_h = np.zeros((nel, nM))
    for e in range(nel):
        Je = np.broadcast_to(J[e], (nQ,))
        grad = [sum(np.broadcast_to(DF[cc * dim + d][e], (nQ,)) * dphi[d] for d in range(dim)) for cc in range(dim)]
        H = lam * (phi * (W * Je)) @ phi.T
        for g in grad:
            H += (g * (W * Je)) @ g.T
        assert np.abs(H - H.T).max() < 1e-12 * np.abs(H).max()
        want_h[e] = H @ c[e]
    got_h = el.helmholtz(nel, deformed, jac, df, lam, c.reshape(-1))
    assert max(rel_errs(got_h, want_h.reshape(-1))) < 1e-12


@pytest.mark.parametrize("nm,deformed", [(3, False), (5, False), (6, True), (4, True)])
def test_triangle_operators_equal_dense_collapsed_statement(nm, deformed):
    """Collapsed-coordinate triangle (StdRegions/StdTriExp.cpp): phi_pq(xi) = A_p(eta1) B_pq(eta2) on the Duffy grid
    eta1 = 2 (1+xi1)/(1-xi2) - 1, eta2 = xi2, mode (p,q) packed p-major / q fastest, the collapsed top vertex
    (packed mode 1) taking its full shape B_01(eta2) [the A_0 + A_1 = 1 correction of the sum-factorised kernels];
        d/dxi1 = 2/(1-eta2) d/deta1,   d/dxi2 = (1+eta1)/(1-eta2) d/deta1 + d/deta2,
    quadrature = GLL(eta1) x Gauss-Radau(1,0)(eta2) with the collapse Jacobian folded in as w2/2."""
    rng = np.random.default_rng(7 + nm)
    el = po.Elem(po.TRI, nm, nm + 1)
    z1, w1, z2, w2 = el.Z[0], el.W[0], el.Z[1], el.W[1]
    A, dA = modified_a(nm, z1)
    B, dB = modified_b(nm, z2)
    nM, nQ = el.nmTot, el.nqTot
    phi, d1, d2 = np.zeros((nM, nQ)), np.zeros((nM, nQ)), np.zeros((nM, nQ))
    m = 0
    for p in range(nm):
        for q in range(nm - p):
            a, da = (np.ones_like(z1), np.zeros_like(z1)) if m == 1 else (A[p], dA[p])
            phi[m] = np.einsum("j,i->ji", B[m], a).reshape(-1)
            de1 = np.einsum("j,i->ji", B[m], da)                    # d/deta1
            de2 = np.einsum("j,i->ji", dB[m], a)                    # d/deta2
            d1[m] = (2.0 / (1.0 - z2))[:, None].repeat(z1.size, 1).reshape(-1) * de1.reshape(-1)
            d2[m] = (((1.0 + z1)[None, :] / (1.0 - z2)[:, None]) * de1 + de2).reshape(-1)
            m += 1
    assert m == nM
    W = np.einsum("j,i->ji", 0.5 * w2, w1).reshape(-1)
    nel, lam = 3, 1.1
    jac, df = random_geometry(rng, 2, nel, nQ, deformed)
    J, DF = jac.reshape(nel, -1), df.reshape(4, nel, -1)
    c = rng.uniform(-1, 1, (nel, nM))
    f = rng.uniform(-1, 1, (nel, nQ))
    assert max(rel_errs(el.bwdtrans(nel, c.reshape(-1)), (c @ phi).reshape(-1))) < 1e-12
    assert max(rel_errs(el.iproduct(nel, deformed, jac, f.reshape(-1)), ((f * W * J) @ phi.T).reshape(-1))) < 1e-12
    want_h = np.zeros((nel, nM))
    dxi = [d1, d2]
    for e in range(nel):
        Je = np.broadcast_to(J[e], (nQ,))
        H = lam * (phi * (W * Je)) @ phi.T
        for cc in range(2):
            g = sum(np.broadcast_to(DF[cc * 2 + d][e], (nQ,)) * dxi[d] for d in range(2))
            H += (g * (W * Je)) @ g.T
        want_h[e] = H @ c[e]
    assert max(rel_errs(el.helmholtz(nel, deformed, jac, df, lam, c.reshape(-1)), want_h.reshape(-1))) < 1e-11
