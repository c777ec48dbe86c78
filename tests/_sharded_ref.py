"""Host-side reference of the SHARDED conjugate-gradient algorithm (test infrastructure).

Same sequence of operations as ithaca-sem_b200/csrc/cg.cu and comm.cu -- local gather, elemental
Helmholtz, local assemble, pairwise interface exchange-add, ownership-masked dot products, one
3-value all-reduce per iteration -- but with the CPU oracle as the elemental operator and
torch.distributed (gloo) as the transport, so the partition / interface / mask logic of
ithaca-sem_b200/mesh.py can be validated with world_size > 1 on a machine without GPUs.
"""
import numpy as np

import pyoracle as po


def exchange_add(dist, glob, peers, lists):
    import torch
    if not peers:
        return
    send = [torch.from_numpy(np.ascontiguousarray(glob[l])) for l in lists]
    recv = [torch.empty_like(s) for s in send]
    ops = []
    for p, s, r in zip(peers, send, recv):
        ops.append(dist.P2POp(dist.isend, s, p))
        ops.append(dist.P2POp(dist.irecv, r, p))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for l, r in zip(lists, recv):
        np.add.at(glob, l, r.numpy())


def allreduce(dist, vals):
    import torch
    if dist is None:
        return vals
    t = torch.tensor(vals, dtype=torch.float64)
    dist.all_reduce(t)
    return t.tolist()


def sharded_cg(dist, mesh, el, jac, df, lam, rhs, invdiag=None, tol=1e-9, maxiter=5000):
    """returns (x, iterations, eps) for this rank's slab; dist may be None for a single rank."""
    nD, nG = mesh.nDir, mesh.nGlobal
    l2g, mask = mesh.localToGlobal, mesh.ownerMask
    mnd = mask[nD:]

    def matvec(w):
        loc = po.global_to_local(l2g, None, w)
        out = el.helmholtz(mesh.nElmt, False, jac, df, lam, loc)
        s = po.assemble(l2g, None, out, nG)
        if dist is not None:
            exchange_add(dist, s, mesh.peers, mesh.interface_lists)
        return s

    def precon(r):
        return r * invdiag if invdiag is not None else r.copy()

    x = np.zeros(nG)
    r = rhs[nD:].copy()
    w, p, q = np.zeros(nG), np.zeros(nG - nD), np.zeros(nG - nD)
    eps = allreduce(dist, [float(np.dot(r * mnd, r))])[0]
    rhs_mag = allreduce(dist, [float(np.dot(rhs * mask, rhs))])[0]
    rhs_mag = rhs_mag if rhs_mag > 1e-6 else 1.0
    if eps < tol * tol * rhs_mag:
        return x, 0, eps
    w[nD:] = precon(r)
    s = matvec(w)
    rho, mu = allreduce(dist, [float(np.dot(r * mnd, w[nD:])), float(np.dot(s[nD:] * mnd, w[nD:]))])
    beta, alpha, its, k = 0.0, rho / mu, 1, 0
    while k < maxiter:
        p = beta * p + w[nD:]
        q = beta * q + s[nD:]
        x[nD:] += alpha * p
        r -= alpha * q
        w[nD:] = precon(r)
        s = matvec(w)
        rho_new, mu, eps = allreduce(dist, [float(np.dot(r * mnd, w[nD:])), float(np.dot(s[nD:] * mnd, w[nD:])),
                                            float(np.dot(r * mnd, r))])
        its += 1
        if eps < tol * tol * rhs_mag:
            break
        beta = rho_new / rho
        alpha = rho_new / (mu - rho_new * beta / alpha)
        rho = rho_new
        k += 1
    return x, its, eps


def helmholtz_rhs(dist, mesh, el, jac, lam):
    """assembled right-hand side of  lap(u) - lam u = f  with u = sin(pi x) sin(pi y) sin(pi z):
    rhs = -IProductWRTBase(f) (ContField::v_HelmSolve, ContField.cpp:894-897), f = -(lam + 3 pi^2) u."""
    X, Y, Z = mesh.quad_coords(el.Z[0])
    u = np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z)
    f = -(lam + 3 * np.pi ** 2) * u
    loc = -el.iproduct(mesh.nElmt, False, jac, f)
    rhs = po.assemble(mesh.localToGlobal, None, loc, mesh.nGlobal)
    if dist is not None:
        exchange_add(dist, rhs, mesh.peers, mesh.interface_lists)
    rhs[:mesh.nDir] = 0.0
    return rhs, u
