"""Sharded device CG / HelmSolve chain against the serial CPU oracle (test infrastructure; launched under torchrun by
tests/test_gpu_cg.py::test_sharded_cg_two_gpus, and called by bench.py for its `cg_parity` field).  Builds the
problem with tools/bench_cg.setup() on this rank's z-slab (or box of a --part px,py,pz partition) and compares, on the FULL mesh, with the oracle:

  * one mat-vec s = Assemble(Helmholtz(GlobalToLocal(w))) + interface exchange on a random global vector
    (well-posed: relative Linf <= 1e-12, the north-star tolerance);
  * the converged CG solution (tol 1e-12; two converged solves agree to ~cond * tol, checked at 1e-9 -- a fixed
    iteration count is NOT a usable check, the single-reduction recurrence amplifies 1e-16 perturbations of the
    right-hand side to O(10 %) in r.r after ~30 iterations);
  * the whole ContField::v_HelmSolve + BwdTrans chain through nekmf_helmsolve with host arrays."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench_cg  # noqa: E402
import pyoracle as po  # noqa: E402
import _sharded_ref as sr  # noqa: E402

TOL_SOLVE = 1e-12


def run_check(S, a):
    """-> dict with the three errors, iteration counts and `ok` (this rank's verdict; the caller reduces over ranks)"""
    rank, dev, mesh, mesh_mod, std, lam, cg, rhs, x, nk = (S[k] for k in (
        "rank", "dev", "mesh", "mesh_mod", "std", "lam", "cg", "rhs", "x", "nk"))
    full = mesh_mod.StructuredHexMesh(a.nx, a.ny, a.nz, a.nm)
    el = po.Elem(po.HEX, a.nm, a.nm + 1)
    jf, dff = full.geometry()
    mine_ids = mesh.lattice_ids                                  # [Gzl, Gy, Gx] rank-local global ids
    # the same lattice points in the numbering of the unpartitioned mesh
    full_ids = full.lattice_ids[mesh.gz0:mesh.gz1 + 1, mesh.gy0:mesh.gy1 + 1, mesh.gx0:mesh.gx1 + 1]
    # ---- mat-vec
    wf = np.random.default_rng(17).uniform(-1, 1, full.nGlobal)
    want_s = po.assemble(full.localToGlobal, None, el.helmholtz(full.nElmt, False, jf, dff, lam, po.global_to_local(
        full.localToGlobal, None, wf)), full.nGlobal)
    w_slab = np.empty(mesh.nGlobal)
    w_slab[mine_ids.reshape(-1)] = wf[full_ids.reshape(-1)]
    w_d = torch.tensor(w_slab, device=dev)
    s_d = torch.zeros(mesh.nGlobal, dtype=torch.float64, device=dev)
    cg.matvec(w_d, s_d)
    torch.cuda.synchronize()
    err_mv = np.abs(s_d.cpu().numpy()[mine_ids] - want_s[full_ids]).max() / np.abs(want_s).max()
    # ---- converged solve
    its, eps = cg.solve(rhs, x, tol=TOL_SOLVE, maxiter=5000)
    rhs_o, _ = sr.helmholtz_rhs(None, full, el, jf, lam)
    dg = full.helmholtz_diagonal(std.basis[0], lam)
    xo, itso, _ = el.cg(full.nElmt, False, jf, dff, lam, full.nGlobal, full.nDir, full.localToGlobal, None,
                        1.0 / dg[full.nDir:], rhs_o, tol=TOL_SOLVE)
    err_x = np.abs(x.cpu().numpy()[mine_ids] - xo[full_ids]).max() / np.abs(xo).max()
    # ---- the whole ContField::v_HelmSolve chain (nekmf_helmsolve), sharded: host arrays in / out
    hs = nk.HelmSolver(cg, S["ipr"], S["bwd"])
    f_slab = -(lam + 3 * np.pi ** 2) * S["u_exact"]
    coef, phys = np.zeros(mesh.nLocal), np.zeros(f_slab.size)
    its_h, _ = hs.HelmSolve(f_slab, coef, phys, tol=TOL_SOLVE)
    X, Y, Z = full.quad_coords(el.Z[0])
    f_full = -(lam + 3 * np.pi ** 2) * np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z)
    ch = po.Chain(el, full.nElmt, False, jf, dff, lam, full.localToGlobal, None, full.nGlobal, full.nDir,
                  1.0 / dg[full.nDir:])
    want_c, want_p = np.zeros(full.nLocal), np.zeros(f_full.size)
    its_ho, _ = ch.helmsolve(f_full, want_c, want_p, tol=TOL_SOLVE)
    mine_c = want_c.reshape(full.nElmt, el.nmTot)[mesh.element_ids].reshape(-1)
    mine_p = want_p.reshape(full.nElmt, el.nqTot)[mesh.element_ids].reshape(-1)
    err_c = np.abs(coef - mine_c).max() / np.abs(want_c).max()
    err_p = np.abs(phys - mine_p).max() / np.abs(want_p).max()
    del hs
    # iteration counts: a sanity bound only -- at tol = 1e-12 the residual sits on its round-off plateau and the
    # crossing iteration moves by 10-15 % with the summation order of the dot products
    ok = (err_mv < 1e-12 and err_x < 1e-9 and err_c < 1e-9 and err_p < 1e-9 and
          abs(its - itso) <= max(3, 0.25 * itso) and abs(its_h - its_ho) <= max(3, 0.25 * its_ho))
    comm = S["keepalive"][2]
    return {"ok": bool(ok), "matvec_rel_linf": float(err_mv), "solution_rel_linf": float(err_x),
            "chain_coeff_rel_linf": float(err_c), "chain_phys_rel_linf": float(err_p),
            "iterations": [int(its), int(itso)], "chain_iterations": [int(its_h), int(its_ho)],
            "tol": TOL_SOLVE, "mesh": [a.nx, a.ny, a.nz], "ranks": S["world"], "partition": list(mesh.part),
            "neighbours_of_this_rank": len(mesh.peers),
            "transport": comm.transport if comm is not None else "none",
            "oracle": "oracle/libmforacle.so (mfo_cg_helmholtz, mfo_chain_helmsolve) on the unpartitioned mesh"}


def reduce_ok(S, res):
    """worst case over the ranks: ok = all, errors = max"""
    dist = S["dist"]
    keys = ("matvec_rel_linf", "solution_rel_linf", "chain_coeff_rel_linf", "chain_phys_rel_linf")
    t = torch.tensor([0.0 if res["ok"] else 1.0] + [res[k] for k in keys], dtype=torch.float64, device=S["dev"])
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = dict(res)
    res["ok"] = float(t[0].item()) == 0.0
    for i, k in enumerate(keys):
        res[k] = float(t[1 + i].item())
    return res


def main():
    a = bench_cg.parse_args()
    S = bench_cg.setup(a)
    res = reduce_ok(S, run_check(S, a))
    if S["rank"] == 0:
        print(res)
        print("CHECK OK" if res["ok"] else "CHECK FAILED")
    if S["dist"] is not None:
        S["dist"].destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
