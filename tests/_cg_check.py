"""Sharded device CG against the serial CPU oracle (test infrastructure; launched under torchrun by
tests/test_gpu_cg.py::test_sharded_cg_two_gpus).  Builds the problem with tools/bench_cg.setup(), solves to
convergence on the device(s) and compares with NekLinSysIterCG's restatement in oracle/ on the full mesh."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench_cg  # noqa: E402
import pyoracle as po  # noqa: E402
import _sharded_ref as sr  # noqa: E402


def main():
    a = bench_cg.parse_args()
    S = bench_cg.setup(a)
    rank, dev, dist, mesh, mesh_mod, std, lam, cg, rhs, x = (S[k] for k in (
        "rank", "dev", "dist", "mesh", "mesh_mod", "std", "lam", "cg", "rhs", "x"))
    its, eps = cg.solve(rhs, x, tol=1e-13, maxiter=5000)
    full = mesh_mod.StructuredHexMesh(a.nx, a.ny, a.nz, a.nm)
    el = po.Elem(po.HEX, a.nm, a.nm + 1)
    jf, dff = full.geometry()
    rhs_o, _ = sr.helmholtz_rhs(None, full, el, jf, lam)
    dg = full.helmholtz_diagonal(std.basis[0], lam)
    xo, itso, _ = el.cg(full.nElmt, False, jf, dff, lam, full.nGlobal, full.nDir, full.localToGlobal, None,
                        1.0 / dg[full.nDir:], rhs_o, tol=1e-13)
    mine = x.cpu().numpy()[mesh.lattice_ids]
    want = xo[full.lattice_ids][mesh.gz0:mesh.gz1 + 1]
    err = np.abs(mine - want).max() / np.abs(xo).max()
    # tol=1e-13 is at the round-off plateau: the iteration at which r.r crosses it depends on the
    # summation order of the (ownership-masked, all-reduced) dot products, so allow a few percent
    ok = err < 1e-10 and abs(its - itso) <= max(2, 0.05 * itso)
    t = torch.tensor([0.0 if ok else 1.0], device=dev)
    if dist is not None:
        dist.all_reduce(t)
    if rank == 0:
        comm = S["keepalive"][2]
        print("rank0 its=%d (oracle %d) err=%.2e transport=%s ranks=%d" % (
            its, itso, err, comm.transport if comm is not None else "none", S["world"]))
        print("CHECK OK" if float(t.item()) == 0.0 else "CHECK FAILED")
    if dist is not None:
        dist.destroy_process_group()
    sys.exit(0 if float(t.item()) == 0.0 else 1)


if __name__ == "__main__":
    main()
