#!/usr/bin/env python
"""Matrix-free CG Helmholtz solve on a structured hex mesh at P=4, element-partitioned in z-slabs
over the ranks of one box (BASELINE.json configs[4]).

    python tools/bench_cg.py [--nx 64 --ny 128 --nz 128] [--iters 50]
    torchrun --nproc-per-node N tools/bench_cg.py ...

Every rank builds its slab (ithaca-sem_b200/mesh.py), creates the device operator / assembly map /
NCCL exchange through the C ABI and runs nekmf_cg_solve with a fixed iteration cap.  The convergence /
parity check against the serial CPU oracle lives in tests/_cg_check.py (it reuses setup() from here)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from _util import load_pkg_module, nekmf  # noqa: E402


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=64)
    ap.add_argument("--ny", type=int, default=128)
    ap.add_argument("--nz", type=int, default=128)
    ap.add_argument("--nm", type=int, default=5)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--part", default=None, help="px,py,pz box partition over the ranks (default: z-slabs)")
    return ap.parse_args(argv)


def setup(a, comm=None):
    """builds this rank's slab, operators, assembly map, exchange, right-hand side and the CG object.  The
    process group is initialised here unless the caller already did; `comm` reuses an existing nekmf communicator."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nk = nekmf()
    mesh_mod = load_pkg_module("mesh")
    dist = None
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=dev)
        if comm is None:
            comm = nk.Comm.from_torch_distributed()
    lam = 1.0
    if getattr(a, "part", None):
        part = tuple(int(v) for v in a.part.split(","))
        if part[0] * part[1] * part[2] != world:
            raise SystemExit("--part %s does not match %d ranks" % (a.part, world))
        mesh = mesh_mod.StructuredHexMesh(a.nx, a.ny, a.nz, a.nm, part=part, rank=rank)
    else:
        mesh = mesh_mod.StructuredHexMesh(a.nx, a.ny, a.nz, a.nm, slab=(rank, world))
    std = nk.StdExpansion(nk.eHexahedron, a.nm)
    jac, df = mesh.geometry()
    geom = nk.CoalescedGeomData(jac, df, False)
    helm = nk.Operator(std, mesh.nElmt, geom, nk.eHelmholtz)
    helm.SetLambda(lam)
    ipr = nk.Operator(std, mesh.nElmt, geom, nk.eIProductWRTBase)
    bwd = nk.Operator(std, mesh.nElmt, geom, nk.eBwdTrans)
    amap = nk.AssemblyMap(mesh.localToGlobal, mesh.nGlobal)
    ex = nk.Exchange(comm, mesh.peers, mesh.interface_lists, mesh.nGlobal) if world > 1 else None
    # right-hand side  -(v, f),  f = -(lam + 3 pi^2) sin sin sin  (device IProduct + Assemble + exchange)
    z = std.basis[0].Z
    X, Y, Z = mesh.quad_coords(z)
    u_exact = np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z)
    f = torch.tensor((lam + 3 * np.pi ** 2) * u_exact, device=dev)
    del X, Y, Z
    loc = torch.empty(mesh.nLocal, dtype=torch.float64, device=dev)
    ipr.apply([f], [loc])
    rhs = torch.empty(mesh.nGlobal, dtype=torch.float64, device=dev)
    amap.Assemble(loc, rhs)
    if ex is not None:
        ex.add(rhs)
    rhs[:mesh.nDir] = 0.0
    torch.cuda.synchronize()
    # Jacobi preconditioner: elemental diagonals from the operator itself, assembled / exchanged / inverted on the
    # device (nekmf_cg_set_jacobi, the PreconditionerDiagonal replacement)
    cg = nk.HelmholtzCG(helm, amap, mesh.nDir, None, exchange=ex, comm=comm, ownerMask=mesh.ownerMask)
    cg.set_jacobi()
    x = torch.zeros(mesh.nGlobal, dtype=torch.float64, device=dev)
    del f, loc
    return dict(rank=rank, world=world, dev=dev, nk=nk, mesh_mod=mesh_mod, dist=dist, mesh=mesh, std=std, lam=lam, helm=helm,
                cg=cg, rhs=rhs, x=x, ipr=ipr, bwd=bwd, u_exact=u_exact, keepalive=(amap, ex, comm, geom, ipr, bwd))


def measure(S, a):
    """fixed-iteration timing of the sharded solve: -> dict (identical on every rank)"""
    rank, world, dev, nk, dist, mesh, helm, cg, rhs, x = (S[k] for k in ("rank", "world", "dev", "nk", "dist", "mesh", "helm", "cg", "rhs", "x"))
    # ---- timing: fixed iteration cap (tolerance 0 never triggers), max over ranks
    cg.solve(rhs, x, tol=0.0, maxiter=8, raise_on_maxiter=False)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = nk.launch_count()
    t0 = time.perf_counter()
    its_total, eps = cg.solve(rhs, x, tol=0.0, maxiter=a.iters, raise_on_maxiter=False)
    torch.cuda.synchronize()
    dt_wall = time.perf_counter() - t0
    loop_ms, its = cg.last_loop()  # CUDA events on the solver's stream around the iteration loop
    t = torch.tensor([dt_wall, loop_ms * 1e-3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt_wall, dt = float(t[0].item()), float(t[1].item())
    nel_total = a.nx * a.ny * a.nz
    gdof = (a.nx * (a.nm - 1) + 1) * (a.ny * (a.nm - 1) + 1) * (a.nz * (a.nm - 1) + 1)
    # per-rank bytes one iteration must move with the kernels as fused: update 7 reads + 5 writes
    # of nNonDir doubles; mat-vec: map (4 B) + local output written and read back (16 B) + transpose-CSR column
    # (4 B) per local DOF, gathered w + rowptr + s written + w re-read for s.w per global DOF
    nN = mesh.nGlobal - mesh.nDir
    by = 12 * 8 * nN + mesh.nLocal * (4 + 16 + 4) + mesh.nGlobal * (8 + 4 + 8 + 8) + mesh.nElmt * 32
    sys.path.insert(0, ROOT)
    import bench
    peak, peak_src = bench.measured_peaks()
    gbs = by * its / dt / 1e9
    return {"metric": "matrix-free CG Helmholtz, hex P=%d" % (a.nm - 1), "n_gpus": world,
            "kernel": helm.kernel_name,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                         "bytes_per_iteration_per_rank": by, "peak_source": peak_src},
            "elements": nel_total, "global_dof": gdof, "iterations": its,
            "ms_per_iteration": dt / its * 1e3, "ms_per_iteration_wall": dt_wall / its_total * 1e3,
            "transport": S["keepalive"][2].transport if world > 1 else "none",
            "gdof_per_s_local": nel_total * a.nm ** 3 * its / dt / 1e9,
            "launches": nk.launch_count() - l0, "final_eps": eps}


def main():
    a = parse_args()
    S = setup(a)
    res = measure(S, a)
    if S["rank"] == 0:
        print(json.dumps(res))
    if S["dist"] is not None:
        S["dist"].destroy_process_group()


if __name__ == "__main__":
    main()
