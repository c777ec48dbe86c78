"""bench.py --config 1 | 4: the BASELINE.json configurations that are neither the default (configs[1], bench.py
itself), the per-order sweep (configs[2], tools/sweep.py) nor the sharded CG (configs[4], tools/bench_cg.py).

  config 1  2-D Helmholtz on a structured N x N quad mesh, P=5 (nm=6, nq=7), lambda=1 -- the reference's
            Helmholtz2D_modal session scaled up: operator apply (device-resident and host arrays) and the
            device CG iteration on the assembled problem.
  config 4  Tet and Prism (plus the hexahedra they are mixed with) Helmholtz apply at P=6 on a synthetic mixed
            mesh: an n^3 structured grid of cubes, 60 % kept as hexahedra, 30 % cut into 2 prisms, 10 % into 6
            tetrahedra, one collection per shape as CreateCollections forms them (MultiRegions/ExpList.cpp:
            5005-5151).  Headline: the box-cut mesh (axis-aligned affine elements, extruded prisms -- the best
            case of three special-case kernels); `geometry_variants` repeats the step on the same mesh under a
            global rotation + shear (general affine elements) and under a smooth warp (deformed elements).

Every line carries `roofline` (dominant kernel, algorithmic bytes of SURVEY.md 8(d)) and `cpu_baseline` (the
reference's own kernels from oracle/_ref on a bounded sample, all host threads).  Imported by bench.py only."""
import itertools
import json
import os
import time

import numpy as np

LAMBDA = 1.0


def _time_apply(torch, op, ins, outs, reps, warm=3):
    for _ in range(warm):
        op.apply(ins, outs)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        op.apply(ins, outs)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2], ev[0].elapsed_time(ev[reps]) / reps


def _cpu_helmholtz(po, shape, nm, nq0, nel, jac, df, seconds=4.0):
    """reference MatrixFree Helmholtz (AVX2 build) on `nel` affine elements, all host threads; GDOF/s"""
    el = po.Elem(shape, nm, nq0)
    try:
        ref = po.Ref("avx2")
        kind = "reference"
    except Exception:
        ref, kind = None, "port"
    threads = os.cpu_count() or 1
    x = np.random.default_rng(1234).uniform(-1, 1, nel * el.nmTot)
    if ref is not None:
        op = ref.operator(po.OP_HELM, el, nel, False, jac, df)
        out = [np.zeros(nel * el.nmTot)]
        run = lambda: op(x, lam=LAMBDA, nthreads=threads, outs=out)
    else:
        po.set_threads(threads)
        run = lambda: el.helmholtz(nel, False, jac, df, LAMBDA, x)
    run()
    t0 = time.perf_counter()
    run()
    t1 = time.perf_counter() - t0
    reps = max(3, min(100, int(seconds / max(t1, 1e-4))))
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return nel * el.nmTot / ts[len(ts) // 2] / 1e9, ts[len(ts) // 2], kind, threads, reps


def config1(args, torch, nk, po, peak, peak_src, ClockSampler):
    from _util import load_pkg_module
    mesh_mod = load_pkg_module("mesh")
    dev = torch.device("cuda", 0)
    nm, nq, N = 6, 7, args.nx if args.nx != 64 else 1024
    mesh = mesh_mod.StructuredQuadMesh(N, N, nm)
    std = nk.StdExpansion(nk.eQuadrilateral, nm)
    jac, df = mesh.geometry()
    geom = nk.CoalescedGeomData(jac, df, False)
    helm = nk.Operator(std, mesh.nElmt, geom, nk.eHelmholtz)
    helm.SetLambda(LAMBDA)
    ndof = mesh.nLocal
    gen = torch.Generator(device="cpu").manual_seed(1234)
    x_host = (torch.rand(ndof, dtype=torch.float64, generator=gen) * 2 - 1).pin_memory()
    y_host = torch.empty(ndof, dtype=torch.float64).pin_memory()
    x, y = x_host.to(dev), torch.empty(ndof, dtype=torch.float64, device=dev)
    sampler = ClockSampler(0)
    sampler.start()
    med, avg = _time_apply(torch, helm, [x], [y], args.steps, max(args.warmup, 3))
    if len(sampler.samples) < 8:
        t_end = time.perf_counter() + 0.4
        while time.perf_counter() < t_end:
            for _ in range(20):
                helm.apply([x], [y])
            torch.cuda.synchronize()
    clocks = sampler.stop()
    for _ in range(2):
        helm.apply([x_host], [y_host])
    t0 = time.perf_counter()
    Ke = 5
    for _ in range(Ke):
        helm.apply([x_host], [y_host])
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / Ke
    # device CG on the assembled problem (Jacobi preconditioned), fixed iteration count
    amap = nk.AssemblyMap(mesh.localToGlobal, mesh.nGlobal)
    diag = mesh.helmholtz_diagonal(std.basis[0], LAMBDA)
    cg = nk.HelmholtzCG(helm, amap, mesh.nDir, 1.0 / diag[mesh.nDir:])
    rhs = torch.rand(mesh.nGlobal, dtype=torch.float64, device=dev)
    rhs[:mesh.nDir] = 0.0
    xs = torch.zeros(mesh.nGlobal, dtype=torch.float64, device=dev)
    cg.solve(rhs, xs, tol=0.0, maxiter=3, raise_on_maxiter=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    its, _ = cg.solve(rhs, xs, tol=0.0, maxiter=30, raise_on_maxiter=False)
    torch.cuda.synchronize()
    cg_ms = (time.perf_counter() - t0) / its * 1e3
    bytes_el = 8 * (2 * nm * nm + 5)  # 616 B per quad at P=5, regular geometry
    ach = bytes_el * mesh.nElmt / (avg * 1e-3) / 1e9
    ns = 65536
    cpu, cpu_t, kind, threads, reps = _cpu_helmholtz(po, po.QUAD, nm, nq, ns, jac[:ns].copy(),
                                                     df.reshape(4, -1)[:, :ns].reshape(-1).copy())
    return {
        "metric": "GDOF/s FP64 Helmholtz apply (quad P=5)", "value": ndof / (avg * 1e-3) / 1e9, "unit": "GDOF/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": avg, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "2D Helmholtz, structured %dx%d quad mesh, P=5 (nm=6, nq=7), lambda=1 "
                               "(BASELINE configs[0], Helmholtz2D_modal scaled up)" % (N, N),
                   "elements": mesh.nElmt, "l2": "inputs larger than L2 (in+out %d MB), no flush" % (ndof * 16 >> 20)},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_element": bytes_el, "kernel": helm.kernel_name,
                     "kernel_ms": avg},
        "e2e": {"value": ndof / e2e_s / 1e9, "unit": "GDOF/s", "h2d_bytes_per_step": ndof * 8,
                "d2h_bytes_per_step": ndof * 8, "ms_per_step": e2e_s * 1e3},
        "gpu_launches": args.steps, "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")},
        "cg": {"ms_per_iteration": cg_ms, "iterations": its, "global_dof": mesh.nGlobal,
               "note": "Jacobi-preconditioned device CG, gather -> quad Helmholtz -> assemble per iteration"},
        "cpu_baseline": {"value": cpu, "unit": "GDOF/s", "cores": threads, "kind": kind,
                         "sample": "%d of %d quads, median of %d applies, reference MatrixFreeOps kernels "
                                   "(oracle/_ref AVX2 build)" % (ns, mesh.nElmt, reps)},
    }


def affine_factors(edges):
    """jac, df[9] of the affine map x = x0 + sum_d (xi_d + 1)/2 * edges[d]  (GeomFactors.cpp:399-474:
    df[c*3+d] = d xi_d / d x_c, jac = det(dx/dxi))"""
    J = np.stack(edges, axis=1) / 2.0          # dx_c / dxi_d
    Ji = np.linalg.inv(J)                       # dxi_d / dx_c  at [d][c]
    df = np.array([Ji[d, c] for c in range(3) for d in range(3)])
    return float(np.linalg.det(J)), df


# a fixed global linear map (rotation about z by 0.3 rad, then a shear): turns the box-cut mesh into GENERAL affine
# elements -- parallelepipeds with a full Laplacian metric, non-extruded prisms, skewed tetrahedra
_c, _s = np.cos(0.3), np.sin(0.3)
GENERAL_MAP = np.array([[1.0, 0.25, 0.10], [0.0, 1.0, 0.20], [0.15, 0.0, 1.0]]) @ np.array([[_c, -_s, 0.0], [_s, _c, 0.0], [0.0, 0.0, 1.0]])


def element_frames(n, h, A=None):
    """origin x0 [nel,3] and edge vectors E [nel,3(d),3(c)] of every element of the synthetic mixed mesh, per shape:
    x = x0 + sum_d (xi_d + 1)/2 E_d.  Cube c (lexicographic) keeps its type by c mod 10: 0-5 hex, 6-8 two prisms,
    9 six tetrahedra (Kuhn simplices).  A: optional global linear map applied to the whole mesh."""
    c = np.arange(n ** 3)
    corner = np.stack([c % n, (c // n) % n, c // (n * n)], axis=1) * h
    kinds = c % 10
    I3 = np.eye(3) * h
    ex, ey, ez = I3
    out = {}
    hx = corner[kinds < 6]
    out["Hex"] = (hx, np.broadcast_to(np.stack([ex, ey, ez]), (hx.shape[0], 3, 3)).copy())
    pc = corner[(kinds >= 6) & (kinds < 9)]
    # prism reference: triangle in (xi_0, xi_2) extruded along xi_1; the second prism is the point-reflected half
    px0 = np.stack([pc, pc + ex + ez], axis=1).reshape(-1, 3)
    pE = np.broadcast_to(np.stack([np.stack([ex, ey, ez]), np.stack([-ex, ey, -ez])]), (pc.shape[0], 2, 3, 3)).reshape(-1, 3, 3).copy()
    out["Prism"] = (px0, pE)
    tc = corner[kinds == 9]
    tE = []
    for perm in itertools.permutations(range(3)):
        a_, b_, c_ = I3[perm[0]], I3[perm[1]], I3[perm[2]]
        E = [a_, a_ + b_, a_ + b_ + c_]          # Kuhn simplex of this permutation: vertices 0, E0, E1, E2
        if np.linalg.det(np.stack(E, axis=1)) < 0:
            E = [E[1], E[0], E[2]]
        tE.append(np.stack(E))
    out["Tet"] = (np.repeat(tc, 6, axis=0), np.broadcast_to(np.stack(tE), (tc.shape[0], 6, 3, 3)).reshape(-1, 3, 3).copy())
    if A is not None:
        out = {k: (x0 @ A.T, E @ A.T) for k, (x0, E) in out.items()}
    return out


def mixed_mesh(n, h, A=None):
    """regular (affine) geometric factors of the mixed mesh: {shape: (nElmt, jac[nElmt], df[9*nElmt])}
    (GeomFactors.cpp:399-474: df[c*3+d] = d xi_d / d x_c, jac = det(dx/dxi))"""
    out = {}
    for name, (x0, E) in element_frames(n, h, A).items():
        J = np.transpose(E, (0, 2, 1)) / 2.0        # J[e][c][d] = dx_c / dxi_d
        Ji = np.linalg.inv(J)                        # Ji[e][d][c] = dxi_d / dx_c
        df = np.stack([Ji[:, d, c_] for c_ in range(3) for d in range(3)])
        out[name] = (x0.shape[0], np.linalg.det(J), np.ascontiguousarray(df).reshape(-1))
    return out


def mixed_mesh_deformed(n, h, std_of, torch, dev, A=None, amp=0.05):
    """the same mesh warped by X = x + amp sin(pi x) sin(pi y) sin(pi z) (1,1,1): per-quadrature-point factors
    {shape: (nElmt, jac[nElmt*nq], df[9*nElmt*nq])} w.r.t. the Cartesian reference coordinates xi (the collapsed
    shapes' quadrature points are mapped eta -> xi first), computed on the device.  dX/dxi = (I + 1 g^T) J_e with
    g = grad w, inverted with Sherman-Morrison."""
    import math
    out = {}
    for name, (x0, E) in element_frames(n, h, A).items():
        std = std_of[name]
        z = [torch.tensor(std.basis[d].Z, dtype=torch.float64, device=dev) for d in range(3)]
        e0, e1, e2 = torch.meshgrid(z[2], z[1], z[0], indexing="ij")[::-1]   # eta_0 fastest: [k][j][i]
        if name == "Prism":
            xi = [(1 + e0) * (1 - e2) / 2 - 1, e1, e2]
        elif name == "Tet":
            xi = [(1 + e0) * (1 - e1) * (1 - e2) / 4 - 1, (1 + e1) * (1 - e2) / 2 - 1, e2]
        else:
            xi = [e0, e1, e2]
        T = torch.stack([(v.reshape(-1) + 1) / 2 for v in xi], dim=1)        # [nq, 3(d)]
        x0t, Et = torch.tensor(x0, device=dev), torch.tensor(E, device=dev)   # [nel,3], [nel,3(d),3(c)]
        X = x0t[:, None, :] + torch.einsum("qd,edc->eqc", T, Et)            # [nel, nq, 3]
        sn, cs = torch.sin(math.pi * X), math.pi * torch.cos(math.pi * X)
        g = amp * torch.stack([cs[..., 0] * sn[..., 1] * sn[..., 2], sn[..., 0] * cs[..., 1] * sn[..., 2],
                               sn[..., 0] * sn[..., 1] * cs[..., 2]], dim=-1)  # grad w, [nel, nq, 3]
        J = torch.transpose(Et, 1, 2) / 2.0                                   # [nel, c, d]
        Ji = torch.linalg.inv(J)                                              # [nel, d, c]
        sg = 1.0 + g.sum(-1)                                                  # det(I + 1 g^T)
        jac = torch.linalg.det(J)[:, None] * sg
        # inv(dX/dxi) = Ji (I - 1 g^T / (1 + sum g)):  inv[d][c] = Ji[d][c] - (sum_c' Ji[d][c']) g_c / sg
        rs = Ji.sum(-1)                                                       # [nel, d]
        inv = Ji[:, None, :, :] - rs[:, None, :, None] * (g / sg[..., None])[:, :, None, :]   # [nel, nq, d, c]
        df = torch.stack([inv[:, :, d, c_] for c_ in range(3) for d in range(3)])             # [9, nel, nq]
        out[name] = (x0.shape[0], jac.reshape(-1).contiguous(), df.reshape(-1).contiguous())
    return out


SHAPES4 = ("Hex", "Prism", "Tet")


def _mixed_operators(torch, nk, mesh, nm, deformed, gen, dev):
    shp = {"Hex": nk.eHexahedron, "Prism": nk.ePrism, "Tet": nk.eTetrahedron}
    ops, bufs = {}, {}
    for name in SHAPES4:
        nel, jac, df = mesh[name]
        std = nk.StdExpansion(shp[name], nm)
        op = nk.Operator(std, nel, nk.CoalescedGeomData(jac, df, deformed), nk.eHelmholtz)
        op.SetLambda(LAMBDA)
        x = torch.rand(nel * std.GetNcoeffs(), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
        ops[name], bufs[name] = op, (x, torch.empty_like(x), std)
    return ops, bufs


def _mixed_step_ms(torch, ops, bufs, steps):
    """one step = the three collections back to back (what ExpList::GeneralMatrixOp does), timed as a whole"""
    for name in SHAPES4:
        for _ in range(3):
            ops[name].apply([bufs[name][0]], [bufs[name][1]])
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(steps):
        for name in SHAPES4:
            x, y, _ = bufs[name]
            ops[name].apply([x], [y])
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / steps


def _geometry_variant(torch, nk, peak, mesh, nm, deformed, gen, dev, steps, workload):
    """the mixed-mesh step on another geometry: per-shape kernels and times, aggregate GDOF/s, HBM fraction of the step"""
    ops, bufs = _mixed_operators(torch, nk, mesh, nm, deformed, gen, dev)
    per = {}
    for name in SHAPES4:
        x, y, std = bufs[name]
        _, avg = _time_apply(torch, ops[name], [x], [y], max(3, steps // 2))
        nq = std.GetTotPoints()
        by = 8 * (2 * std.GetNcoeffs() + 10 * (nq if deformed else 1))
        nel = mesh[name][0]
        per[name] = {"elements": nel, "ms": avg, "gdof_per_s": nel * std.GetNcoeffs() / (avg * 1e-3) / 1e9,
                     "algorithmic_bytes_per_element": by, "frac_hbm": by * nel / (avg * 1e-3) / 1e9 / peak,
                     "kernel": ops[name].kernel_name}
    step_ms = _mixed_step_ms(torch, ops, bufs, steps)
    ndof = sum(mesh[n_][0] * bufs[n_][2].GetNcoeffs() for n_ in SHAPES4)
    by_step = sum(per[n_]["algorithmic_bytes_per_element"] * per[n_]["elements"] for n_ in SHAPES4)
    return {"workload": workload, "value": ndof / (step_ms * 1e-3) / 1e9, "unit": "GDOF/s", "ms_per_step": step_ms,
            "frac_hbm_of_step": by_step / (step_ms * 1e-3) / 1e9 / peak, "per_shape": per}


def config4(args, torch, nk, po, peak, peak_src, ClockSampler):
    dev = torch.device("cuda", 0)
    nm, n = 7, args.nx if args.nx != 64 else 60
    mesh = mixed_mesh(n, 1.0 / n)
    shapes = {"Hex": (nk.eHexahedron, po.HEX), "Prism": (nk.ePrism, po.PRISM), "Tet": (nk.eTetrahedron, po.TET)}
    per = {}
    gen = torch.Generator(device=dev).manual_seed(1234)
    ops, bufs = _mixed_operators(torch, nk, mesh, nm, False, gen, dev)
    sampler = ClockSampler(0)
    sampler.start()
    for name in shapes:
        x, y, std = bufs[name]
        med, avg = _time_apply(torch, ops[name], [x], [y], args.steps, max(args.warmup, 3))
        nel = mesh[name][0]
        by = 8 * (2 * std.GetNcoeffs() + 10)
        per[name] = {"elements": nel, "ncoeffs": std.GetNcoeffs(), "ms": avg, "gdof_per_s": nel * std.GetNcoeffs() / (avg * 1e-3) / 1e9,
                     "algorithmic_bytes_per_element": by, "gb_per_s": by * nel / (avg * 1e-3) / 1e9,
                     "frac_hbm": by * nel / (avg * 1e-3) / 1e9 / peak, "kernel": ops[name].kernel_name}
        if ops[name].kernel_name.startswith("prism_helm_kernel"):
            # nm triangle problems per element: per 16-column warp tile (16 // nm elements) 4 terms of
            # (8 MT) x (4 KS) x 16 DMMA work, idle columns included
            ntri = nm * (nm + 1) // 2
            fl = 2 * (8 * ((ntri + 7) // 8)) * 4 * (4 * ((ntri + 3) // 4)) * 16 // (16 // nm)
            per[name].update({"bound": "tensor (FP64 DMMA)", "flops_per_element_issued": fl,
                              "tflops": fl * nel / (avg * 1e-3) / 1e12, "frac_dmma": fl * nel / (avg * 1e-3) / 1e12 / 37.1,
                              "dmma_peak_tflops": 37.1})
        if ops[name].kernel_name.startswith("dense_helm_kernel"):
            # the DMMA GEMM issued per element: (8 MT) rows x 7 terms x (4 KS) padded columns; peak = the DMMA
            # microbenchmark of tools/fp64_peak.cu (profiles/r01_fp64_peak.jsonl)
            nc = std.GetNcoeffs()
            fl = 2 * (8 * ((nc + 7) // 8)) * 7 * (4 * ((nc + 3) // 4))
            per[name].update({"bound": "tensor (FP64 DMMA)", "flops_per_element_issued": fl,
                              "tflops": fl * nel / (avg * 1e-3) / 1e12, "frac_dmma": fl * nel / (avg * 1e-3) / 1e12 / 37.1,
                              "dmma_peak_tflops": 37.1})
    clocks = sampler.stop()
    step_ms = _mixed_step_ms(torch, ops, bufs, args.steps)
    ndof = sum(p["elements"] * p["ncoeffs"] for p in per.values())
    dom = max(per, key=lambda k: per[k]["ms"])
    # CPU baseline: the same three collections, a bounded sample of each, aggregated by DOF / time
    cpu_dof, cpu_t, kind, threads = 0.0, 0.0, "reference", 1
    for name, (_, pshape) in shapes.items():
        nel, jac, df = mesh[name]
        ns = min(nel, 6000)
        ns -= ns % 12
        g, t, kind, threads, _ = _cpu_helmholtz(po, pshape, nm, nm + 1, ns, jac[:ns].copy(),
                                                df.reshape(9, -1)[:, :ns].reshape(-1).copy(), seconds=3.0)
        cpu_dof += ns * per[name]["ncoeffs"] * (nel / ns)
        cpu_t += t * (nel / ns)
    # the same mesh with GENERAL geometry: (i) a global rotation + shear (full metric, non-extruded prisms, skewed
    # tetrahedra), (ii) warped by a smooth displacement (per-quadrature-point factors) on a smaller grid
    del ops, bufs
    torch.cuda.empty_cache()
    geo = {}
    geo["general_affine"] = _geometry_variant(
        torch, nk, peak, mixed_mesh(n, 1.0 / n, GENERAL_MAP), nm, False, gen, dev, args.steps,
        "same %d^3 mixed mesh under a global rotation + shear: general affine elements (full Laplacian metric)" % n)
    torch.cuda.empty_cache()
    nd = min(n, 40)
    shp = {"Hex": nk.eHexahedron, "Prism": nk.ePrism, "Tet": nk.eTetrahedron}
    std_of = {k: nk.StdExpansion(v, nm) for k, v in shp.items()}
    geo["deformed"] = _geometry_variant(
        torch, nk, peak, mixed_mesh_deformed(nd, 1.0 / nd, std_of, torch, dev, GENERAL_MAP), nm, True, gen, dev,
        max(3, args.steps // 2),
        "%d^3 mixed mesh, rotation + shear + smooth warp 0.05 sin sin sin: factors per quadrature point" % nd)
    torch.cuda.empty_cache()
    return {
        "geometry_variants": geo,
        "metric": "GDOF/s FP64 Helmholtz apply (mixed Hex/Prism/Tet mesh, P=6)", "value": ndof / (step_ms * 1e-3) / 1e9,
        "unit": "GDOF/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Helmholtz apply on a synthetic mixed mesh: %d^3 cubes, 60%% hexahedra, 30%% cut into 2 "
                               "prisms, 10%% into 6 tetrahedra, P=6 (nm=7), regular geometry, one collection per shape "
                               "(BASELINE configs[3])" % n,
                   "elements": {k: v["elements"] for k, v in per.items()},
                   "l2": "coefficient arrays of the three collections total %d MB per apply, no flush" % (ndof * 16 >> 20)},
        "roofline": ({"bound": "tensor", "achieved": per[dom]["tflops"], "peak": 37.1, "unit": "TFLOP/s",
                      "frac": per[dom]["frac_dmma"], "traffic": None,
                      "peak_source": "FP64 DMMA (mma.sync.m8n8k4.f64) microbenchmark tools/fp64_peak.cu, "
                                     "profiles/r01_fp64_peak.jsonl (MEASURED_PEAKS.json holds bf16 and HBM only)",
                      "flops_per_element_issued": per[dom]["flops_per_element_issued"], "kernel": per[dom]["kernel"],
                      "kernel_ms": per[dom]["ms"], "frac_hbm": per[dom]["frac_hbm"],
                      "note": "dominant kernel of the step is an FP64 tensor-core GEMM (DESIGN.md 4.3a/4.3b)"}
                     if "frac_dmma" in per[dom] else
                     {"bound": "hbm", "achieved": per[dom]["gb_per_s"], "peak": peak, "unit": "GB/s",
                     "frac": per[dom]["frac_hbm"], "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_element": per[dom]["algorithmic_bytes_per_element"], "kernel": per[dom]["kernel"],
                     "kernel_ms": per[dom]["ms"],
                     "note": "regular-geometry Helmholtz at P=6 is FP64 bound, not HBM bound: per_shape carries the DMMA "
                             "throughput of the tensor-core kernels (frac_dmma, against the measured 37.1 TFLOP/s)"}),
        "per_shape": per, "gpu_launches": 3 * args.steps,
        "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")},
        "cpu_baseline": {"value": cpu_dof / cpu_t / 1e9, "unit": "GDOF/s", "cores": threads, "kind": kind,
                         "sample": "up to 6000 elements of each shape, scaled to the mesh's element counts, reference "
                                   "MatrixFreeOps kernels (oracle/_ref AVX2 build)"},
    }


def run(args, torch, nk, po, peak, peak_src, ClockSampler):
    fn = {1: config1, 4: config4}[args.config]
    print(json.dumps(fn(args, torch, nk, po, peak, peak_src, ClockSampler)))
