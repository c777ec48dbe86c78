#!/usr/bin/env python
"""bench.py -- GDOF/s of the FP64 Helmholtz operator apply on a synthetic 64^3 hex mesh at P=4
(nm=5 modes, nq=6 Gauss-Lobatto-Legendre points per direction; BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one Helmholtz apply over every element of the rank's mesh (262 144 elements,
32.768 M local DOF, 262 MB in + 262 MB out: larger than the 126 MB L2, so no flush is needed).
Under torchrun each rank owns its own 64^3 mesh (weak scaling, no data-path collective: the
elemental operator has no inter-element coupling).  Prints ONE JSON line (see the repo prompt for
the contract).  `--impl reference` times the reference's own CPU kernels (oracle/_ref, compiled
from /root/reference/library/MatrixFreeOps in place) on the host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

NM, NQ = 5, 6
NX = 64
LAMBDA = 1.0
SEED = 1234
E2E_TOL = 1e-9  # NekConstants::kNekIterativeTol, the reference's default IterativeSolverTolerance


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                d = json.load(f)
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded(name, default=None):
    """numbers taken from committed ncu captures (profiles/recorded.json), never measured here"""
    p = os.path.join(ROOT, "profiles", "recorded.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return json.load(f).get(name, default)
        except Exception:
            pass
    return default


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML DURING the timed region (the same counters
    as the nvidia-smi clocks line of B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._ready = threading.Event()
        self._thr = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self._stop.is_set():
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self._ready.set()
                r = int(get_reasons(h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.0005)
        except Exception as ex:  # pragma: no cover
            self.error = repr(ex)
            self._ready.set()

    def start(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        self._ready.wait(timeout=10)
        self.samples.clear()  # keep only samples taken under load

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=5)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(self.samples)}
        if self.samples:
            s = sorted(self.samples)
            out["sm_mhz"] = s[len(s) // 2]
        return out


def hex_mesh_geometry(torch, dev, nx, deformed):
    """Geometric factors of a structured nx^3 hex mesh on [0,1]^3 in the reference layout.
    regular : jac[nElmt], df[9][nElmt]  (df = diag(2/h), jac = (h/2)^3)
    deformed: same mesh warped by x += 0.05 sin(pi X) sin(pi Y) sin(pi Z) in every component,
              factors per quadrature point (SpatialDomains/GeomFactors.cpp:399-474 formulas:
              df[c*3+d] = d xi_d / d x_c, jac = det(d x / d xi))."""
    nel = nx ** 3
    h = 1.0 / nx
    if not deformed:
        jac = torch.full((nel,), (h / 2) ** 3, dtype=torch.float64, device=dev)
        df = torch.zeros((9, nel), dtype=torch.float64, device=dev)
        df[0] = df[4] = df[8] = 2.0 / h
        return jac, df.reshape(-1)
    from _util import nekmf
    nk = nekmf()
    z, _, _ = nk.points(nk.eGaussLobattoLegendre, NQ)
    zt = torch.tensor(z, dtype=torch.float64, device=dev)
    idx = torch.arange(nx, dtype=torch.float64, device=dev)
    # 1-D physical coordinates of every (element, point): [nx, NQ]
    X1 = (idx[:, None] + 0.5 * (zt[None, :] + 1.0)) * h
    import math
    s, c = torch.sin(math.pi * X1), math.pi * torch.cos(math.pi * X1)
    # element index e = ex + nx*(ey + nx*ez); point index q = i + NQ*(j + NQ*k)
    def bc(a, axis):  # broadcast a [nx,NQ] table to [ez,ey,ex,k,j,i]
        shape = [1] * 6
        shape[2 - axis] = nx
        shape[5 - axis] = NQ
        return a.reshape(shape)
    sx, sy, sz = bc(s, 0), bc(s, 1), bc(s, 2)
    cx, cy, cz = bc(c, 0), bc(c, 1), bc(c, 2)
    a = 0.05
    g = [a * cx * sy * sz, a * sx * cy * sz, a * sx * sy * cz]  # d disp / d X_d (same for every component)
    full = (nx, nx, nx, NQ, NQ, NQ)
    # F[c][d] = d x_c / d xi_d = (h/2) (delta_cd + g_d)
    F = [[(h / 2) * ((1.0 if cc == d else 0.0) + g[d]).expand(full) for d in range(3)] for cc in range(3)]
    det = (F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) - F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
           F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]))
    inv = [[None] * 3 for _ in range(3)]  # inv[d][c] = d xi_d / d x_c
    for d in range(3):
        for cc in range(3):
            r0, r1 = [r for r in range(3) if r != cc], [r for r in range(3) if r != d]
            minor = F[r0[0]][r1[0]] * F[r0[1]][r1[1]] - F[r0[0]][r1[1]] * F[r0[1]][r1[0]]
            inv[d][cc] = ((-1.0) ** (d + cc)) * minor / det
    df = torch.stack([inv[d][cc].reshape(-1) for cc in range(3) for d in range(3)])  # row c*3+d
    return det.reshape(-1).contiguous(), df.reshape(-1).contiguous()


def algorithmic_bytes_per_element(deformed):
    # 8*(2*nmTot + ndf + 1) regular, 8*(2*nmTot + (ndf+1)*nqTot) deformed
    return 8 * (2 * NM ** 3 + (10 * NQ ** 3 if deformed else 10))


def _ref_engine():
    import pyoracle as po
    try:
        return po.Ref("avx2"), "reference", "oracle/_ref libnekref_avx2.so (reference MatrixFreeOps kernels, AVX2 width 4)"
    except Exception:
        return None, "port", "oracle/libmforacle.so (plain-C restatement)"


def cpu_reference_leg(seconds_target=12.0, max_threads=None, nx=NX):
    """The reference's own MatrixFree Helmholtz kernels (AVX2 build, width 4) on the host cores, elements split over
    all threads, on the WHOLE nx^3 mesh of the benchmark; bounded by the number of repetitions."""
    import numpy as np
    import pyoracle as po
    ref, kind, variant = _ref_engine()
    threads = max_threads or (os.cpu_count() or 1)
    nel = nx ** 3
    el = po.Elem(po.HEX, NM, NQ)
    rng = np.random.default_rng(SEED)
    x = rng.uniform(-1, 1, nel * el.nmTot)
    h = 1.0 / nx
    jac = np.full(nel, (h / 2) ** 3)
    df = np.zeros((9, nel))
    df[0] = df[4] = df[8] = 2.0 / h
    df = df.reshape(-1).copy()
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
    reps = max(3, min(200, int(seconds_target / max(t1, 1e-4))))
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    tmed = ts[len(ts) // 2]
    return {"value": nel * el.nmTot / tmed / 1e9, "unit": "GDOF/s", "cores": threads, "kind": kind,
            "sample": "all %d elements (regular geometry), median of %d applies, %s" % (nel, reps, variant),
            "ms_per_step": tmed * 1e3}


def cpu_chain_leg(max_threads=None, nx=NX, max_iters=40):
    """The reference's ContField::v_HelmSolve chain + BwdTrans on the host cores (oracle/mf_oracle.c:
    mfo_chain_helmsolve driving the reference's own IProductWRTBase / Helmholtz / BwdTrans kernels from oracle/_ref;
    elements and global vector loops over all threads): the same problem as the B200 arm's `e2e` on one rank's
    nx^3 mesh.  Bounded sample: the solve is capped at max_iters iterations (constant cost per iteration)."""
    import numpy as np
    import pyoracle as po
    from _util import load_pkg_module, nekmf
    ref, kind, variant = _ref_engine()
    threads = max_threads or (os.cpu_count() or 1)
    mesh = load_pkg_module("mesh").StructuredHexMesh(nx, nx, nx, NM)
    el = po.Elem(po.HEX, NM, NQ)
    jac, df = mesh.geometry()
    nk = nekmf()
    diag = mesh.helmholtz_diagonal(nk.StdExpansion(nk.eHexahedron, NM, NQ).basis[0], LAMBDA)
    X, Y, Z = mesh.quad_coords(el.Z[0])
    f = -(LAMBDA + 3 * np.pi ** 2) * np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z)
    del X, Y, Z
    ch = po.Chain(el, mesh.nElmt, False, jac, df, LAMBDA, mesh.localToGlobal, None, mesh.nGlobal, mesh.nDir,
                  1.0 / diag[mesh.nDir:], engine=ref if ref is not None else "oracle", threads=threads)
    coef, phys = np.zeros(mesh.nLocal), np.zeros(f.size)
    ch.helmsolve(f, coef, phys, tol=E2E_TOL, maxiter=2)  # warm-up: page in, spin up the thread team
    coef[:] = 0.0
    t0 = time.perf_counter()
    its, _ = ch.helmsolve(f, coef, phys, tol=E2E_TOL, maxiter=max_iters)
    dt = time.perf_counter() - t0
    po.set_threads(1)
    applies = abs(its) + 1
    return {"value": mesh.nLocal * applies / dt / 1e9, "unit": "GDOF/s", "cores": threads, "kind": kind,
            "sample": "HelmSolve chain on the %d^3 mesh capped at %d of its CG iterations (%d Helmholtz applies), %s"
                      % (nx, max_iters, applies, variant),
            "ms_per_step": dt * 1e3, "helmholtz_applies": applies, "ms_per_apply": dt * 1e3 / applies}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=NX, help="elements per direction (default 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the compact config 1 / 3 / 4 legs (N=1 only)")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 4],
                    help="BASELINE.json configuration, 1-based: 2 (default) = configs[1], the metric's workload; "
                         "1 = 2-D quad Helmholtz P=5; 4 = mixed Hex/Prism/Tet Helmholtz P=6 (bench_configs.py)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    config = {"workload": "3D Helmholtz operator apply, synthetic %d^3 hex mesh per GPU, P=4 (nm=5, nq=6), FP64, "
                          "regular (affine) geometry, lambda=1" % args.nx,
              "elements_per_gpu": args.nx ** 3, "nm": NM, "nq": NQ,
              "l2": "inputs larger than L2 (in+out 524 MB per apply), no flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        # per-apply on the whole mesh (what `value` of the B200 arm measures) and the HelmSolve chain (what its `e2e`
        # measures): `value` below is the former, `e2e.value` the latter, so each ratio compares like with like
        cb = cpu_reference_leg(seconds_target=max(5.0, min(30.0, 0.04 * (K + W) * 10)), nx=args.nx)
        chain = cpu_chain_leg(nx=args.nx)
        line = {"impl": "reference", "metric": "GDOF/s FP64 Helmholtz apply (hex P=4)", "value": cb["value"],
                "unit": "GDOF/s", "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": cb["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": chain["value"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "ms_per_step": chain["ms_per_step"], "helmholtz_applies": chain["helmholtz_applies"],
                        "note": "Helmholtz-apply DOF/s delivered by the reference's ContField::HelmSolve chain "
                                "(IProductWRTBase -> Helmholtz lift -> Assemble -> CG -> GlobalToLocal -> BwdTrans) "
                                "on the host cores: " + chain["sample"]}}
        print(json.dumps(line))
        return

    import torch
    from _util import nekmf
    nk = nekmf()
    if not torch.cuda.is_available() or nk.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config != 2:
        if world > 1:
            raise SystemExit("bench.py --config %d is a single-GPU measurement" % args.config)
        import pyoracle as po
        import bench_configs
        peak, peak_src = measured_peaks()
        bench_configs.run(args, torch, nk, po, peak, peak_src, ClockSampler)
        return

    nel = args.nx ** 3
    ndof = nel * NM ** 3
    stdexp = nk.StdExpansion(nk.eHexahedron, NM, NQ)
    gen = torch.Generator(device="cpu").manual_seed(SEED + rank)
    x_host = (torch.rand(ndof, dtype=torch.float64, generator=gen) * 2 - 1).pin_memory()
    y_host = torch.empty(ndof, dtype=torch.float64).pin_memory()
    x = x_host.to(dev)
    y = torch.empty_like(x)

    results = {}
    for variant in ("regular", "deformed"):
        deformed = variant == "deformed"
        jac, df = hex_mesh_geometry(torch, dev, args.nx, deformed)
        coll = nk.Collection(stdexp, nel, nk.CoalescedGeomData(jac, df, deformed))
        coll.Initialise(nk.eHelmholtz)
        op = coll.m_ops[nk.eHelmholtz]
        del jac, df
        torch.cuda.empty_cache()
        op.SetLambda(LAMBDA)
        sampler = ClockSampler(local_rank)
        if variant == "regular":
            sampler.start()
        for _ in range(W):
            op.apply([x], [y])
        barrier()
        sampler.samples.clear()
        l0 = nk.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(K):
            op.apply([x], [y])  # enqueued on the current (default) stream, back to back
        ev1.record()
        barrier()
        launches = nk.launch_count() - l0
        clocks = None
        if variant == "regular":
            if len(sampler.samples) < 8:
                # K applies last only a few ms, shorter than an NVML poll: keep the same kernel running
                # back to back (untimed) until the sampler has seen the clocks under this load
                t_end = time.perf_counter() + 0.4
                while time.perf_counter() < t_end:
                    for _ in range(50):
                        op.apply([x], [y])
                    torch.cuda.synchronize()
            clocks = sampler.stop()
        total_ms = ev0.elapsed_time(ev1)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms_max = float(t.item())
        results[variant] = {"total_ms": total_ms_max, "ms_per_step": total_ms_max / K, "kernel_ms_avg": total_ms / K,
                            "launches": launches, "clocks": clocks, "kernel": op.kernel_name}
        if variant == "regular":
            # ---- end-to-end through the public host-array call: pinned host in, pinned host out
            Ke = max(3, min(K, 10))
            for _ in range(2):
                op.apply([x_host], [y_host])
            barrier()
            t0 = time.perf_counter()
            for _ in range(Ke):
                op.apply([x_host], [y_host])  # synchronous: H2D + kernel + D2H
            torch.cuda.synchronize()
            te = (time.perf_counter() - t0) / Ke
            tt = torch.tensor([te], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            results["e2e_apply_s"] = float(tt.item())
            chk = float(torch.linalg.vector_norm(y).item())
            chk_host = float(torch.linalg.vector_norm(y_host).item())
            results["checksum"] = (chk, chk_host)
        del coll, op
        torch.cuda.empty_cache()

    del x, y, x_host, y_host
    torch.cuda.empty_cache()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import numpy as np
    import bench_cg
    import _cg_check
    comm = nk.Comm.from_torch_distributed() if world > 1 else None

    # ---- e2e: the reference's solve chain through the public host-array call (nekmf_helmsolve =
    # ContField::v_HelmSolve + BwdTrans): forcing at the quadrature points in pinned host memory -> H2D ->
    # IProductWRTBase -> Dirichlet lift -> Assemble (+ interface exchange) -> CG to the reference's default tolerance
    # -> GlobalToLocal -> BwdTrans -> D2H of coefficients and physical values.  One nx^3 slab per rank of a
    # nx x nx x (nx*N) mesh (weak scaling; N > 1 exchanges the slab interfaces every iteration).
    ae = bench_cg.parse_args(["--nx", str(args.nx), "--ny", str(args.nx), "--nz", str(args.nx * world)])
    S = bench_cg.setup(ae, comm=comm)
    hs = nk.HelmSolver(S["cg"], S["ipr"], S["bwd"])
    f_host = torch.tensor(-(LAMBDA + 3 * np.pi ** 2) * S["u_exact"]).pin_memory()
    coef_host = torch.zeros(S["mesh"].nLocal, dtype=torch.float64).pin_memory()
    phys_host = torch.zeros(f_host.numel(), dtype=torch.float64).pin_memory()
    hs.HelmSolve(f_host, coef_host, phys_host, tol=E2E_TOL)  # warm-up solve (also captures the iteration graphs)
    Ke = 3
    l0 = nk.launch_count()
    te = 0.0
    zero_dev = torch.zeros(S["mesh"].nLocal, dtype=torch.float64, device=dev)
    for _ in range(Ke):
        # the caller's initial guess / Dirichlet values (zero): not part of the solve.  Reset by a DMA write, not by the
        # CPU: 262 MB of lines the host has just written are read back by the next H2D at 42 instead of 55 GB/s
        # (tools/e2e_breakdown.py: 17.0 against 12.9 ms for the pair of input arrays) -- an artefact of timing the call
        # right after a host memset, not a property of the solve
        coef_host.copy_(zero_dev)
        barrier()
        t0 = time.perf_counter()
        its_e, eps_e = hs.HelmSolve(f_host, coef_host, phys_host, tol=E2E_TOL)  # synchronous
        te += (time.perf_counter() - t0) / Ke
    tt = torch.tensor([te], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    u_err = float(np.abs(phys_host.numpy() - S["u_exact"]).max())
    e2e = {"s": float(tt.item()), "iterations": its_e, "applies": its_e + 1, "final_eps": eps_e,
           "launches": (nk.launch_count() - l0) // Ke, "device_ms": hs.last_ms(), "u_err": u_err,
           "h2d": (f_host.numel() + coef_host.numel()) * 8, "d2h": (coef_host.numel() + phys_host.numel()) * 8,
           "ndof_local": S["mesh"].nLocal}
    e2e["phases_ms"] = dict(zip(("h2d", "iproduct_lift_assemble", "cg", "globaltolocal_bwdtrans", "d2h"),
                                [round(v, 3) for v in hs.last_phases()]))
    del hs, S, f_host, coef_host, phys_host, zero_dev
    torch.cuda.empty_cache()

    # ---- BASELINE configs[4]: 2^20 hex elements at P=4 split in z-slabs over the ranks (STRONG scaling),
    # fixed iteration count, device-event time of the iteration loop, max over ranks
    ac = bench_cg.parse_args(["--nx", "64", "--ny", "128", "--nz", "128", "--iters", "60"])
    S = bench_cg.setup(ac, comm=comm)
    cg_line = bench_cg.measure(S, ac)
    del S
    torch.cuda.empty_cache()
    # ---- parity of the sharded solve on a mesh the CPU oracle solves in seconds
    ap_ = bench_cg.parse_args(["--nx", "8", "--ny", "8", "--nz", str(max(8, 4 * world))])
    S = bench_cg.setup(ap_, comm=comm)
    cg_parity = _cg_check.reduce_ok(S, _cg_check.run_check(S, ap_))
    del S
    torch.cuda.empty_cache()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            del comm
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    reg, dfm = results["regular"], results["deformed"]

    def roof(r, deformed):
        bytes_launch = algorithmic_bytes_per_element(deformed) * nel
        ach = bytes_launch / (r["kernel_ms_avg"] * 1e-3) / 1e9
        traffic = recorded("traffic_bytes_deformed" if deformed else "traffic_bytes_regular")
        out = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
               "traffic": traffic, "peak_source": peak_src,
               "algorithmic_bytes_per_element": algorithmic_bytes_per_element(deformed),
               "kernel": r["kernel"], "kernel_ms": r["kernel_ms_avg"]}
        if traffic:
            out["frac_by_dram_bytes"] = traffic / (r["kernel_ms_avg"] * 1e-3) / 1e9 / peak
            out["traffic_source"] = recorded("traffic_source")
        return out

    value = world * ndof / (reg["ms_per_step"] * 1e-3) / 1e9
    line = {
        "metric": "GDOF/s FP64 Helmholtz apply (hex P=4)", "value": value, "unit": "GDOF/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": reg["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "roofline": roof(reg, False),
        "e2e": {"value": world * e2e["ndof_local"] * e2e["applies"] / e2e["s"] / 1e9, "unit": "GDOF/s",
                "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": e2e["s"] * 1e3,
                "helmholtz_applies": e2e["applies"], "cg_iterations": e2e["iterations"], "tol": E2E_TOL,
                "device_ms": e2e["device_ms"], "phases_ms": e2e["phases_ms"], "gpu_launches": e2e["launches"],
                "max_abs_error_vs_analytic_solution": e2e["u_err"],
                "note": "Helmholtz-apply DOF/s delivered by one nekmf_helmsolve call (= ContField::HelmSolve + "
                        "BwdTrans) on pinned host arrays: H2D forcing+coefficients, IProductWRTBase, Dirichlet lift, "
                        "Assemble, CG to tol, GlobalToLocal, BwdTrans, D2H coefficients+physical values; one %d^3 "
                        "slab per rank%s" % (args.nx, ", slab interfaces exchanged over NVLink every iteration"
                                             if world > 1 else "")},
        "e2e_apply": {"value": world * ndof / results["e2e_apply_s"] / 1e9, "unit": "GDOF/s",
                      "h2d_bytes_per_step": ndof * 8, "d2h_bytes_per_step": ndof * 8,
                      "ms_per_step": results["e2e_apply_s"] * 1e3,
                      "note": "one nekmf_op_apply(NEKMF_HOST) on pinned host arrays: H2D + kernel + D2H per apply "
                              "(PCIe-bound by construction)"},
        "gpu_launches": reg["launches"],
        "clocks": {k: reg["clocks"][k] for k in ("sm_mhz", "sm_max_mhz", "reasons")} if reg["clocks"] else None,
        "variants": {"deformed": {"value": world * ndof / (dfm["ms_per_step"] * 1e-3) / 1e9, "unit": "GDOF/s",
                                  "ms_per_step": dfm["ms_per_step"], "roofline": roof(dfm, True),
                                  "workload": "same mesh warped by 0.05 sin(pi x) sin(pi y) sin(pi z): geometric "
                                              "factors per quadrature point"},
                     "cg": dict(cg_line, scaling="strong",
                                workload="matrix-free CG Helmholtz solve, 2^20 hex elements P=4 in z-slabs over the "
                                         "ranks (BASELINE configs[4]), Jacobi preconditioner, fixed 60 iterations")},
        "cg_parity": cg_parity,
        "checksum_l2": results["checksum"][0],
    }
    fp64_peak = recorded("fp64_tflops_measured")
    flops_elt = recorded("helm_flops_per_element_regular", 2 * 14700)
    if fp64_peak:
        ach = flops_elt * nel / (reg["kernel_ms_avg"] * 1e-3) / 1e12
        line["roofline_fp64"] = {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                                 "frac": ach / fp64_peak, "flops_per_element": flops_elt}
    if world == 1 and not args.no_variants:
        # the other BASELINE.json configurations, compact, so that they are on the driver's record:
        # configs[0] (quad P=5 Helmholtz), configs[2] (hex operator sweep points), configs[3] (mixed mesh P=6)
        import bench_configs
        import pyoracle as po
        import sweep

        def compact(d):
            r = d.get("roofline", {})
            return {"value": d.get("value"), "unit": d.get("unit"), "ms_per_step": d.get("ms_per_step"),
                    "workload": d.get("config", {}).get("workload"),
                    "roofline": {k: r.get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "kernel") if k in r},
                    "e2e": d.get("e2e", {}).get("value"), "cpu_baseline": (d.get("cpu_baseline") or {}).get("value"),
                    **({"geometry_variants": {k: {kk: v[kk] for kk in ("workload", "value", "unit", "ms_per_step",
                                                                       "frac_hbm_of_step")} |
                                                  {"kernels": {sh: [ps["kernel"], round(ps["ms"], 4), round(ps["frac_hbm"], 3)]
                                                               for sh, ps in v["per_shape"].items()}}
                                              for k, v in d["geometry_variants"].items()}}
                       if "geometry_variants" in d else {})}

        class _A:
            nx, steps, warmup = 64, 20, 3
        for name, fn in (("config1_quad_p5", bench_configs.config1), ("config4_mixed_p6", bench_configs.config4)):
            try:
                line["variants"][name] = compact(fn(_A, torch, nk, po, peak, peak_src, ClockSampler))
            except Exception as ex:
                line["variants"][name] = {"failed": repr(ex)}
            torch.cuda.empty_cache()
        try:
            pts = []
            for nm_ in (3, 5, 8, 11):
                for rec in sweep.iter_points("Hex", nm_, nm_, "regular,deformed", ["BwdTrans", "IProductWRTBase", "PhysDeriv"],
                                             5, 1 << 26):
                    pts.append({k: rec[k] for k in ("op", "nm", "geometry", "ms", "gdof_per_s", "gb_per_s", "frac_hbm",
                                                    "kernel")})
            line["variants"]["config3_hex_sweep_points"] = pts
        except Exception as ex:
            line["variants"]["config3_hex_sweep_points"] = {"failed": repr(ex)}
        torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_reference_leg(seconds_target=6.0, nx=args.nx)
            line["cpu_baseline"]["chain"] = cpu_chain_leg(nx=args.nx)
        except Exception as ex:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": "GDOF/s", "cores": 0, "kind": "reference",
                                    "sample": "failed: %r" % (ex,)}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        del comm
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
