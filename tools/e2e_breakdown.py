#!/usr/bin/env python
"""Where the time of one nekmf_helmsolve call on host arrays goes (bench.py's e2e): the same call on device arrays
(no PCIe), on pinned host arrays, the CG loop inside it, and the two bare PCIe copies.
    python tools/e2e_breakdown.py [--nx 64]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import bench_cg  # noqa: E402
from _util import nekmf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=64)
    a = ap.parse_args()
    nk = nekmf()
    ae = bench_cg.parse_args(["--nx", str(a.nx), "--ny", str(a.nx), "--nz", str(a.nx)])
    S = bench_cg.setup(ae, comm=None)
    hs = nk.HelmSolver(S["cg"], S["ipr"], S["bwd"])
    f = torch.tensor(-(bench.LAMBDA + 3 * np.pi ** 2) * S["u_exact"])
    f_host, f_dev = f.pin_memory(), f.cuda()
    nL = S["mesh"].nLocal
    coef_host, phys_host = torch.zeros(nL, dtype=torch.float64).pin_memory(), torch.zeros_like(f).pin_memory()
    coef_dev, phys_dev = torch.zeros(nL, dtype=torch.float64, device="cuda"), torch.zeros_like(f_dev)
    out = {}
    for name, (ff, cc, pp) in {"device": (f_dev, coef_dev, phys_dev), "host": (f_host, coef_host, phys_host)}.items():
        hs.HelmSolve(ff, cc, pp, tol=bench.E2E_TOL)
        ts, ms, loops = [], [], []
        for _ in range(3):
            cc.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            its, eps = hs.HelmSolve(ff, cc, pp, tol=bench.E2E_TOL)
            ts.append((time.perf_counter() - t0) * 1e3)
            ms.append(hs.last_ms())
            loops.append(S["cg"].last_loop())
            phases = [round(v, 3) for v in hs.last_phases()]
        out[name] = {"wall_ms": round(min(ts), 3), "device_ms": round(min(ms), 3), "iterations": its,
                     "cg_loop_ms": round(min(l[0] for l in loops), 3), "cg_iterations_launched": loops[-1][1],
                     "phases_ms_h2d_pre_cg_post_d2h": phases}
    # bare copies of the same arrays
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(2):
        ev[0].record()
        f_dev.copy_(f_host, non_blocking=True)
        coef_dev.copy_(coef_host, non_blocking=True)
        ev[1].record()
        coef_host.copy_(coef_dev, non_blocking=True)
        phys_host.copy_(phys_dev, non_blocking=True)
        ev[2].record()
        torch.cuda.synchronize()
    out["h2d_ms"], out["d2h_ms"] = round(ev[0].elapsed_time(ev[1]), 3), round(ev[1].elapsed_time(ev[2]), 3)
    out["bytes_each_way"] = (f.numel() + nL) * 8

    # what precedes the copy matters: the same H2D pair after the host zeroed its coefficient array (as a caller does
    # before a solve), after an idle gap, and back to back
    def h2d_ms(prep):
        prep()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f_dev.copy_(f_host, non_blocking=True)
        coef_dev.copy_(coef_host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1), 3)
    out["h2d_after_host_zero_ms"] = [h2d_ms(lambda: coef_host.zero_()) for _ in range(3)]
    out["h2d_after_idle_50ms_ms"] = [h2d_ms(lambda: time.sleep(0.05)) for _ in range(3)]
    out["h2d_back_to_back_ms"] = [h2d_ms(lambda: None) for _ in range(3)]
    out["h2d_after_kernel_burst_ms"] = [h2d_ms(lambda: [phys_dev.mul_(1.0) for _ in range(200)]) for _ in range(3)]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
