#!/usr/bin/env python
"""Minimal driver for ncu captures: `reps` Helmholtz applies of one geometry variant on the bench mesh.
    ncu --set full ... python tools/prof_helm.py --variant regular|deformed|sheared [--nx 64] [--nm 5]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench  # noqa: E402
from _util import nekmf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="regular")
ap.add_argument("--nx", type=int, default=64)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--op", default="helm")
ap.add_argument("--ab", default=None, help="ENVVAR=v1,v2: alternate the knob between rounds of `reps` applies (same process, same buffers)")
a = ap.parse_args()
nk = nekmf()
dev = torch.device("cuda", 0)
nel = a.nx ** 3
std = nk.StdExpansion(nk.eHexahedron, bench.NM, bench.NQ)
jac, df = bench.hex_mesh_geometry(torch, dev, a.nx, a.variant == "deformed")
if a.variant == "sheared":
    df = df.clone()
    df.view(9, -1)[1, -1] = 1e-300  # metric no longer exactly diagonal -> quadrature-space kernel
coll = nk.Collection(std, nel, nk.CoalescedGeomData(jac, df, a.variant == "deformed"))
x = torch.rand(nel * bench.NM ** 3, dtype=torch.float64, device=dev) * 2 - 1
y = torch.empty_like(x)
coll.Initialise(nk.eHelmholtz)
op = coll.m_ops[nk.eHelmholtz]
op.SetLambda(1.0)
if a.ab:
    name, vals = a.ab.split("=")
    for rnd in range(4):
        for v in vals.split(","):
            if v == "-":
                os.environ.pop(name, None)
            else:
                os.environ[name] = v
            for _ in range(3):
                op.apply([x], [y])
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            evs[0].record()
            for i in range(a.reps):
                op.apply([x], [y])
            evs[1].record()
            torch.cuda.synchronize()
            print("round %d %s=%s: %.4f ms per apply" % (rnd, name, v, evs[0].elapsed_time(evs[1]) / a.reps))
    sys.exit(0)
evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
evs[0].record()
for i in range(a.reps):
    op.apply([x], [y])
    evs[i + 1].record()
torch.cuda.synchronize()
ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.reps)]
print(op.kernel_name, "ms:", " ".join("%.4f" % t for t in ts))
