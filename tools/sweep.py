#!/usr/bin/env python
"""Operator sweep (BASELINE.json configs[2] and [3]): every operator x shape x order on device-resident
arrays, timed with CUDA events on the launching stream, reported against the measured HBM peak.

    python tools/sweep.py [--shapes Hex,Quad,Tri,Prism,Tet] [--nm 2..11] [--ops all] [--out file.jsonl]

Element counts follow SURVEY.md 8(d) config 3: nElmt = 2^27 / (nmTot + nqTot) (about 1 GB of in+out
arrays, far larger than L2), capped so the deformed geometry (10 doubles per point) stays below 24 GB.
One JSON line per (shape, operator, nm, geometry): ms, GDOF/s, algorithmic GB/s, fraction of peak."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from _util import nekmf  # noqa: E402
import bench  # noqa: E402
from flop_model import algorithmic_bytes, algorithmic_flops  # noqa: E402

OPS = {"BwdTrans": 0, "Helmholtz": 1, "IProductWRTBase": 2, "IProductWRTDerivBase": 3, "PhysDeriv": 4}


def iter_points(shapes_arg, lo, hi, geom_arg, ops, reps, words):
    """yields one record per (shape, order, geometry, operator); device-resident arrays, CUDA-event timing"""
    nk = nekmf()
    dev = torch.device("cuda", torch.cuda.current_device())
    shapes = {"Quad": nk.eQuadrilateral, "Tri": nk.eTriangle, "Hex": nk.eHexahedron, "Prism": nk.ePrism,
              "Pyr": nk.ePyramid, "Tet": nk.eTetrahedron}
    peak, peak_src = bench.measured_peaks()
    fp64_peak = bench.recorded("fp64_tflops_measured")  # DFMA microbenchmark, profiles/r01_fp64_peak.jsonl
    gen = torch.Generator(device=dev).manual_seed(1234)
    for sname in shapes_arg.split(","):
        shape = shapes[sname]
        for nm in range(int(lo), int(hi) + 1):
            std = nk.StdExpansion(shape, nm)
            dim, nmTot, nqTot = std.dim, std.GetNcoeffs(), std.GetTotPoints()
            nel = words // (nmTot + nqTot)
            nel = min(nel, int(24e9 / (8 * (dim * dim + 1) * nqTot)))
            for gname in geom_arg.split(","):
                deformed = gname == "deformed"
                npt = nel * (nqTot if deformed else 1)
                jac = torch.rand(npt, dtype=torch.float64, device=dev, generator=gen) + 0.5
                df = (torch.rand(dim * dim, npt, dtype=torch.float64, device=dev, generator=gen) - 0.5) * 0.6
                for d in range(dim):
                    df[d * dim + d] += 1.5
                if gname == "regular_diag":  # axis-aligned boxes: the coefficient-space Helmholtz kernel
                    for n in range(dim * dim):
                        if n % (dim + 1) != 0:
                            df[n] = 0.0
                geom = nk.CoalescedGeomData(jac, df.reshape(-1), deformed)
                coll = nk.Collection(std, nel, geom)
                for opn in ops:
                    op = OPS[opn]
                    cin = opn in ("BwdTrans", "Helmholtz")
                    cout = opn not in ("BwdTrans", "PhysDeriv")
                    nin = dim if opn == "IProductWRTDerivBase" else 1
                    nout = dim if opn == "PhysDeriv" else 1
                    ins = [torch.rand(nel * (nmTot if cin else nqTot), dtype=torch.float64, device=dev, generator=gen)
                           for _ in range(nin)]
                    outs = [torch.empty(nel * (nmTot if cout else nqTot), dtype=torch.float64, device=dev)
                            for _ in range(nout)]
                    coll.Initialise(op)
                    o = coll.m_ops[op]
                    if opn == "Helmholtz":
                        o.SetLambda(1.0)
                    for _ in range(3):
                        o.apply(ins, outs)
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
                    ev[0].record()
                    for i in range(reps):
                        o.apply(ins, outs)
                        ev[i + 1].record()
                    torch.cuda.synchronize()
                    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
                    ms = ts[len(ts) // 2]
                    by = algorithmic_bytes(opn, dim, nmTot, nqTot, deformed)
                    gbs = by * nel / (ms * 1e-3) / 1e9
                    rec = {"shape": sname, "op": opn, "nm": nm, "P": nm - 1, "nq": std.nq, "geometry": gname,
                           "nElmt": nel, "ms": round(ms, 4), "gdof_per_s": round(nel * nmTot / (ms * 1e-3) / 1e9, 2),
                           "bytes_per_element": by, "gb_per_s": round(gbs, 1), "frac_hbm": round(gbs / peak, 3),
                           "kernel": o.kernel_name}
                    fl = algorithmic_flops(opn, sname, nm, std.nq[0])
                    if fl is not None and fp64_peak:
                        tf = fl * nel / (ms * 1e-3) / 1e12
                        # flops of the REFERENCE's algorithm per second over the DFMA peak: a throughput ratio in the
                        # reference's currency, NOT a pipe utilisation -- a kernel that needs fewer flops
                        # (coefficient-space Helmholtz) exceeds 1.0 here.  Issued-flop fractions (frac_dmma below,
                        # bench.py's roofline_fp64) are the roofline numbers.
                        rec.update({"flops_per_element_reference_algorithm": fl, "tflops": round(tf, 2),
                                    "frac_fp64_of_reference_flops": round(tf / fp64_peak, 3),
                                    "bound": "fp64" if tf / fp64_peak > gbs / peak else "hbm"})
                    if o.kernel_name.startswith("dense_helm_kernel"):
                        # the DMMA GEMM actually issued: (8 MT) rows x (nT * 4 KS) padded columns per element,
                        # against the measured DMMA peak (profiles/r01_fp64_peak.jsonl: 37.1 TFLOP/s)
                        nT = 1 + dim * (dim + 1) // 2
                        fl = 2 * (8 * ((nmTot + 7) // 8)) * nT * (4 * ((nmTot + 3) // 4))
                        tf = fl * nel / (ms * 1e-3) / 1e12
                        rec.update({"flops_per_element_issued": fl, "tflops": round(tf, 2),
                                    "frac_dmma": round(tf / 37.1, 3), "bound": "fp64"})
                    if o.kernel_name.startswith(("prism_helm_kernel", "prism_gen_kernel")):
                        # nm triangle problems per element, 4 (extruded) or 8 (general) terms; per 16-column warp tile
                        # (16 // nm elements) terms x (8 MT) x (4 KS) x 16 DMMA work, idle columns included
                        ntri, terms = nm * (nm + 1) // 2, (4 if o.kernel_name.startswith("prism_helm") else 8)
                        fl = 2 * (8 * ((ntri + 7) // 8)) * terms * (4 * ((ntri + 3) // 4)) * 16 // (16 // nm)
                        tf = fl * nel / (ms * 1e-3) / 1e12
                        rec.update({"flops_per_element_issued": fl, "tflops": round(tf, 2),
                                    "frac_dmma": round(tf / 37.1, 3), "bound": "fp64"})
                    yield rec
                    del ins, outs, o
                    coll.m_ops.pop(op)
                del coll, geom, jac, df
                torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="Hex")
    ap.add_argument("--nm", default="3..11")
    ap.add_argument("--ops", default="all")
    ap.add_argument("--geom", default="regular,deformed")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--words", type=int, default=1 << 27, help="in+out doubles per apply (sets nElmt)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    lo, hi = (a.nm.split("..") + [a.nm])[:2] if ".." in a.nm else (a.nm, a.nm)
    ops = list(OPS) if a.ops == "all" else a.ops.split(",")
    out = open(a.out, "w") if a.out else None
    for rec in iter_points(a.shapes, lo, hi, a.geom, ops, a.reps, a.words):
        line = json.dumps(rec)
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()
    if out:
        peak, peak_src = bench.measured_peaks()
        out.write(json.dumps({"hbm_peak_gb_per_s": peak, "peak_source": peak_src,
                              "fp64_peak_tflops": bench.recorded("fp64_tflops_measured")}) + "\n")
        out.close()


if __name__ == "__main__":
    main()
