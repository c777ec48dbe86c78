#!/usr/bin/env python
"""Generates tests/golden/ref_vectors.npz from the REFERENCE's own kernels.

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
Inputs are seeded; outputs come from oracle/_ref/libnekref_scalar.so, i.e. the reference's
MatrixFreeOps *Kernels.hpp / Helmholtz.h templates compiled in place (default width-1 build),
fed with tables produced by the reference's own Polylib.cpp.  The fixtures pin the oracle
(tests/test_oracle.py) and the CUDA path (tests/test_gpu_parity.py) on machines where
/root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
import pyoracle as po  # noqa: E402
from _util import random_geometry  # noqa: E402

CASES = [  # shape, nm, nq0
    (po.QUAD, 4, 5), (po.QUAD, 6, 7), (po.QUAD, 4, 6),
    (po.TRI, 4, 5), (po.TRI, 6, 7),
    (po.HEX, 3, 4), (po.HEX, 5, 6), (po.HEX, 4, 6), (po.HEX, 8, 9),
    (po.PRISM, 4, 5), (po.PRISM, 7, 8),
    (po.TET, 4, 5), (po.TET, 7, 8), (po.TET, 5, 8),
    (po.PYR, 4, 5), (po.PYR, 6, 7), (po.PYR, 4, 6),  # appended last: earlier cases keep their random streams
]
NEL = 2
LAMBDA = 1.5


def main():
    ref = po.Ref("scalar")
    rng = np.random.default_rng(20240917)
    out = {}
    # 1-D tables straight from the reference's Polylib
    for pt in (0, 1, 2):
        for n in (3, 6, 9):
            z, w, D = ref.points(pt, n)
            out["points_%d_%d_z" % (pt, n)] = z
            out["points_%d_%d_w" % (pt, n)] = w
            out["points_%d_%d_D" % (pt, n)] = D
    for shape, nm, nq0 in CASES:
        el = po.Elem(shape, nm, nq0)
        for deformed in (0, 1):
            key = "%s_%d_%d_%s" % (po.SHAPE_NAMES[shape], nm, nq0, "def" if deformed else "reg")
            jac, df = random_geometry(rng, el.dim, NEL, el.nqTot, deformed)
            x = rng.uniform(-1, 1, NEL * el.nmTot)
            f = [rng.uniform(-1, 1, NEL * el.nqTot) for _ in range(el.dim)]
            out[key + "_jac"], out[key + "_df"], out[key + "_x"] = jac, df, x
            for d in range(el.dim):
                out[key + "_f%d" % d] = f[d]
            out[key + "_bwd"] = ref.operator(po.OP_BWD, el, NEL, deformed, jac, df)(x)
            out[key + "_iprod"] = ref.operator(po.OP_IPROD, el, NEL, deformed, jac, df)(f[0])
            pd = ref.operator(po.OP_PHYSDERIV, el, NEL, deformed, jac, df)(f[0])
            for d in range(el.dim):
                out[key + "_pd%d" % d] = pd[d]
            out[key + "_helm"] = ref.operator(po.OP_HELM, el, NEL, deformed, jac, df)(x, lam=LAMBDA)
            out[key + "_ipwdb"] = ref.operator(po.OP_IPWDB, el, NEL, deformed, jac, df)(f)
    # segments in 1, 2 and 3 space dimensions (appended last: the random streams above are unchanged)
    for nm, nq0 in ((4, 5), (7, 8), (5, 8)):
        for cd in (1, 2, 3):
            el = po.Elem(po.SEG, nm, nq0, coordim=cd)
            for deformed in (0, 1):
                key = "Seg%d_%d_%d_%s" % (cd, nm, nq0, "def" if deformed else "reg")
                npt = NEL * (el.nqTot if deformed else 1)
                jac = rng.uniform(0.5, 1.5, npt)
                df = rng.uniform(-1.5, 1.5, cd * npt)
                x = rng.uniform(-1, 1, NEL * el.nmTot)
                f = [rng.uniform(-1, 1, NEL * el.nqTot) for _ in range(cd)]
                out[key + "_jac"], out[key + "_df"], out[key + "_x"] = jac, df, x
                for d in range(cd):
                    out[key + "_f%d" % d] = f[d]
                out[key + "_bwd"] = ref.operator(po.OP_BWD, el, NEL, deformed, jac, df)(x)
                out[key + "_iprod"] = ref.operator(po.OP_IPROD, el, NEL, deformed, jac, df)(f[0])
                pd = ref.operator(po.OP_PHYSDERIV, el, NEL, deformed, jac, df)(f[0])
                pd = pd if isinstance(pd, list) else [pd]
                for d in range(cd):
                    out[key + "_pd%d" % d] = pd[d]
                out[key + "_ipwdb"] = ref.operator(po.OP_IPWDB, el, NEL, deformed, jac, df)(f)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
