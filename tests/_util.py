"""Shared helpers for the test-suite: module loading, synthetic geometry, error norms."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ithaca-sem_b200")


def load_pkg_module(name):
    """The package directory is named `ithaca-sem_b200` (not an identifier): load modules by path."""
    key = "ithaca_sem_b200_" + name
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(PKG, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod


def nekmf():
    return load_pkg_module("nekmf")


def rel_errs(got, ref):
    """relative L2 and Linf errors (north_star: 1e-12)."""
    got, ref = np.asarray(got), np.asarray(ref)
    d = got - ref
    l2 = np.sqrt(np.sum(d * d)) / max(np.sqrt(np.sum(ref * ref)), 1e-300)
    li = np.abs(d).max() / max(np.abs(ref).max(), 1e-300)
    return l2, li


def random_geometry(rng, dim, nel, nq_tot, deformed):
    """Random but well-conditioned geometric factors in the reference layout:
    jac [nel] | [nel*nq], df [dim*dim][nel | nel*nq] with df[c*dim+d] = d xi_d / d x_c."""
    npt = nel * (nq_tot if deformed else 1)
    jac = rng.uniform(0.5, 1.5, npt)
    df = rng.uniform(-0.3, 0.3, (dim * dim, npt))
    for d in range(dim):
        df[d * dim + d] += 1.5
    return np.ascontiguousarray(jac), np.ascontiguousarray(df.reshape(-1))
