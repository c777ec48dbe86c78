"""Algorithmic bytes and flops per element of the five operators; shared by tools/sweep.py and
tools/sweep_table.py."""


def algorithmic_bytes(op, dim, nmTot, nqTot, deformed):
    """per element, SURVEY.md 8(a)/(d): every input and output array once, geometry once"""
    g = nqTot if deformed else 1
    ndf = dim * dim
    return 8 * {"BwdTrans": nmTot + nqTot, "IProductWRTBase": nqTot + nmTot + g,
                "PhysDeriv": nqTot + dim * nqTot + ndf * g, "Helmholtz": 2 * nmTot + (ndf + 1) * g,
                "IProductWRTDerivBase": dim * nqTot + nmTot + (ndf + 1) * g}[op]


def algorithmic_flops(op, shape, nm, nq):
    """per element with the default quadrature (nq = nm + 1), counted on the reference's algorithm
    (SURVEY.md 8(a)/(d)): sum-factorised passes 2*(nq nm^d-1 ... ) flops, tensor derivatives 2*d*nq^(d+1),
    pointwise metric work."""
    if shape == "Hex":
        sf = 2 * (nq * nm ** 3 + nq ** 2 * nm ** 2 + nq ** 3 * nm)
        der, pts = 2 * 3 * nq ** 4, nq ** 3
        return {"BwdTrans": sf, "IProductWRTBase": sf + 3 * pts, "PhysDeriv": der + 15 * pts,
                "Helmholtz": 5 * sf + der + 50 * pts, "IProductWRTDerivBase": 3 * sf + 24 * pts}[op]
    if shape == "Quad":
        sf = 2 * (nq * nm ** 2 + nq ** 2 * nm)
        der, pts = 2 * 2 * nq ** 3, nq ** 2
        return {"BwdTrans": sf, "IProductWRTBase": sf + 2 * pts, "PhysDeriv": der + 6 * pts,
                "Helmholtz": 4 * sf + der + 20 * pts, "IProductWRTDerivBase": 2 * sf + 10 * pts}[op]
    # collapsed shapes (default quadrature: Gauss-Radau directions carry nm points): pass lengths follow the mode
    # triangles / pyramids of the reference kernels (BwdTransKernels.hpp:78-484); pointwise constants as for Hex / Quad
    npair = nm * (nm + 1) // 2
    if shape == "Tri":
        nq0, nq1 = nq, nm
        sf = 2 * (nq1 * npair + nq1 * nq0 * nm)
        pts = nq0 * nq1
        der = 2 * pts * (nq0 + nq1)
        return {"BwdTrans": sf, "IProductWRTBase": sf + 2 * pts, "PhysDeriv": der + 6 * pts,
                "Helmholtz": 4 * sf + der + 20 * pts, "IProductWRTDerivBase": 2 * sf + 10 * pts}[op]
    if shape in ("Prism", "Pyr", "Tet"):
        nq0, nq1, nq2 = nq, (nm if shape == "Tet" else nq), nm
        if shape == "Prism":
            s1, s2 = nq2 * nm * npair, nq2 * nq1 * nm * nm
        elif shape == "Pyr":
            s1, s2 = nq2 * (nm * (nm + 1) * (2 * nm + 1) // 6), nq2 * nq1 * nm * nm
        else:
            s1, s2 = nq2 * (nm * (nm + 1) * (nm + 2) // 6), nq2 * nq1 * npair
        sf = 2 * (s1 + s2 + nq2 * nq1 * nq0 * nm)
        pts = nq0 * nq1 * nq2
        der = 2 * pts * (nq0 + nq1 + nq2)
        return {"BwdTrans": sf, "IProductWRTBase": sf + 3 * pts, "PhysDeriv": der + 15 * pts,
                "Helmholtz": 5 * sf + der + 50 * pts, "IProductWRTDerivBase": 3 * sf + 24 * pts}[op]
    return None
