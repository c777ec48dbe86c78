"""Synthetic structured hexahedral meshes: the assembly-map data the CG path consumes.

The reference builds `m_localToGlobalMap`, `m_localToGlobalSign`, the Dirichlet-first global numbering
and the universal (cross-rank) numbering in AssemblyMapCG's constructor (MultiRegions/AssemblyMap/
AssemblyMapCG.cpp, ~2.9k lines, out of scope: SURVEY.md section 2).  For the structured meshes of the
benchmark configurations the same arrays are generated here directly:

  * C0 connectivity of the modal hex expansion (StdRegions/StdHexExp.cpp:960-1240): in every direction
    mode 0 is the vertex at xi=-1, mode 1 the vertex at xi=+1, modes >= 2 are interior; all elements
    share the global axis orientation, so no sign changes are needed (m_signChange == false);
  * global numbering with the Dirichlet DOFs (the surface of the box) first, as AssemblyMapCG orders
    them (CG then runs on [nDir, nGlobal), NekLinSysIterCG.cpp:112-126);
  * an element partition into z-slabs, one per rank, with the interface-plane DOF lists (identically
    ordered on both sides) and the 0/1 ownership mask that Gs::Unique would produce
    (AssemblyMapCG.cpp:2551-2569).

Pure numpy: this is host-side setup, not the data path.
"""
import numpy as np


def _gi(e, p, nm):
    """1-D lattice index of local mode p of element e (vertex modes 0,1; interior 2..nm-1)."""
    base = e * (nm - 1)
    return base + np.where(p == 0, 0, np.where(p == 1, nm - 1, p - 1))


class StructuredHexMesh:
    """nx x ny x nz hexahedra on [0,Lx]x[0,Ly]x[0,Lz], nm modes per direction.  `slab=(rank, nranks)` restricts the
    mesh to this rank's z-slab of elements; `part=(px, py, pz)` with `rank` to its box of a px x py x pz partition
    (rank = rx + px (ry + py rz); up to 26 neighbours, edge / corner DOFs held by 4 / 8 ranks).  Numbering is then
    rank-local; `universal` holds the cross-rank id of every local DOF, from which the
    interface lists and the ownership mask are derived exactly as Gs::Init / Gs::Unique derive them
    (interface_from_universal_maps below)."""

    def __init__(self, nx, ny, nz, nm, lengths=(1.0, 1.0, 1.0), slab=(0, 1), part=None, rank=None):
        self.nx, self.ny, self.nz, self.nm = nx, ny, nz, nm
        self.L = tuple(float(v) for v in lengths)
        if part is None:
            self.rank, self.nranks = slab
            part = (1, 1, self.nranks)
        else:
            self.rank, self.nranks = int(rank), part[0] * part[1] * part[2]
        self.part = tuple(int(v) for v in part)
        if not 0 <= self.rank < self.nranks or self.part[0] > nx or self.part[1] > ny or self.part[2] > nz:
            raise ValueError("bad partition")
        self.h = (self.L[0] / nx, self.L[1] / ny, self.L[2] / nz)
        m1 = nm - 1
        self.Gx, self.Gy, self.Gz = nx * m1 + 1, ny * m1 + 1, nz * m1 + 1
        (self.ex0, self.ex1), (self.ey0, self.ey1), (self.ez0, self.ez1) = self._ranges(self.rank)
        self.nxl, self.nyl, self.nzl = self.ex1 - self.ex0, self.ey1 - self.ey0, self.ez1 - self.ez0
        self.nElmt = self.nxl * self.nyl * self.nzl
        # inclusive global lattice ranges of this rank's box
        self.gx0, self.gx1 = self.ex0 * m1, self.ex1 * m1
        self.gy0, self.gy1 = self.ey0 * m1, self.ey1 * m1
        self.gz0, self.gz1 = self.ez0 * m1, self.ez1 * m1
        self.Gxl, self.Gyl, self.Gzl = self.gx1 - self.gx0 + 1, self.gy1 - self.gy0 + 1, self.gz1 - self.gz0 + 1
        self._number()

    def _ranges(self, rank):
        """element ranges [e0, e1) per axis of `rank`: contiguous, remainder spread over the first boxes"""
        px, py, pz = self.part
        r3 = (rank % px, (rank // px) % py, rank // (px * py))
        out = []
        for n, p, r in zip((self.nx, self.ny, self.nz), self.part, r3):
            base, rem = divmod(n, p)
            counts = [base + (1 if i < rem else 0) for i in range(p)]
            out.append((sum(counts[:r]), sum(counts[:r]) + counts[r]))
        return out

    def _universal_of(self, rank):
        """universal ids (1 + global lattice index) of `rank`'s local lattice [Gzl,Gyl,Gxl], and the Dirichlet flags.
        Dirichlet DOFs take part in the exchange like any other, as they do in AssemblyMapCG's universal map."""
        m1 = self.nm - 1
        (x0, x1), (y0, y1), (z0, z1) = self._ranges(rank)
        gx, gy, gz = np.arange(x0 * m1, x1 * m1 + 1), np.arange(y0 * m1, y1 * m1 + 1), np.arange(z0 * m1, z1 * m1 + 1)
        uid = 1 + gx[None, None, :] + self.Gx * (gy[None, :, None] + self.Gy * gz[:, None, None]).astype(np.int64)
        bnd = ((gx == 0) | (gx == self.Gx - 1))[None, None, :] | ((gy == 0) | (gy == self.Gy - 1))[None, :, None] | \
              ((gz == 0) | (gz == self.Gz - 1))[:, None, None]
        return uid, bnd

    # ------------------------------------------------------------------ numbering
    def _number(self):
        uid, on_bnd = self._universal_of(self.rank)
        flat = on_bnd.reshape(-1)
        self.nGlobal = flat.size
        self.nDir = int(flat.sum())
        ids = np.empty(flat.size, dtype=np.int64)
        ids[flat] = np.arange(self.nDir)
        ids[~flat] = self.nDir + np.arange(flat.size - self.nDir)
        self.lattice_ids = ids.reshape(self.Gzl, self.Gyl, self.Gxl)  # rank-local global id of every lattice point
        self.dirichlet = on_bnd
        nm = self.nm
        p = np.arange(nm)
        ex, ey, ez = np.arange(self.nxl), np.arange(self.nyl), np.arange(self.nzl)
        gx = _gi(ex[:, None], p[None, :], nm)                       # [nxl, nm] box-local lattice column
        gy = _gi(ey[:, None], p[None, :], nm)
        gzl = _gi(ez[:, None], p[None, :], nm)
        # local index: e = ex + nxl*(ey + nyl*ez), mode = p + nm*(q + nm*r)
        l2g = self.lattice_ids[gzl[:, None, None, :, None, None], gy[None, :, None, None, :, None],
                               gx[None, None, :, None, None, :]]   # [ez, ey, ex, r, q, p]
        self.localToGlobal = np.ascontiguousarray(l2g.reshape(-1), dtype=np.int32)
        self.nLocal = self.localToGlobal.size
        # index of every local element in the unpartitioned mesh (e = ex + nx*(ey + ny*ez))
        self.element_ids = ((self.ex0 + ex)[None, None, :] + self.nx * ((self.ey0 + ey)[None, :, None] +
                            self.ny * (self.ez0 + ez)[:, None, None])).reshape(-1)
        # universal id of every rank-local global DOF
        self.universal = np.zeros(self.nGlobal, dtype=np.int64)
        self.universal[self.lattice_ids.reshape(-1)] = uid.reshape(-1)
        # ---- partition interfaces from the universal ids of the (at most 26) adjacent boxes
        px, py, pz = self.part
        rx, ry, rz = self.rank % px, (self.rank // px) % py, self.rank // (px * py)
        maps = [None] * self.nranks
        maps[self.rank] = self.universal
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    qx, qy, qz = rx + dx, ry + dy, rz + dz
                    if (dx, dy, dz) != (0, 0, 0) and 0 <= qx < px and 0 <= qy < py and 0 <= qz < pz:
                        q = qx + px * (qy + py * qz)
                        maps[q] = self._universal_of(q)[0].reshape(-1)
        self.peers, self.interface_lists, self.ownerMask = interface_from_universal_maps(maps, self.rank)

    # ------------------------------------------------------------------ geometry (regular boxes)
    def geometry(self):
        """jac[nElmt], df[9*nElmt] of the axis-aligned boxes (GeomFactors.cpp:399-474)."""
        hx, hy, hz = self.h
        jac = np.full(self.nElmt, hx * hy * hz / 8.0)
        df = np.zeros((9, self.nElmt))
        df[0], df[4], df[8] = 2.0 / hx, 2.0 / hy, 2.0 / hz
        return jac, df.reshape(-1).copy()

    def quad_coords(self, z):
        """physical coordinates of every quadrature point, arrays [nElmt*nq^3] in the reference's
        [elmt][k][j][i] order; z = 1-D quadrature nodes on [-1,1]."""
        nq = len(z)
        hx, hy, hz = self.h
        X1 = (self.ex0 + np.arange(self.nxl)[:, None] + 0.5 * (z[None, :] + 1.0)) * hx
        Y1 = (self.ey0 + np.arange(self.nyl)[:, None] + 0.5 * (z[None, :] + 1.0)) * hy
        Z1 = (self.ez0 + np.arange(self.nzl)[:, None] + 0.5 * (z[None, :] + 1.0)) * hz
        shape = (self.nzl, self.nyl, self.nxl, nq, nq, nq)
        X = np.broadcast_to(X1[None, None, :, None, None, :], shape).reshape(-1)
        Y = np.broadcast_to(Y1[None, :, None, None, :, None], shape).reshape(-1)
        Z = np.broadcast_to(Z1[:, None, None, :, None, None], shape).reshape(-1)
        return np.ascontiguousarray(X), np.ascontiguousarray(Y), np.ascontiguousarray(Z)

    # ------------------------------------------------------------------ matrix-free Jacobi
    def helmholtz_diagonal(self, basis, lam):
        """Diagonal of the assembled Helmholtz matrix without forming elemental matrices (replaces
        PreconditionerDiagonal::DiagonalPreconditionerSum, PreconditionerDiagonal.cpp:98-162): for an
        axis-aligned box diag_e[pqr] = J (lam m_r m_q m_p + G00 m_r m_q k_p + G11 m_r k_q m_p +
        G22 k_r m_q m_p) with m = diag(B W B^T), k = diag(DB W DB^T).  Returns this rank's LOCAL sum;
        interface DOFs still need the cross-rank exchange."""
        nm, nq = self.nm, basis.nq
        B = basis.bdata.reshape(nm, nq)
        dB = basis.dbdata.reshape(nm, nq)
        m = (B * B) @ basis.W
        k = (dB * dB) @ basis.W
        hx, hy, hz = self.h
        J = hx * hy * hz / 8.0
        g0, g1, g2 = (2.0 / hx) ** 2, (2.0 / hy) ** 2, (2.0 / hz) ** 2
        mr, mq, mp = m[:, None, None], m[None, :, None], m[None, None, :]
        kr, kq, kp = k[:, None, None], k[None, :, None], k[None, None, :]
        d = J * (lam * mr * mq * mp + g0 * mr * mq * kp + g1 * mr * kq * mp + g2 * kr * mq * mp)
        diag = np.zeros(self.nGlobal)
        np.add.at(diag, self.localToGlobal, np.tile(d.reshape(-1), self.nElmt))
        return diag


class StructuredQuadMesh:
    """nx x ny quadrilaterals on [0,Lx]x[0,Ly], nm modes per direction: the 2-D counterpart of
    StructuredHexMesh for BASELINE configs[0] (the reference's Helmholtz2D_modal session: structured quads,
    P=5, homogeneous Dirichlet on the whole boundary).  C0 connectivity of the modal quad expansion
    (StdRegions/StdQuadExp.cpp): mode 0 / 1 are the vertices at xi=-1 / +1, modes >= 2 interior; all
    elements share the global orientation, so no sign changes.  Dirichlet DOFs are numbered first."""

    def __init__(self, nx, ny, nm, lengths=(1.0, 1.0)):
        self.nx, self.ny, self.nm = nx, ny, nm
        self.L = tuple(float(v) for v in lengths)
        self.h = (self.L[0] / nx, self.L[1] / ny)
        self.nElmt = nx * ny
        m1 = nm - 1
        self.Gx, self.Gy = nx * m1 + 1, ny * m1 + 1
        on_bnd = np.zeros((self.Gy, self.Gx), dtype=bool)
        on_bnd[:, 0] = on_bnd[:, -1] = True
        on_bnd[0, :] = on_bnd[-1, :] = True
        flat = on_bnd.reshape(-1)
        self.nGlobal, self.nDir = flat.size, int(flat.sum())
        ids = np.empty(flat.size, dtype=np.int64)
        ids[flat] = np.arange(self.nDir)
        ids[~flat] = self.nDir + np.arange(flat.size - self.nDir)
        self.lattice_ids = ids.reshape(self.Gy, self.Gx)
        p = np.arange(nm)
        gx = _gi(np.arange(nx)[:, None], p[None, :], nm)  # [nx, nm]
        gy = _gi(np.arange(ny)[:, None], p[None, :], nm)
        # local index: e = ex + nx*ey, mode = p + nm*q
        l2g = self.lattice_ids[gy[:, None, :, None], gx[None, :, None, :]]  # [ey, ex, q, p]
        self.localToGlobal = np.ascontiguousarray(l2g.reshape(-1), dtype=np.int32)
        self.nLocal = self.localToGlobal.size
        self.ownerMask = np.ones(self.nGlobal)
        self.peers, self.interface_lists = [], []

    def geometry(self):
        """jac[nElmt], df[4*nElmt] of the axis-aligned rectangles (df[c*2+d] = d xi_d / d x_c)."""
        hx, hy = self.h
        jac = np.full(self.nElmt, hx * hy / 4.0)
        df = np.zeros((4, self.nElmt))
        df[0], df[3] = 2.0 / hx, 2.0 / hy
        return jac, df.reshape(-1).copy()

    def quad_coords(self, z):
        """physical coordinates of every quadrature point, [nElmt*nq^2] in [elmt][j][i] order."""
        nq = len(z)
        hx, hy = self.h
        X1 = (np.arange(self.nx)[:, None] + 0.5 * (z[None, :] + 1.0)) * hx
        Y1 = (np.arange(self.ny)[:, None] + 0.5 * (z[None, :] + 1.0)) * hy
        shape = (self.ny, self.nx, nq, nq)
        X = np.broadcast_to(X1[None, :, None, :], shape).reshape(-1)
        Y = np.broadcast_to(Y1[:, None, :, None], shape).reshape(-1)
        return np.ascontiguousarray(X), np.ascontiguousarray(Y)

    def helmholtz_diagonal(self, basis, lam):
        """matrix-free Jacobi diagonal, as StructuredHexMesh.helmholtz_diagonal"""
        nm, nq = self.nm, basis.nq
        B = basis.bdata.reshape(nm, nq)
        dB = basis.dbdata.reshape(nm, nq)
        m = (B * B) @ basis.W
        k = (dB * dB) @ basis.W
        hx, hy = self.h
        J = hx * hy / 4.0
        g0, g1 = (2.0 / hx) ** 2, (2.0 / hy) ** 2
        mq, mp = m[:, None], m[None, :]
        kq, kp = k[:, None], k[None, :]
        d = J * (lam * mq * mp + g0 * mq * kp + g1 * kq * mp)
        diag = np.zeros(self.nGlobal)
        np.add.at(diag, self.localToGlobal, np.tile(d.reshape(-1), self.nElmt))
        return diag


# ----------------------------------------------------------------------------------------------------------------
# gslib set-up from universal ids: the generic form of the partition interfaces above.
def interface_from_universal_maps(universal_maps, rank):
    """What `Gs::Init` + `Gs::Unique` give AssemblyMapCG from `m_globalToUniversalMap` (AssemblyMapCG.h:118-126; call
    sites AssemblyMapCG.cpp:2563-2569, 2840, 2922), restated for the pairwise exchange of comm.cu: `universal_maps[r]`
    holds rank r's universal id of each of its rank-local global DOFs (id 0 = not taking part, gslib's convention).
    Returns (peers, lists, ownerMask) for `rank`: the ranks it shares ids with in ascending order, for each of them
    the rank-local indices of the shared DOFs ordered by universal id (both sides of a pair build the same order), and
    the 0/1 mask that is 1 where `rank` is the lowest rank holding the id (so a masked dot product counts every DOF
    once).  An id held by k ranks appears in k-1 lists of each holder: after every holder has added what the others
    SENT (their pre-exchange values) all copies hold the sum of the k contributions -- gs_add."""
    mine = np.asarray(universal_maps[rank], dtype=np.int64)
    order = np.argsort(mine, kind="stable")
    sorted_ids = mine[order]
    if sorted_ids.size and np.any((sorted_ids[1:] == sorted_ids[:-1]) & (sorted_ids[1:] != 0)):
        raise ValueError("universal ids must be unique within a rank")
    peers, lists = [], []
    owner = np.ones(mine.size)
    for r, other in enumerate(universal_maps):
        if r == rank or other is None:  # None: a rank known to share nothing (not adjacent)
            continue
        other = np.asarray(other, dtype=np.int64)
        shared = np.intersect1d(sorted_ids[sorted_ids != 0], other[other != 0], assume_unique=True)
        if shared.size == 0:
            continue
        pos = order[np.searchsorted(sorted_ids, shared)]  # rank-local indices, ascending universal id
        peers.append(r)
        lists.append(pos.astype(np.int32))
        if r < rank:
            owner[pos] = 0.0
    return peers, lists, owner


def interface_from_universal_map(dist, universal_map):
    """the same for this process under torch.distributed: all-gathers the ranks' universal-id arrays"""
    maps = [None] * dist.get_world_size()
    dist.all_gather_object(maps, np.asarray(universal_map, dtype=np.int64))
    return interface_from_universal_maps(maps, dist.get_rank())
