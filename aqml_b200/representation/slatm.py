"""Host mirror of ``cml.representation.slatm`` (coreml/cml/representation/slatm.py:17-130,278-331)."""
from __future__ import annotations

import itertools as itl

import numpy as np

from .. import _lib as L

T, F = True, False


def get_boa(z1, zs_):
    """slatm.py:17-18."""
    return z1 * np.array([(np.asarray(zs_) == z1).sum(), ])


def _grid2(rcut, dgrid):
    return int((rcut - 0.1) / dgrid) + 1  # slatm.py:60-62


def _grid3(dgrid):
    d2r = np.pi / 180  # slatm.py:118-122
    a0 = -20.0 * d2r
    a1 = np.pi + 20.0 * d2r
    return int((a1 - a0) / dgrid) + 1


class MolPBC(object):
    """slatm.py:134-210: finite cluster of periodic images around one atom (host side, NumPy only).

    ``cell`` rows are the lattice vectors (images are ``coords + idx @ cell``); the number of images tried
    along an axis uses the column norms of ``cell`` and the fractional coordinate, exactly as upstream."""

    def __init__(self, zs, coords, cell, rcut=9.0):
        self.zs = np.asarray(zs)
        self.coords = np.asarray(coords, dtype=np.float64)
        self.cell = np.array(cell, dtype=np.float64)
        self.na = len(self.zs)
        self.coords_f = np.linalg.solve(self.cell.T, self.coords.T).T
        self.ls = np.linalg.norm(self.cell, axis=0)
        self.rcut = rcut

    def get_nmax(self, ia, axis, sign):
        n = sign
        while True:
            far = self.ls[axis] * np.abs(n - self.coords_f[ia][axis]) > self.rcut
            n += sign
            if far:
                return n

    def get_cluster(self, ia):
        lo_hi = [(self.get_nmax(ia, ax, -1), self.get_nmax(ia, ax, +1)) for ax in range(3)]
        zs = [self.zs[ia]]
        coords = [self.coords[ia]]
        for i in range(lo_hi[0][0], lo_hi[0][1] + 1):
            for j in range(lo_hi[1][0], lo_hi[1][1] + 1):
                for k in range(lo_hi[2][0], lo_hi[2][1] + 1):
                    shifted = self.coords + np.dot((i, j, k), self.cell)
                    d = np.sqrt(((shifted - self.coords[ia]) ** 2).sum(axis=1))
                    for ja in np.nonzero(d <= self.rcut)[0]:
                        if ja == ia and (i, j, k) == (0, 0, 0):
                            continue
                        zs.append(self.zs[ja])
                        coords.append(shifted[ja])
        coords = np.array(coords)
        if len(coords) > 1:
            dd = np.sqrt(((coords[:, None] - coords[None]) ** 2).sum(axis=2))
            assert np.all(dd[np.triu_indices(len(coords), 1)] > 0)  # slatm.py:209
        return [np.array(zs), coords]


def _single(fn, mbtype, obj, izeff, iloc, ia, normalize, sigma, rcut, dgrid, nx, rpower, pbc):
    zs, coords, cell = obj
    if iloc:
        assert ia is not None, '#ERROR: plz specify `za and `ia '
    if pbc:
        assert iloc, '#ERROR: for periodic system, plz use atomic rpst'  # slatm.py:43,99
        zs, coords = MolPBC(zs, coords, cell, rcut=rcut).get_cluster(ia)
        ia = 0  # slatm.py:68,125: the query atom is the first atom of the cluster
    coords = L.f64(coords)
    zs = L.i32(zs)
    if coords.shape != (len(zs), 3):
        raise ValueError(f"{coords.shape[0]} coordinates, but {len(zs)} atom_types!")  # fslatm.f90:52-57
    coeff = 1 / np.sqrt(2 * sigma ** 2 * np.pi) if normalize else 1.0
    ys = np.zeros(nx)
    iatm = int(ia) if iloc else -1
    args = [L.ptr(coords), L.ptr(zs), len(zs), iatm, int(bool(izeff))] + [int(z) for z in mbtype]
    args += [float(rcut), int(nx), float(dgrid), float(sigma), float(coeff)]
    if len(mbtype) == 2:
        args.append(float(rpower))
    L.call(fn, *args, L.ptr(ys))
    return ys


def get_sbop(mbtype, obj, cg=None, izeff=F, iloc=False, ia=None, normalize=True, sigma=0.05,
             rcut=4.8, dgrid=0.03, pbc=F, rpower=6):
    """slatm.py:22-73 -> fslatm.f90:413 calc_sbop / :531 calc_sbop_local (``cg`` is ignored upstream too)."""
    return _single("aqml_calc_sbop", mbtype, obj, izeff, iloc, ia, normalize, sigma, rcut, dgrid,
                   _grid2(rcut, dgrid), rpower, pbc)


def get_sbot(mbtype, obj, cg=None, izeff=F, iloc=F, ia=None, normalize=T, sigma=0.05,
             rcut=4.8, dgrid=0.0262, pbc=F):
    """slatm.py:75-130 -> fslatm.f90:4 calc_sbot / :206 calc_sbot_local."""
    return _single("aqml_calc_sbot", mbtype, obj, izeff, iloc, ia, normalize, sigma, rcut, dgrid,
                   _grid3(dgrid), 6, pbc)


def mbtypes_from_counts(zsu, nzmax):
    """slatm.py:286-331 / core.py:364-377: zsu ascending, nzmax[i] = max count of zsu[i] in any molecule."""
    zsu = [int(z) for z in zsu]
    boas = [[zi, ] for zi in zsu]
    bops = [[zi, zi] for zi in zsu] + [list(p) for p in itl.combinations(zsu, 2)]
    bots = []
    for i in zsu:
        for bop in bops:
            j, k = bop
            for tasi in ([i, j, k], [i, k, j], [j, i, k]):
                if (tasi not in bots) and (tasi[::-1] not in bots):
                    nzsi = [sum(1 for q in tasi if q == zj) for zj in zsu]
                    if all(n <= m for n, m in zip(nzsi, nzmax)):
                        bots.append(tasi)
    return boas + bops + bots


class NBody(object):
    """slatm.py:278-331.  ``obj`` needs ``.nzs`` (nmol, n_elements) and ``.zsu`` (ascending)."""

    def __init__(self, obj, pbc=F, rcut=4.8):
        self.obj = obj
        self.pbc = pbc
        self.rcut = rcut

    def get_slatm_mbtypes(self):
        nzmax = np.max(self.obj.nzs, axis=0)
        return mbtypes_from_counts(self.obj.zsu, nzmax)
