"""NumPy restatement of the AQML hot path (TEST INFRASTRUCTURE, not product code).

CPU oracle: a from-scratch restatement of the reference algorithm for
(1) SLATM / aSLATM generation and (2) global / local Gaussian + Laplacian kernel
assembly, the L2-distance width heuristic and the KRR learning-curve consumer.
All ``file:line`` citations are relative to the upstream reference tree
(binghuang2018/aqml).  Arithmetic is IEEE FP64 throughout; every distance is a
direct difference loop as in the Fortran (no ||x||^2+||y||^2-2xy expansion here).

Pinning: this oracle reproduces the reference's only known-answer test, the demo
learning curve printed in ``demo/example/README.md:56-63`` (see
``tests/test_oracle_kat.py`` and ``tests/golden/make_golden.py``).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import itertools
import math

import numpy as np

H2KC = 627.5094738898777  # io2/__init__.py:45

# effective nuclear charges, atomdb.f90:97-106
_ZSTAR_Z = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 35, 53]
_ZSTAR_V = [1.0, 1.688, 1.279, 1.912, 2.421, 3.136, 3.834, 4.453, 5.100, 5.758,
            2.507, 3.308, 4.066, 4.285, 4.886, 5.482, 6.116, 6.764, 9.028, 11.612]
# the Fortran literals are single precision (no d0 suffix) promoted to real*8
ZSTAR = {z: float(np.float32(v)) for z, v in zip(_ZSTAR_Z, _ZSTAR_V)}

EPS = np.finfo(np.float64).eps  # epsilon(0.0d0), fslatm.f90:37


# --------------------------------------------------------------------------------------
# many-body types
# --------------------------------------------------------------------------------------
def get_slatm_mbtypes(nuclear_charges, pbc=False):
    """Enumerate 1-/2-/3-body types.  representation/core.py:324-379, slatm.py:286-331.

    ``nuclear_charges``: list of per-molecule integer arrays.  Elements are taken in
    ascending order (np.unique in NBody; list(set) of small ints in core.py).
    """
    zsu = sorted({int(z) for zs in nuclear_charges for z in zs})
    counts = np.array([[int((np.asarray(zs) == z).sum()) for z in zsu] for zs in nuclear_charges])
    nzmax = counts.max(axis=0)
    if pbc:  # core.py:356-361
        nzmax = np.where(nzmax <= 2, 3, nzmax)
    boas = [[z] for z in zsu]
    bops = [[z, z] for z in zsu] + [list(p) for p in itertools.combinations(zsu, 2)]
    bots = []
    for i in zsu:
        for (j, k) in bops:
            for t in ([i, j, k], [i, k, j], [j, i, k]):
                if t in bots or t[::-1] in bots:
                    continue
                mult = [sum(1 for q in t if q == z) for z in zsu]
                if all(m <= n for m, n in zip(mult, nzmax)):
                    bots.append(t)
    return boas + bops + bots


# --------------------------------------------------------------------------------------
# grids
# --------------------------------------------------------------------------------------
def linspace(x0, x1, nx):
    """utils.f90:29-50: xs(i) = x0 + (i-1)*step, step=(x1-x0)/(nx-1)."""
    step = (x1 - x0) / (nx - 1)
    return x0 + np.arange(nx, dtype=np.float64) * step


def grid2(rcut, dgrid):
    """slatm.py:60-62."""
    r0 = 0.1
    nx = int((rcut - r0) / dgrid) + 1
    return nx


def grid3(dgrid):
    """slatm.py:118-122."""
    d2r = np.pi / 180
    a0 = -20.0 * d2r
    a1 = np.pi + 20.0 * d2r
    nx = int((a1 - a0) / dgrid) + 1
    return nx


def _weight(z, izeff):
    return ZSTAR[z] if izeff else float(z % 1000)


# --------------------------------------------------------------------------------------
# two-body (London) spectrum: fslatm.f90:413-527 (global) and :531-655 (local)
# --------------------------------------------------------------------------------------
def calc_sbop(coords, zs, z1, z2, rcut, nx, dgrid, sigma, coeff, rpower, izeff=False, ia=None):
    """``ia`` None -> calc_sbop (global); else calc_sbop_local for 0-based atom ``ia``."""
    coords = np.asarray(coords, dtype=np.float64)
    zi = np.asarray(zs).astype(np.int64)
    ias1 = np.nonzero(zi == z1)[0]
    ias2 = np.nonzero(zi == z2)[0]
    xs = linspace(0.1, rcut, nx)
    ys = np.zeros(nx)
    pref = 0.5 if ia is not None else 1.0  # local: all(cg==1) -> 1/2, :610-611
    c0 = pref * (_weight(z1, izeff) * _weight(z2, izeff)) * coeff
    inv_sigma = -0.5 / sigma ** 2
    xs0 = c0 / (xs ** rpower) * dgrid
    rcut2 = rcut ** 2
    same = (z1 == z2)
    for p1, i in enumerate(ias1):
        start = p1 + 1 if same else 0
        for j in ias2[start:]:
            if ia is not None and (i != ia and j != ia):
                continue
            d = coords[i] - coords[j]
            r = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
            if r < rcut2:
                ys += xs0 * np.exp(inv_sigma * (xs - math.sqrt(r)) ** 2)
    return ys


# --------------------------------------------------------------------------------------
# three-body (Axilrod-Teller-Muto) spectrum: fslatm.f90:4-202 (global), :206-409 (local)
# --------------------------------------------------------------------------------------
def _unit(v):
    return v / math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def _cos_angle(a, b, c):
    """utils.f90:82-103: cosine of the angle at b."""
    v1 = _unit(a - b)
    v2 = _unit(c - b)
    return v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2]


def _angle(a, b, c):
    """utils.f90:52-80: clipped acos."""
    ca = _cos_angle(a, b, c)
    ca = min(1.0, max(-1.0, ca))
    return math.acos(ca)


def calc_sbot(coords, zs, z1, z2, z3, rcut, nx, dgrid, sigma, coeff, izeff=False, ia=None,
              shipped_index_slip=False):
    """``ia`` None -> calc_sbot (global, centre loops over all Z==z2); else calc_sbot_local.

    ``shipped_index_slip``: reproduce fslatm.f90:174-175 (global, z1!=z3 tests
    ``ias1(ia2)`` instead of ``ias1(ia1)``); only defined when that index is in range.
    Default False = intended semantics (Trash/fslatm_0.f90:263 and the local routine).
    """
    coords = np.asarray(coords, dtype=np.float64)
    zi = np.asarray(zs).astype(np.int64)
    na = len(zi)
    dm = np.zeros((na, na))
    for i in range(na):
        for j in range(i + 1, na):
            d = coords[j] - coords[i]
            dm[i, j] = dm[j, i] = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
    ias1 = np.nonzero(zi == z1)[0]
    ias2 = np.nonzero(zi == z2)[0]
    ias3 = np.nonzero(zi == z3)[0]
    ys = np.zeros(nx)
    if ia is not None and ia not in ias2:  # :312-316
        return ys
    d2r = np.pi / 180.0
    a0 = -20.0 * d2r
    a1 = np.pi + 20.0 * d2r
    xs = linspace(a0, a1, nx)
    prefactor = 1.0 / 3.0  # all(cg==1), :112-113/:324-325
    c0 = prefactor * (_weight(z1, izeff) * _weight(z2, izeff) * _weight(z3, izeff)) * coeff * dgrid
    inv_sigma = -1.0 / (2 * sigma ** 2)
    cos_xs = np.cos(xs) * c0
    centres = ias2 if ia is None else [ia]
    same = (z1 == z3)
    for p1, i in enumerate(ias1):
        for p2, j in enumerate(centres):
            dij = dm[i, j]
            if not (dij > EPS and dij <= rcut):
                continue
            ks = ias3[p1 + 1:] if same else ias3
            for k in ks:
                if shipped_index_slip and (ia is None) and (not same):
                    dik = dm[ias1[p2], k]  # fslatm.f90:175 (raises IndexError where Fortran is OOB)
                else:
                    dik = dm[i, k]
                if dik > rcut:
                    continue
                dkj = dm[j, k]
                if not (dkj > EPS and dkj <= rcut):
                    continue
                ang = _angle(coords[i], coords[j], coords[k])
                cak = _cos_angle(coords[i], coords[k], coords[j])
                cai = _cos_angle(coords[k], coords[i], coords[j])
                r = dm[i, j] * dm[i, k] * dm[k, j]
                ys += (c0 + cos_xs * cak * cai) / (r ** 3) * np.exp((xs - ang) ** 2 * inv_sigma)
    return ys


# --------------------------------------------------------------------------------------
# Python-level wrappers, representation/slatm.py:17-130, core.py:382-556
# --------------------------------------------------------------------------------------
def get_boa(z1, zs_):
    return z1 * np.array([(np.asarray(zs_) == z1).sum(), ])


def get_sbop(mbtype, obj, izeff=False, iloc=False, ia=None, normalize=True, sigma=0.05,
             rcut=4.8, dgrid=0.03, rpower=6):
    z1, z2 = mbtype
    zs, coords, _ = obj
    nx = grid2(rcut, dgrid)
    coeff = 1 / np.sqrt(2 * sigma ** 2 * np.pi) if normalize else 1.0
    return calc_sbop(coords, zs, z1, z2, rcut, nx, dgrid, sigma, coeff, rpower, izeff,
                     ia=(ia if iloc else None))


def get_sbot(mbtype, obj, izeff=False, iloc=False, ia=None, normalize=True, sigma=0.05,
             rcut=4.8, dgrid=0.0262, shipped_index_slip=False):
    z1, z2, z3 = mbtype
    zs, coords, _ = obj
    nx = grid3(dgrid)
    coeff = 1 / np.sqrt(2 * sigma ** 2 * np.pi) if normalize else 1.0
    return calc_sbot(coords, zs, z1, z2, z3, rcut, nx, dgrid, sigma, coeff, izeff,
                     ia=(ia if iloc else None), shipped_index_slip=shipped_index_slip)


def generate_slatm(coordinates, nuclear_charges, mbtypes, izeff=False, local=False,
                   sigmas=(0.05, 0.05), dgrids=(0.03, 0.03), rcut=4.8, rpower=6, ias=None,
                   shipped_index_slip=False):
    """core.py:382-556 without the alchemy / pbc branches."""
    zs = np.asarray(nuclear_charges)
    coords = np.asarray(coordinates, dtype=np.float64)
    obj = [zs, coords, None]
    if local:
        out = []
        for ia in (range(len(zs)) if ias is None else ias):
            parts = []
            for mb in mbtypes:
                if len(mb) == 1:
                    parts.append(get_boa(mb[0], np.array([zs[ia]])).astype(np.float64))
                elif len(mb) == 2:
                    parts.append(get_sbop(mb, obj, izeff=izeff, iloc=True, ia=ia, sigma=sigmas[0],
                                          dgrid=dgrids[0], rcut=rcut, rpower=rpower))
                else:
                    parts.append(get_sbot(mb, obj, izeff=izeff, iloc=True, ia=ia, sigma=sigmas[1],
                                          dgrid=dgrids[1], rcut=rcut))
            out.append(np.concatenate(parts))
        return out
    parts = []
    for mb in mbtypes:
        if len(mb) == 1:
            parts.append(get_boa(mb[0], zs).astype(np.float64))
        elif len(mb) == 2:
            parts.append(get_sbop(mb, obj, izeff=izeff, sigma=sigmas[0], dgrid=dgrids[0], rcut=rcut,
                                  rpower=rpower))
        else:
            parts.append(get_sbot(mb, obj, izeff=izeff, sigma=sigmas[1], dgrid=dgrids[1], rcut=rcut,
                                  shipped_index_slip=shipped_index_slip))
    return np.concatenate(parts)


# --------------------------------------------------------------------------------------
# global kernels / distances: fkernels.f90:327-387, fdistance.f90:25-44
# --------------------------------------------------------------------------------------
def _rowblocks(n, blk=256):
    for s in range(0, n, blk):
        yield slice(s, min(n, s + blk))


def fl2_distance(A, B):
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    D = np.empty((A.shape[0], B.shape[0]))
    for j in range(A.shape[0]):
        t = A[j][None, :] - B
        D[j] = np.sqrt(np.sum(t * t, axis=1))
    return D


def fl1_distance(A, B):
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    D = np.empty((A.shape[0], B.shape[0]))
    for j in range(A.shape[0]):
        D[j] = np.sum(np.abs(A[j][None, :] - B), axis=1)
    return D


def gaussian_kernel(A, B, sigma):
    """kernels.py:40-67 -> fkernels.f90:327-359.  Returns (N1,N2)."""
    d = fl2_distance(A, B)
    return np.exp((-0.5 / (sigma * sigma)) * (d * d))


def gaussian_kernel_direct(A, B, sigma):
    """Same, without the sqrt round trip (sum of squared differences directly)."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    K = np.empty((A.shape[0], B.shape[0]))
    inv = -0.5 / (sigma * sigma)
    for j in range(A.shape[0]):
        t = A[j][None, :] - B
        K[j] = np.exp(inv * np.sum(t * t, axis=1))
    return K


def laplacian_kernel(A, B, sigma):
    """kernels.py:11-38 -> fkernels.f90:361-387."""
    return np.exp((-1.0 / sigma) * fl1_distance(A, B))


# --------------------------------------------------------------------------------------
# local kernels: fkernels.f90:23-93 and :97-183
# --------------------------------------------------------------------------------------
def calc_lk_gaussian(x1, x2, zs1, zs2, nas1, nas2, zmax, iab, sigmas):
    """K[k,a,b] = sum_{i in a, j in b} exp(d2_ij * (-0.5/sig[k, Z1_i-1]^2)),
    d2 = |x1_i-x2_j|^2 if (iab or Z1_i==Z2_j) else 1e9 (single-precision literal 1.e9)."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    zs1 = np.asarray(zs1).astype(np.int64)
    zs2 = np.asarray(zs2).astype(np.int64)
    nas1 = np.asarray(nas1).astype(np.int64)
    nas2 = np.asarray(nas2).astype(np.int64)
    sigmas = np.asarray(sigmas, dtype=np.float64).reshape(len(sigmas), -1)
    S = sigmas.shape[0]
    inv_sigma2 = -0.5 / sigmas ** 2
    ias = np.concatenate(([0], np.cumsum(nas1)))
    jas = np.concatenate(([0], np.cumsum(nas2)))
    # atom-pair values, then molecule-pair sums
    E = np.zeros((S, x1.shape[0], x2.shape[0]))
    for i in range(x1.shape[0]):
        t = x1[i][None, :] - x2
        d2 = np.sum(t * t, axis=1)
        if not iab:
            d2 = np.where(zs2 == zs1[i], d2, np.float64(np.float32(1.e9)))
        E[:, i, :] = np.exp(d2[None, :] * inv_sigma2[:, zs1[i] - 1][:, None])
    K = np.zeros((S, len(nas1), len(nas2)))
    for a in range(len(nas1)):
        Ea = E[:, ias[a]:ias[a + 1], :].sum(axis=1)
        for b in range(len(nas2)):
            K[:, a, b] = Ea[:, jas[b]:jas[b + 1]].sum(axis=1)
    return K


def calc_lk_laplacian(x1, x2, nas1, nas2, sigmas):
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    nas1 = np.asarray(nas1).astype(np.int64)
    nas2 = np.asarray(nas2).astype(np.int64)
    sigmas = np.asarray(sigmas, dtype=np.float64).ravel()
    S = len(sigmas)
    inv = -1.0 / sigmas
    ias = np.concatenate(([0], np.cumsum(nas1)))
    jas = np.concatenate(([0], np.cumsum(nas2)))
    D = fl1_distance(x1, x2)
    E = np.exp(D[None, :, :] * inv[:, None, None])
    K = np.zeros((S, len(nas1), len(nas2)))
    for a in range(len(nas1)):
        Ea = E[:, ias[a]:ias[a + 1], :].sum(axis=1)
        for b in range(len(nas2)):
            K[:, a, b] = Ea[:, jas[b]:jas[b + 1]].sum(axis=1)
    return K


# --------------------------------------------------------------------------------------
# kernel width heuristic: algo/aqml.py:204-255
# --------------------------------------------------------------------------------------
def get_kernel_width(nas, zs, x, cab=False, kernel='g'):
    zs = np.asarray(zs).astype(np.int64)
    x = np.asarray(x, dtype=np.float64)
    zsu = np.unique(zs)
    dist = fl2_distance if kernel[0] == 'g' else fl1_distance
    Nz = len(zsu)
    zmax = zsu[-1]
    if Nz == 1:
        dsmax = np.array([np.max(dist(x, x))] * Nz)
    elif cab:
        f1 = zs == zsu[-2]
        f2 = zs == zsu[-1]
        dsmax = np.array([np.max(dist(x[f1], x[f2]))] * Nz)
    else:
        dsmax = np.array([np.max(dist(x[zs == z], x[zs == z])) for z in zsu])
    dso = np.ones(zmax) * 1.e-9
    for i, z in enumerate(zsu):
        dso[z - 1] = dsmax[i]
    return dso


def width_factor(kernel='g'):
    """aqml.py:193-199."""
    return 1. / np.sqrt(2.0 * np.log(2.0)) if kernel[0] == 'g' else 1. / np.log(2.0)


# --------------------------------------------------------------------------------------
# KRR consumer: aqml.py:129-181, 799-897; math/matrix.py
# --------------------------------------------------------------------------------------
# free-atom energies (Ha) used by the rank-deficient fallback, cheminfo/data/atoms.py:206-214
ORCA_B3LYPVDZ = {1: -0.497858961664, 6: -37.830625604890, 7: -54.564148277678,
                 8: -75.039064184306, 9: -99.692719367701}


def _indep_cols(M):
    """math/matrix.py:18-33."""
    mr = np.linalg.matrix_rank
    idxc = [0]
    for i in range(1, M.shape[1]):
        if mr(M[:, idxc + [i]]) > mr(M[:, idxc]):
            idxc.append(i)
    return idxc


def calc_ae_dressed(ngs, ys, idx1, idx2, zsu, free_atom=ORCA_B3LYPVDZ):
    """aqml.py:129-181 (``ref is None`` branch)."""
    ns1 = ngs[idx1]
    ns2 = ngs[idx2]
    ns = np.concatenate((ns1, ns2), axis=0)
    nel = ns1.shape[1]
    ys1, ys2 = ys[idx1], ys[idx2]
    mr = np.linalg.matrix_rank
    if mr(ns1) < nel:
        ok = (mr(ns1) == mr(ns)) and (_indep_cols(ns1) == _indep_cols(ns))
        if ok:
            cols = _indep_cols(ns)
            ns1 = ns1[:, cols]
            ns2 = ns2[:, cols]
            esb = np.linalg.lstsq(ns1, ys1, rcond=None)[0]
        else:
            esb = np.array([H2KC * free_atom[int(z)] for z in zsu])
    else:
        esb = np.linalg.lstsq(ns1, ys1, rcond=None)[0]
    dys1 = ys1 - np.dot(ns1, esb)
    dys2 = ys2 - np.dot(ns2, esb)
    return dys1, dys2, esb


def learning_curve(ks1, ks2, ngs, ys, nheav1, zsu, llambda=1e-4):
    """aqml.py:806-897 for one coeff / one lambda, single or multiple test molecules.

    ks1 (N1,N1), ks2 (N2,N1); ys in kcal/mol for train then test; returns rows
    (ni, n_train, mae, signed-or-rmse, -ys2[-1] (the 'DressedAtom' column))."""
    n1 = ks1.shape[0]
    n2 = ks2.shape[0]
    ims1 = np.arange(n1)
    idx2 = np.arange(n1, n1 + n2)
    nheav1 = np.asarray(nheav1)
    rows = []
    for ni in range(1, 1 + nheav1.max()):
        idx1 = ims1[nheav1 <= ni]
        if len(idx1) == 0:
            continue
        k1 = ks1[idx1][:, idx1].copy()
        k2 = ks2[:, idx1].copy()
        ys1, ys2, esb = calc_ae_dressed(ngs, ys, idx1, idx2, zsu)
        k1[np.diag_indices_from(k1)] += llambda
        alphas = np.linalg.solve(k1, ys1)
        ys2p = np.dot(k2, alphas)
        dys2 = ys2p - ys2
        mae = np.sum(np.abs(dys2)) / len(dys2)
        rows.append((ni, len(idx1), mae, dys2[-1], -ys2[-1], ys2p.copy()))
    return rows



def calc_ka(x2, x1, zs2, zs1, nas2, nas1, zmax, sigmas, kernel='g'):
    """aqml.py:258-309 (cab=False, zaq=None): atomic kernel matrix, summed over the atoms of each training molecule.
    Returns ks (nsigmas, A2, N1): ks[k, j, m] = sum_{i in molecule m, Z_i == Z_j} exp(-0.5 (d_ij / sigmas[k, Z_j-1])^2)."""
    zs1, zs2 = np.asarray(zs1), np.asarray(zs2)
    nsg, na1, na2, n1 = len(sigmas), len(zs1), len(zs2), len(nas1)
    ksa = np.zeros((nsg, na2, na1))
    for zi in np.unique(zs2):
        ias1 = np.nonzero(zs1 == zi)[0]
        ias2 = np.nonzero(zs2 == zi)[0]
        if len(ias1) == 0:
            continue
        ds = fl2_distance(x2[ias2], x1[ias1]) if kernel[0] == 'g' else fl1_distance(x2[ias2], x1[ias1])
        for k in range(nsg):
            ksa[k][np.ix_(ias2, ias1)] = np.exp(-0.5 * (ds / sigmas[k][zi - 1]) ** 2)
    off = np.concatenate(([0], np.cumsum(nas1)))
    return np.stack([ksa[:, :, off[i]:off[i + 1]].sum(axis=2) for i in range(n1)], axis=2)


def atomic_energies(ksa21, alphas, zs2, zsu, esb, free_atom=ORCA_B3LYPVDZ):
    """aqml.py:903-930 (``-ieaq -prog orca``): per-atom contribution of the query molecule,
    E_A[j] = (ksa21 . alphas)[j] + esb[Z_j] - H2KC * E_free_atom[Z_j]   (kcal/mol).
    ksa21 (A2, N1) for one sigma, alphas (N1,), esb = dressed-atom energies in the order of zsu."""
    easq = np.dot(ksa21, alphas)
    edct = dict(zip([int(z) for z in zsu], esb))
    return np.array([easq[j] + edct[int(z)] - H2KC * free_atom[int(z)] for j, z in enumerate(zs2)]), easq


def read_xyz(path, prop='b3lypvdz'):
    """cheminfo/rw/xyz.py:10-80 (geometry + one 'key=value' property on line 2)."""
    sym2z = {'H': 1, 'C': 6, 'N': 7, 'O': 8, 'F': 9, 'P': 15, 'S': 16, 'Cl': 17}
    with open(path) as fh:
        lines = fh.readlines()
    na = int(lines[0])
    props = {}
    for tok in lines[1].split():
        if '=' in tok:
            k, v = tok.split('=')
            props[k] = float(v)
    zs, coords = [], []
    for ln in lines[2:2 + na]:
        s = ln.split()
        zs.append(sym2z[s[0]])
        coords.append([float(s[1]), float(s[2]), float(s[3])])
    return np.array(zs, dtype=np.int64), np.array(coords, dtype=np.float64), props.get(prop)


# --------------------------------------------------------------------------------------
# periodic systems: representation/slatm.py:134-210 (MolPBC.get_cluster)
# --------------------------------------------------------------------------------------
def pbc_cluster(zs, coords, cell, ia, rcut):
    """Finite cluster around atom ia: the atom itself first, then every periodic image within rcut
    (images indexed by integer triples whose range per axis follows get_nmax, slatm.py:166-181)."""
    zs = np.asarray(zs)
    coords = np.asarray(coords, dtype=np.float64)
    cell = np.array(cell, dtype=np.float64)
    frac = np.linalg.solve(cell.T, coords.T).T
    ls = np.linalg.norm(cell, axis=0)

    def nmax(axis, sign):
        n = sign
        while True:
            if ls[axis] * abs(n - frac[ia][axis]) > rcut:
                return n + sign
            n += sign

    rng = [range(nmax(ax, -1), nmax(ax, 1) + 1) for ax in range(3)]
    out_z, out_x = [zs[ia]], [coords[ia]]
    for idx in itertools.product(*rng):
        img = coords + np.dot(idx, cell)
        ds = np.sqrt(((img - coords[ia]) ** 2).sum(axis=1))
        for ja in range(len(zs)):
            if ds[ja] <= rcut and not (ja == ia and idx == (0, 0, 0)):
                out_z.append(zs[ja])
                out_x.append(img[ja])
    return np.array(out_z), np.array(out_x)


def generate_slatm_pbc(coordinates, nuclear_charges, cell, mbtypes, **kw):
    """generate_slatm(..., local=True, pbc=True): row of the query atom in its own cluster."""
    rows = []
    rcut = kw.get('rcut', 4.8)
    for ia in range(len(nuclear_charges)):
        z, x = pbc_cluster(nuclear_charges, coordinates, cell, ia, rcut)
        rows.append(generate_slatm(x, z, mbtypes, local=True, ias=[0], **kw)[0])
    return rows
