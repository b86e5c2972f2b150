"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the pair-wise / bond aSLATM of the reference:
coreml/cml/representation/fslatm_pairwise.f90 (atom rows + bond rows of `calc_local_spectrum`, the molecular kernels
`calc_mk1` / `calc_mk2` and `calc_dij_max` in their `local` form) and the helpers of its driver representation/xb.py.

PARITY UNPINNED.  The reference holds no test, fixture or known answer for this file, the file is not part of the
reference's build (coreml/setup.py:43-47) and as shipped it cannot run:
  * `calc_sbop_local` allocates `ias2(na)` with `na` never assigned (fslatm_pairwise.f90:394,403);
  * `calc_spectrum` (the global form) passes `ys(jb:je)` slices that are one element too long and shifted (:457-471).
This oracle therefore DEFINES the semantics it checks the CUDA path against: the statements of the file taken literally,
with `na := size(zs)` in `calc_sbop_local`.  The global forms (`calc_spectrum`, `local = .false.` in calc_mk1/mk2/dij_max)
are not restated.  Everything else follows the file line by line, including its quirks:
  * `r0 = 0.1` is a single-precision literal stored in a real*8 (:5) -> 0.100000001490116;
  * the bond two-body grid has nsx(2) points but still ends at rscut(1) (:418), its cut-off is rscut(2) (:409);
  * three-body terms are summed over ALL ordered (ia1, ia3) combinations -- for z1 = z3 every neighbour pair enters twice (:264-292);
  * the one-body slot of an atom is its nuclear charge itself (:582);
  * `calc_mk1` builds the second molecule's spectrum with the FIRST molecule's neighbour table (:803) -- harmless, the bond rows
    of that call are discarded.
Only tests/ may import this module.
"""
from __future__ import annotations

import itertools as itl

import numpy as np

R0 = float(np.float32(0.1))        # fslatm_pairwise.f90:5
EPS = float(np.finfo(np.float64).eps)
PI = 4.0 * np.arctan(1.0)


def linspace(x0, x1, nx):          # :12-28
    step = (x1 - x0) / (nx - 1)
    return x0 + np.arange(nx) * step


def _unit(v):
    return v / np.sqrt(np.sum(v * v))


def calc_angle(a, b, c):           # :30-48
    cosang = float(np.dot(_unit(a - b), _unit(c - b)))
    return np.arccos(min(1.0, max(-1.0, cosang)))


def calc_cos_angle(a, b, c):       # :51-63
    return float(np.dot(_unit(a - b), _unit(c - b)))


def grid_sizes(racut, rbcut, dgrid):
    """nsx of xb.py:110-124 / calc_mk1 :781-788."""
    return [int((racut - R0) / dgrid) + 1, int((rbcut - R0) / dgrid) + 1, int((PI + 40.0 * (PI / 180.0)) / dgrid) + 1]


def get_slatm_mbtypes(zs, nas):
    """xb.py:72-95 (pbc = '000')."""
    zs, nas = np.asarray(zs, dtype=int), np.asarray(nas, dtype=int)
    zsu = np.unique(zs)
    off = np.concatenate(([0], np.cumsum(nas)))
    nzs = np.array([[(zs[off[i]:off[i + 1]] == z).sum() for z in zsu] for i in range(len(nas))])
    nzsmax = nzs.max(axis=0)
    boas = [[int(z)] for z in zsu]
    bops = [[int(z), int(z)] for z in zsu] + [list(map(int, p)) for p in itl.combinations(zsu, 2)]
    bots = []
    for i in zsu:
        for j, k in bops:
            for t in ([int(i), j, k], [int(i), k, j], [j, int(i), k]):
                if t not in bots and t[::-1] not in bots:
                    if all((np.array(t) == z).sum() <= m for z, m in zip(zsu, nzsmax)):
                        bots.append(t)
    return np.array(boas, dtype=int).reshape(-1), np.array(bops, dtype=int), np.array(bots, dtype=int).reshape(-1, 3)


def get_neighbors(rbcut, coords):
    """xb.py:36-51: atoms within rbcut of each atom (itself included), -1 padded."""
    coords = np.asarray(coords, dtype=float)
    ds = np.sqrt(((coords[:, None, :] - coords[None, :, :]) ** 2).sum(-1))
    lists = [np.nonzero(ds[i] <= rbcut)[0] for i in range(len(coords))]
    nmax = max(len(l) for l in lists)
    nbrs = -np.ones((len(coords), nmax), dtype=int)
    for i, l in enumerate(lists):
        nbrs[i, :len(l)] = l
    return nmax, nbrs


def calc_sbop_local(coords, zs, ia, ib, z1, z2, rscut, nx, dgrid, sigma, w2, rpower2):
    """:372-434.  ia, ib: 0-based here; ib < 0 = all partners (the Fortran's ib <= 0)."""
    ys = np.zeros(nx)
    na = len(zs)
    if zs[ia] != z1:
        return ys
    if ib >= 0 and (ib == ia or zs[ib] != z2):
        return ys
    if ib >= 0:
        partners, racut = [ib], rscut[1]
    else:
        partners, racut = [i for i in range(na) if zs[i] == z2 and i != ia], rscut[0]
    xs = linspace(R0, rscut[0], nx)
    c0 = float((z1 % 1000) * (z2 % 1000)) / np.sqrt(2 * sigma ** 2 * PI)
    inv_sigma = -0.5 / sigma ** 2
    xs0 = w2 * c0 / (xs ** rpower2) * dgrid
    for j in partners:
        r = float(np.sum((coords[ia] - coords[j]) ** 2))
        if r < racut ** 2:
            ys = ys + xs0 * np.exp(inv_sigma * (xs - np.sqrt(r)) ** 2)
    return ys


def calc_sbot_local(coords, zs, ia, ib, z1, z2, z3, rscut, nx, dgrid, sigma, w3, rpower3):
    """:191-301."""
    ys = np.zeros(nx)
    na = len(zs)
    if zs[ia] != z2:
        return ys
    if ib >= 0 and zs[ib] != z1:
        return ys
    ds = np.sqrt(((coords[:, None, :] - coords[None, :, :]) ** 2).sum(-1))
    if ib >= 0:
        if ib == ia:
            return ys
        ias1 = [ib]
    else:
        ias1 = [i for i in range(na) if zs[i] == z1 and i != ia]
    ias3 = [i for i in range(na) if zs[i] == z3 and i != ia]
    xs = linspace(-20.0 * PI / 180.0, PI + 20.0 * PI / 180.0, nx)
    c0 = 1.0 * ((z1 % 1000) * (z2 % 1000) * (z3 % 1000)) * dgrid
    c0 = w3 * c0 / np.sqrt(2 * sigma ** 2 * PI)
    inv_sigma = -1.0 / (2 * sigma ** 2)
    cos_xs = np.cos(xs) * c0
    cut1 = rscut[1] if ib >= 0 else rscut[0]
    for i in ias1:
        if not (ds[i, ia] > EPS and ds[i, ia] <= cut1):
            continue
        for k in ias3:
            if k == ia or k == i:
                continue
            if not (ds[ia, k] > EPS and ds[ia, k] <= rscut[0]):
                continue
            ang = calc_angle(coords[i], coords[ia], coords[k])
            cak = calc_cos_angle(coords[i], coords[k], coords[ia])
            cai = calc_cos_angle(coords[k], coords[i], coords[ia])
            r = ds[i, ia] * ds[i, k] * ds[k, ia]
            ys = ys + (c0 + cos_xs * cak * cai) / (r ** rpower3) * np.exp((xs - ang) ** 2 * inv_sigma)
    return ys


def calc_local_spectrum(zs, coords, nbrs, mbs1, mbs2, mbs3, nsx, rscut, dgrid, sigma, w2, rpower2, w3, rpower3, bonds=True):
    """:525-641 with ia0 = 0 (all atoms) and ib0 = 0 (all listed neighbours) or -1 (bonds=False: no bond rows).
    Returns ys (na, n) and ys2 (na, nbmax, nu)."""
    zs, coords, nbrs = np.asarray(zs, dtype=int), np.asarray(coords, dtype=float), np.asarray(nbrs, dtype=int)
    na, nbmax = len(zs), nbrs.shape[1]
    n1, n2, n3 = len(mbs1), len(mbs2), len(mbs3)
    n = n1 + n2 * nsx[0] + n3 * nsx[2]
    nu = n1 + n2 * nsx[1] + n3 * nsx[2]
    ys, ys2 = np.zeros((na, n)), np.zeros((na, nbmax, nu))
    for ia in range(na):
        for m, z1 in enumerate(mbs1):
            if z1 == zs[ia]:
                ys[ia, m] = z1
        for m, (z1, z2) in enumerate(mbs2):
            b = n1 + m * nsx[0]
            ys[ia, b:b + nsx[0]] = calc_sbop_local(coords, zs, ia, -1, z1, z2, rscut, nsx[0], dgrid, sigma, w2, rpower2)
        for m, (z1, z2, z3) in enumerate(mbs3):
            b = n1 + n2 * nsx[0] + m * nsx[2]
            ys[ia, b:b + nsx[2]] = calc_sbot_local(coords, zs, ia, -1, z1, z2, z3, rscut, nsx[2], dgrid, sigma, w3, rpower3)
    if bonds:
        for ia in range(na):
            cnt = int((nbrs[ia] > -1).sum())                 # nas2(ia) = count(nbrs(:,ia) > -1); the first `cnt` slots are walked (:605)
            for j in range(cnt):
                ja = int(nbrs[ia, j])
                if ja == ia or ja < 0:
                    continue
                for m, (z1, z2) in enumerate(mbs2):
                    b = n1 + m * nsx[1]
                    ys2[ia, j, b:b + nsx[1]] = calc_sbop_local(coords, zs, ia, ja, z1, z2, rscut, nsx[1], dgrid, sigma, w2, rpower2)
                for m, (z1, z2, z3) in enumerate(mbs3):
                    b = n1 + n2 * nsx[1] + m * nsx[2]
                    ys2[ia, j, b:b + nsx[2]] = calc_sbot_local(coords, zs, ia, ja, z1, z2, z3, rscut, nsx[2], dgrid, sigma, w3, rpower3)
    return ys, ys2


def _atom_rows(nas, zs, coords, mb, rscut, dgrid, sigma, w2, rpower2, w3, rpower3):
    nsx = grid_sizes(rscut[0], rscut[1], dgrid)
    off = np.concatenate(([0], np.cumsum(nas)))
    rows = []
    for i in range(len(nas)):
        sl = slice(off[i], off[i + 1])
        dummy = -np.ones((off[i + 1] - off[i], 1), dtype=int)
        rows.append(calc_local_spectrum(zs[sl], coords[sl], dummy, mb[0], mb[1], mb[2], nsx, rscut, dgrid, sigma, w2, rpower2,
                                        w3, rpower3, bonds=False)[0])
    return rows


def calc_local_kij(x1, x2, kwds):
    """:695-718: sum over ALL atom pairs (no element matching) of exp(-|x1_i - x2_j|^2 / (2 kwd^2))."""
    d2 = ((x1[:, None, :] - x2[None, :, :]) ** 2).sum(-1)
    return np.array([np.exp(-d2 / (2.0 * k ** 2)).sum() for k in kwds])


def calc_mk1(nas, zs, coords, mb, rscut, dgrid, sigma, w2, rpower2, w3, rpower3, kwds):
    """:740-829, local = .true.: k1 (nco, nm, nm)."""
    rows = _atom_rows(nas, np.asarray(zs, dtype=int), np.asarray(coords, dtype=float), mb, rscut, dgrid, sigma, w2, rpower2, w3, rpower3)
    nm = len(nas)
    k1 = np.zeros((len(kwds), nm, nm))
    for i in range(nm):
        for j in range(i, nm):
            k1[:, i, j] = k1[:, j, i] = calc_local_kij(rows[i], rows[j], kwds)
    return k1


def calc_mk2(nm1, nas, zs, coords, mb, rscut, dgrid, sigma, w2, rpower2, w3, rpower3, kwds):
    """:832-917, local = .true.: k2 (nco, nm - nm1, nm1), k2(:, j, i) for training molecule i < nm1 and test molecule j >= nm1."""
    rows = _atom_rows(nas, np.asarray(zs, dtype=int), np.asarray(coords, dtype=float), mb, rscut, dgrid, sigma, w2, rpower2, w3, rpower3)
    nm = len(nas)
    k2 = np.zeros((len(kwds), nm - nm1, nm1))
    for i in range(nm1):
        for j in range(nm1, nm):
            k2[:, j - nm1, i] = calc_local_kij(rows[i], rows[j], kwds)
    return k2


def calc_dij_max(nas, zs, coords, mb, rscut, dgrid, sigma, w2, rpower2, w3, rpower3):
    """:920-1005, local = .true.: the largest atom-atom distance of the representation over all molecule pairs."""
    rows = _atom_rows(nas, np.asarray(zs, dtype=int), np.asarray(coords, dtype=float), mb, rscut, dgrid, sigma, w2, rpower2, w3, rpower3)
    dmax = 0.0
    for i in range(len(nas)):
        for j in range(i, len(nas)):
            d2 = ((rows[i][:, None, :] - rows[j][None, :, :]) ** 2).sum(-1)
            dmax = max(dmax, float(np.sqrt(d2.max())))
    return dmax
