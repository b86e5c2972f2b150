"""Pair-wise / bond aSLATM and its molecular kernels on the GPU -- the reference's
coreml/cml/representation/fslatm_pairwise.f90 (`calc_local_spectrum` :525-641, `calc_all_local_spectrum` :644-673,
`calc_mk1` :740-829, `calc_mk2` :832-917, `calc_dij_max` :920-1005, `local = .true.`) and the host helpers of its driver
representation/xb.py (`get_slatm_mbtypes` :72-95, `get_neighbors` :36-51, `init_grids` :110-124).

The file is outside the reference's build and cannot run as shipped; semantics as defined in oracle/np_pairwise.py
(PARITY UNPINNED).  The spectra come from `aqml_pairwise_local_spectrum_dev` (csrc/pairwise.cu); the molecular kernels
K(i, j) = sum over ALL atom pairs exp(-|x_I - x_J|^2 / (2 kwd^2)) are the local-kernel contraction with one pseudo element
(`aqml_calc_lk_gaussian_dev`), the largest distance is the max-reduction pass (`aqml_max_distance_dev`): the reference
rebuilds both molecules' spectra for every molecule pair, here every spectrum is computed once.
The global forms (`local = .false.`, `calc_spectrum`) are not provided.
"""
from __future__ import annotations

import ctypes as C
import itertools as itl

import numpy as np

from .. import _lib as L


def get_slatm_mbtypes(zs, nas):
    """xb.py:72-95 (pbc = '000'): (mbs1 (n1,), mbs2 (n2, 2), mbs3 (n3, 3))."""
    zs, nas = np.asarray(zs, dtype=int), np.asarray(nas, dtype=int)
    zsu = np.unique(zs)
    off = np.concatenate(([0], np.cumsum(nas)))
    nzsmax = np.max([[np.count_nonzero(zs[off[i]:off[i + 1]] == z) for z in zsu] for i in range(len(nas))], axis=0)
    bops = [[int(z), int(z)] for z in zsu] + [[int(a), int(b)] for a, b in itl.combinations(zsu, 2)]
    bots = []
    for i in (int(z) for z in zsu):
        for j, k in bops:
            for t in ([i, j, k], [i, k, j], [j, i, k]):
                if t in bots or t[::-1] in bots:
                    continue
                if all(t.count(int(z)) <= m for z, m in zip(zsu, nzsmax)):
                    bots.append(t)
    return (np.array([int(z) for z in zsu], dtype=np.int32), np.array(bops, dtype=np.int32).reshape(-1, 2),
            np.array(bots, dtype=np.int32).reshape(-1, 3))


def get_neighbors(rbcut, coords):
    """xb.py:36-51: (nbmax, nbrs (na, nbmax)): atoms within rbcut of every atom (itself included), -1 padded."""
    coords = np.asarray(coords, dtype=float)
    ds = np.sqrt(((coords[:, None, :] - coords[None, :, :]) ** 2).sum(-1))
    lists = [np.nonzero(row <= rbcut)[0] for row in ds]
    nbmax = max(len(l) for l in lists)
    nbrs = -np.ones((len(coords), nbmax), dtype=np.int32)
    for i, l in enumerate(lists):
        nbrs[i, :len(l)] = l
    return nbmax, nbrs


def grid_sizes(racut, rbcut, dgrid, mbtypes):
    """(nsx, n, nu) of xb.py:110-124."""
    rs = L.f64([racut, rbcut])
    nsx, n, nu = np.zeros(3, dtype=np.int32), C.c_int(0), C.c_int(0)
    L.call("aqml_pairwise_dims", L.ptr(rs), float(dgrid), len(mbtypes[0]), len(mbtypes[1]), len(mbtypes[2]), L.ptr(nsx),
           C.byref(n), C.byref(nu))
    return [int(v) for v in nsx], n.value, nu.value


def calc_all_local_spectrum(nas, zs, coords, mbtypes, nbrs=None, racut=3.2, rbcut=4.2, dgrid=0.04, sigma=0.05, w2=1.0, rpower2=6,
                            w3=1. / 3, rpower3=3, as_numpy=False):
    """Atom rows ys (A, n) and -- with a neighbour table `nbrs` (A, nbmax; molecule-local indices, -1 padded) -- bond rows
    ys2 (A, nbmax, nu) of all molecules (fslatm_pairwise.f90:644-673 with ia0 = ib0 = 0).  CUDA tensors unless as_numpy."""
    import torch
    nas, zs, coords = L.i32(nas), L.i32(zs), L.f64(coords)
    if coords.shape != (int(nas.sum()), 3) or len(zs) != coords.shape[0]:
        raise ValueError("coords must be (sum(nas), 3) and zs must have sum(nas) entries")
    mbs1, mbs2, mbs3 = (L.i32(m) for m in mbtypes)
    nsx, n, nu = grid_sizes(racut, rbcut, dgrid, (mbs1, mbs2, mbs3))
    A = coords.shape[0]
    nbmax = 0
    if nbrs is not None:
        nbrs = L.i32(nbrs)
        if nbrs.ndim != 2 or nbrs.shape[0] != A:
            raise ValueError("nbrs must be (sum(nas), nbmax)")
        nbmax = nbrs.shape[1]
    dev = torch.device("cuda", torch.cuda.current_device())
    ys = torch.empty((A, n), dtype=torch.float64, device=dev)
    ys2 = torch.empty((A, nbmax, nu), dtype=torch.float64, device=dev) if nbmax else None
    rs = L.f64([racut, rbcut])
    L.call("aqml_pairwise_local_spectrum_dev", L.ptr(coords), L.ptr(zs), L.ptr(nas), len(nas), L.ptr(nbrs) if nbmax else None, nbmax,
           L.ptr(mbs1), len(mbs1), L.ptr(mbs2), len(mbs2), L.ptr(mbs3), len(mbs3), L.ptr(rs), float(dgrid), float(sigma),
           float(w2), float(rpower2), float(w3), float(rpower3), L.ptr(ys), L.ptr(ys2) if nbmax else None, L.current_stream())
    if as_numpy:
        return ys.cpu().numpy(), (ys2.cpu().numpy() if nbmax else None)
    return ys, ys2


def calc_local_spectrum(zs, coords, nbrs, mbtypes, **kw):
    """One molecule (fslatm_pairwise.f90:525-641, xb.py:175 `fget_local_spectrum`)."""
    return calc_all_local_spectrum([len(zs)], zs, coords, mbtypes, nbrs=nbrs, **kw)


def _molecular_kernels(x1, nas1, x2, nas2, kwds):
    from .. import kernels
    kwds = np.asarray(kwds, dtype=float).reshape(-1, 1)
    one1, one2 = np.ones(int(np.sum(nas1)), dtype=np.int32), np.ones(int(np.sum(nas2)), dtype=np.int32)
    return kernels.get_local_kernels_gaussian(x1, x2, one1, one2, nas1, nas2, False, 1, kwds, center=True)


def calc_mk1(nas, zs, coords, mbtypes, kwds, **kw):
    """fslatm_pairwise.f90:740-829 (local): k1 (nco, nm, nm) CUDA tensor."""
    ys, _ = calc_all_local_spectrum(nas, zs, coords, mbtypes, **kw)
    return _molecular_kernels(ys, nas, ys, nas, kwds)


def calc_mk2(nm1, nas, zs, coords, mbtypes, kwds, **kw):
    """fslatm_pairwise.f90:832-917 (local): k2 (nco, nm - nm1, nm1): test molecules nm1.. against training molecules 0..nm1-1."""
    nas = np.asarray(nas, dtype=int)
    ys, _ = calc_all_local_spectrum(nas, zs, coords, mbtypes, **kw)
    a1 = int(nas[:nm1].sum())
    return _molecular_kernels(ys[a1:], nas[nm1:], ys[:a1], nas[:nm1], kwds)


def calc_dij_max(nas, zs, coords, mbtypes, **kw):
    """fslatm_pairwise.f90:920-1005 (local): the largest distance between two atoms' rows."""
    from .. import distance
    ys, _ = calc_all_local_spectrum(nas, zs, coords, mbtypes, **kw)
    return distance.max_distance(ys, ys, 2)
