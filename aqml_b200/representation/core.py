"""Host mirror of ``cml.representation.core`` SLATM routines (coreml/cml/representation/core.py:324-556)."""
from __future__ import annotations

import numpy as np

from .. import _lib as L
from .slatm import mbtypes_from_counts


def get_slatm_mbtypes(nuclear_charges, pbc=False):
    """core.py:324-379: many-body types of a data set (list of per-molecule charge arrays)."""
    zs = nuclear_charges
    zsmax = sorted({int(z) for zsi in zs for z in zsi})
    nass = np.array([[int((np.asarray(zsi) == zi).sum()) for zi in zsmax] for zsi in zs])
    nzmax = np.max(nass, axis=0)
    if pbc:
        nzmax = np.where(nzmax <= 2, 3, nzmax)  # core.py:356-361
    return mbtypes_from_counts(zsmax, nzmax)


def generate_slatm_batch(coordinates, nuclear_charges, nas, mbtypes, izeff=False, local=False,
                         sigmas=[0.05, 0.05], dgrids=[0.03, 0.03], rcut=4.8, rpower=6):
    """Whole data set in one call.  coordinates (A,3), nuclear_charges (A,), nas (nmol,).

    Returns (A, D) aSLATM rows if ``local`` else (nmol, D) SLATM rows; a CUDA tensor when
    ``coordinates`` is a CUDA tensor, else a NumPy array."""
    p = L.slatm_params(mbtypes, izeff, sigmas, dgrids, rcut, rpower)
    nas = L.i32(nas)
    D = L.load().aqml_slatm_dim(p)
    if L.is_tensor(coordinates):
        import torch
        coords = L.dev_f64(coordinates).contiguous()
        zs = nuclear_charges
        if not L.is_tensor(zs):
            zs = torch.as_tensor(L.i32(zs), device=coords.device)
        zs = zs.to(torch.int32).contiguous()
        # the C ABI only sees raw pointers: a mismatch would read out of bounds on the device
        if tuple(coords.shape) != (int(zs.shape[0]), 3) or int(nas.sum()) != int(zs.shape[0]):
            raise ValueError(f"{coords.shape[0]} coordinates, but {int(zs.shape[0])} atom_types "
                             f"(nas sums to {int(nas.sum())})!")
        rows = coords.shape[0] if local else len(nas)
        out = torch.empty((rows, D), dtype=torch.float64, device=coords.device)
        with torch.cuda.device(coords.device):
            L.call("aqml_generate_slatm_batch_dev", L.ptr(coords), L.ptr(zs), L.ptr(nas), len(nas), p, int(bool(local)),
                   L.ptr(out), D, L.current_stream())
        return out
    coords = L.f64(coordinates)
    zs = L.i32(nuclear_charges)
    if coords.shape != (len(zs), 3) or int(nas.sum()) != len(zs):
        raise ValueError(f"{coords.shape[0]} coordinates, but {len(zs)} atom_types!")
    rows = len(zs) if local else len(nas)
    out = np.empty((rows, D))
    L.call("aqml_generate_slatm_batch", L.ptr(coords), L.ptr(zs), L.ptr(nas), len(nas), p, int(bool(local)), L.ptr(out))
    return out


def generate_slatm(coordinates, nuclear_charges, mbtypes, cg=None, izeff=False,
                   cell=None, local=False, sigmas=[0.05, 0.05], dgrids=[0.03, 0.03],
                   rcut=4.8, alchemy=False, pbc=False, rpower=6, ias=None):
    """core.py:382-556.  local -> list of per-atom 1-D arrays, else one 1-D array."""
    if alchemy:
        raise NotImplementedError("alchemy=True (core.py:445-453) is not used by any workflow; out of scope")
    zs = np.asarray(nuclear_charges)
    if pbc:
        # slatm.py:40-45,96-101: one finite cluster of periodic images per query atom, query atom first
        assert local, '#ERROR: for periodic system, plz use atomic rpst'
        from .slatm import MolPBC
        mol = MolPBC(zs, np.asarray(coordinates, dtype=np.float64).reshape(-1, 3), cell, rcut=rcut)
        idx = range(len(zs)) if ias is None else ias
        clusters = [mol.get_cluster(int(i)) for i in idx]
        nas = [len(c[0]) for c in clusters]
        x = generate_slatm_batch(np.concatenate([c[1] for c in clusters]), np.concatenate([c[0] for c in clusters]),
                                 nas, mbtypes, izeff=izeff, local=True, sigmas=sigmas, dgrids=dgrids, rcut=rcut,
                                 rpower=rpower)
        first = np.concatenate(([0], np.cumsum(nas)[:-1]))
        return [x[int(i)] for i in first]
    x = generate_slatm_batch(np.asarray(coordinates, dtype=np.float64).reshape(-1, 3), zs, [len(zs)], mbtypes,
                             izeff=izeff, local=local, sigmas=sigmas, dgrids=dgrids, rcut=rcut, rpower=rpower)
    if local:
        idx = range(len(zs)) if ias is None else ias
        return [x[int(i)] for i in idx]
    return x[0]
