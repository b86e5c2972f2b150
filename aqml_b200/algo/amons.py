"""Amons workflow with many queries (algo/aqml.py:505-559, 619-620, 744-789).

With a map file the reference loops over the queries: query q is trained on ITS amon subset ``maps[q]``
(-1 padded), with many-body types, per-element dmax and hence sigmas taken from that subset + the query.
Here every stage of one iteration runs on the GPU (aSLATM -> per-element max distance -> local Gaussian
kernels); the iteration itself stays a host loop, as upstream.  Using the many-body types of the union of
all molecules instead of the subset's only adds all-zero columns, which change no distance (SURVEY 3.5 ii).
"""
from __future__ import annotations

import numpy as np

from .. import fused
from ..representation.core import get_slatm_mbtypes


def _subset(nas, zs, coords, idx):
    off = np.concatenate(([0], np.cumsum(nas)))
    sel = np.concatenate([np.arange(off[i], off[i + 1]) for i in idx]) if len(idx) else np.zeros(0, dtype=np.int64)
    return np.asarray(nas)[idx].astype(np.int32), np.asarray(zs)[sel].astype(np.int32), np.asarray(coords)[sel]


def query_kernels(amons, queries, maps, coeffs=(1.0,), mbtypes=None, rcut=4.8, dgrids=(0.04, 0.04),
                  widths=(0.05, 0.05), exclude=()):
    """For every query q: (idx, sigmas, ks1, ks2) with ``idx`` the amon indices used (sorted, aqml.py:556-558),
    ``sigmas`` (ncoeffs, zmax), ``ks1`` (ncoeffs, n_q, n_q) and ``ks2`` (ncoeffs, 1, n_q) as NumPy arrays.

    amons / queries: (nas, zs, coords) of the whole amon pool / of the query molecules;
    maps: int array (n_queries, namon_max), -1 padded, entries = amon indices (cheminfo/algo/amon.py:481-493)."""
    nas_a, zs_a, co_a = amons
    nas_q, zs_q, co_q = queries
    if mbtypes is None:
        off_a = np.concatenate(([0], np.cumsum(nas_a)))
        off_q = np.concatenate(([0], np.cumsum(nas_q)))
        zl = [zs_a[off_a[i]:off_a[i + 1]] for i in range(len(nas_a))] + \
             [zs_q[off_q[i]:off_q[i + 1]] for i in range(len(nas_q))]
        mbtypes = get_slatm_mbtypes(zl)
    kw = dict(sigmas=tuple(widths), dgrids=tuple(dgrids), rcut=rcut)
    zmax = int(max(np.max(zs_a), np.max(zs_q)))
    factor = 1.0 / np.sqrt(2.0 * np.log(2.0))  # aqml.py:195
    out = []
    for q in range(len(nas_q)):
        row = np.asarray(maps[q])
        idx = sorted(set(int(i) for i in row[row > -1]).difference(set(exclude)))
        a = _subset(nas_a, zs_a, co_a, idx)
        t = _subset(nas_q, zs_q, co_q, [q])
        both = (np.concatenate((a[0], t[0])), np.concatenate((a[1], t[1])), np.concatenate((a[2], t[2])))
        r_all = fused.PackedRep(both[2], both[1], both[0], mbtypes, **kw)
        dsmax = fused.kernel_width(r_all, r_all, zmax)                        # dmxby='a' (aqml.py:748-749)
        present = np.unique(both[1])
        lone = [int(z) for z in present if dsmax[z - 1] == 0.0]
        if lone:
            raise ValueError(f"query {q}: element(s) {lone} occur once in amon subset + query, so dmax = 0 and sigma = 0 "
                             "(the reference produces NaN kernels here, aqml.py:249-251,775)")
        sig = np.array([factor * dsmax * c for c in coeffs])                  # aqml.py:775
        r1 = fused.PackedRep(a[2], a[1], a[0], mbtypes, **kw)
        r2 = fused.PackedRep(t[2], t[1], t[0], mbtypes, **kw)
        ks1 = fused.local_gaussian(r1, r1, sig, zmax, symmetric=True).cpu().numpy()
        ks2 = fused.local_gaussian(r2, r1, sig, zmax).cpu().numpy()
        for r in (r_all, r1, r2):
            r.close()
        out.append((idx, sig, ks1, ks2))
    return out
