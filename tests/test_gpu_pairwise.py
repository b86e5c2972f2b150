"""Pair-wise / bond aSLATM (fslatm_pairwise.f90) on the GPU against oracle/np_pairwise.py.  PARITY UNPINNED: the
reference ships no test or known answer for this file (it is not even built); the oracle defines the semantics."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _mols(seed, nmol=5):
    from aqml_b200 import synth
    nas, zs, coords = synth.qm9_like(nmol, seed)
    return np.asarray(nas), np.asarray(zs), np.asarray(coords)


def _relmax(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("kw", [dict(), dict(racut=2.6, rbcut=3.1, dgrid=0.05, sigma=0.07, w2=0.7, rpower2=4, w3=0.4, rpower3=2.5)])
def test_atom_and_bond_rows(kw):
    from aqml_b200.representation import pairwise as pw
    from oracle import np_pairwise as o
    nas, zs, coords = _mols(7)
    mb = pw.get_slatm_mbtypes(zs, nas)
    omb = o.get_slatm_mbtypes(zs, nas)
    assert all(np.array_equal(a, b) for a, b in zip(mb, omb))
    p = dict(racut=3.2, rbcut=4.2, dgrid=0.04, sigma=0.05, w2=1.0, rpower2=6, w3=1. / 3, rpower3=3)
    p.update(kw)
    off = np.concatenate(([0], np.cumsum(nas)))
    nbl = [pw.get_neighbors(p["rbcut"], coords[off[i]:off[i + 1]]) for i in range(len(nas))]
    for (nb, t), i in zip(nbl, range(len(nas))):
        onb, ot = o.get_neighbors(p["rbcut"], coords[off[i]:off[i + 1]])
        assert nb == onb and np.array_equal(t, ot)
    nbmax = max(nb for nb, _ in nbl)
    nbrs = -np.ones((int(nas.sum()), nbmax), dtype=np.int32)
    for i, (nb, t) in enumerate(nbl):
        nbrs[off[i]:off[i + 1], :nb] = t
    ys, ys2 = pw.calc_all_local_spectrum(nas, zs, coords, mb, nbrs=nbrs, as_numpy=True, **p)
    nsx = o.grid_sizes(p["racut"], p["rbcut"], p["dgrid"])
    assert pw.grid_sizes(p["racut"], p["rbcut"], p["dgrid"], mb)[0] == nsx
    rs = [p["racut"], p["rbcut"]]
    for i in range(len(nas)):
        sl = slice(off[i], off[i + 1])
        rys, rys2 = o.calc_local_spectrum(zs[sl], coords[sl], nbrs[sl], omb[0], omb[1], omb[2], nsx, rs, p["dgrid"], p["sigma"],
                                          p["w2"], p["rpower2"], p["w3"], p["rpower3"])
        assert rys.shape == ys[sl].shape and rys2.shape == ys2[sl].shape
        assert _relmax(ys[sl], rys) < TOL
        assert _relmax(ys2[sl], rys2) < TOL
        assert np.array_equal(ys2[sl] != 0, rys2 != 0) or _relmax(ys2[sl], rys2) < 1e-13   # the same slots are filled
    assert np.abs(ys2).max() > 0 and np.abs(ys).max() > 0


def test_single_molecule_without_bonds_and_edge_cases():
    from aqml_b200.representation import pairwise as pw
    from oracle import np_pairwise as o
    # one atom; two coincident atoms (distance 0 <= eps: no three-body term, two-body term at r = 0); a linear molecule
    cases = [(np.array([6]), np.zeros((1, 3))),
             (np.array([1, 1, 8]), np.array([[0., 0, 0], [0, 0, 0], [0.9, 0.1, 0]])),
             (np.array([8, 6, 8]), np.array([[-1.16, 0, 0], [0., 0, 0], [1.16, 0, 0]]))]
    for zs, coords in cases:
        nas = np.array([len(zs)])
        mb = pw.get_slatm_mbtypes(zs, nas)
        ys, ys2 = pw.calc_local_spectrum(zs, coords, None, mb, as_numpy=True)
        assert ys2 is None
        nsx = o.grid_sizes(3.2, 4.2, 0.04)
        rys, _ = o.calc_local_spectrum(zs, coords, -np.ones((len(zs), 1), dtype=int), mb[0], mb[1], mb[2], nsx, [3.2, 4.2], 0.04, 0.05,
                                       1.0, 6, 1. / 3, 3, bonds=False)
        assert np.array_equal(np.isnan(ys), np.isnan(rys))
        m = ~np.isnan(rys)
        assert np.max(np.abs(ys[m] - rys[m])) <= TOL * max(np.max(np.abs(rys[m])), 1e-300)


def test_molecular_kernels_and_dmax():
    from aqml_b200.representation import pairwise as pw
    from oracle import np_pairwise as o
    nas, zs, coords = _mols(11, nmol=6)
    mb = pw.get_slatm_mbtypes(zs, nas)
    p = dict(racut=3.2, rbcut=4.2, dgrid=0.04, sigma=0.05, w2=1.0, rpower2=6, w3=1. / 3, rpower3=3)
    rs = [p["racut"], p["rbcut"]]
    args = (mb, rs, p["dgrid"], p["sigma"], p["w2"], p["rpower2"], p["w3"], p["rpower3"])
    dmax = pw.calc_dij_max(nas, zs, coords, mb, **p)
    rd = o.calc_dij_max(nas, zs, coords, *args)
    assert abs(dmax - rd) <= 1e-12 * rd
    kwds = [0.2 * rd, 0.5 * rd, rd]
    k1 = pw.calc_mk1(nas, zs, coords, mb, kwds, **p).cpu().numpy()
    r1 = o.calc_mk1(nas, zs, coords, *args, kwds)
    assert k1.shape == r1.shape and np.max(np.abs(k1 - r1) / np.abs(r1)) < TOL
    nm1 = 4
    k2 = pw.calc_mk2(nm1, nas, zs, coords, mb, kwds, **p).cpu().numpy()
    r2 = o.calc_mk2(nm1, nas, zs, coords, *args, kwds)
    assert k2.shape == r2.shape and np.max(np.abs(k2 - r2) / np.abs(r2)) < TOL
