"""GPU parity: SLATM / aSLATM through the C ABI vs the CPU oracle and the golden demo fixture."""
import numpy as np
import pytest

from conftest import mbtypes_from_array, rel_err
from aqml_b200 import synth
from aqml_b200.representation import core as rcore
from aqml_b200.representation import slatm as rslatm
from oracle import c_oracle as co
from oracle import np_oracle as o

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_golden_demo(demo):
    nas, zs, coords = demo["nas"], demo["zs"], demo["coords"]
    mb = mbtypes_from_array(demo["mbtypes"])
    kw = dict(sigmas=[.05, .05], dgrids=[.04, .04], rcut=4.8)
    xl = rcore.generate_slatm_batch(coords, zs, nas, mb, local=True, **kw)
    assert xl.shape == demo["x_local"].shape
    assert rel_err(xl, demo["x_local"]) < TOL
    xg = rcore.generate_slatm_batch(coords, zs, nas, mb, local=False, **kw)
    assert rel_err(xg, demo["x_global"]) < TOL
    # reference call surface: per molecule, list of per-atom arrays / one array
    off = np.concatenate(([0], np.cumsum(nas)))
    m = 16
    xi = rcore.generate_slatm(coords[off[m]:off[m + 1]], zs[off[m]:off[m + 1]], mb, local=True, **kw)
    assert isinstance(xi, list) and len(xi) == nas[m] and xi[0].shape == (1959,)
    assert rel_err(np.array(xi), demo["x_local"][off[m]:off[m + 1]]) < TOL
    xs = rcore.generate_slatm(coords[off[m]:off[m + 1]], zs[off[m]:off[m + 1]], mb, local=True, ias=[2, 5], **kw)
    assert np.array_equal(np.array(xs), np.array(xi)[[2, 5]])
    xgi = rcore.generate_slatm(coords[off[m]:off[m + 1]], zs[off[m]:off[m + 1]], mb, **kw)
    assert xgi.shape == (1959,) and rel_err(xgi, demo["x_global"][m]) < TOL


@pytest.mark.parametrize("cfg", [
    dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8),
    dict(sigmas=(.05, .05), dgrids=(.03, .03), rcut=4.8),           # generate_slatm defaults
    dict(sigmas=(.07, .04), dgrids=(.2, .15), rcut=3.5, rpower=3, izeff=True),
    dict(sigmas=(.05, .05), dgrids=(.1, .1), rcut=2.0),             # short cut-off: many empty blocks
])
def test_synthetic_vs_oracle(cfg):
    nas, zs, coords = synth.qm9_like(40, 77)
    mb = o.get_slatm_mbtypes(synth.split_zs(nas, zs))
    for local in (True, False):
        ref = co.generate_slatm_batch(coords, zs, nas, mb, local, **cfg)
        got = rcore.generate_slatm_batch(coords, zs, nas, mb, local=local, **cfg)
        assert got.shape == ref.shape
        assert rel_err(got, ref) < TOL, (local, cfg)


def test_global_is_sum_of_local():
    nas, zs, coords = synth.qm9_like(25, 5)
    mb = o.get_slatm_mbtypes(synth.split_zs(nas, zs))
    kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
    xl = rcore.generate_slatm_batch(coords, zs, nas, mb, local=True, **kw)
    xg = rcore.generate_slatm_batch(coords, zs, nas, mb, local=False, **kw)
    off = np.concatenate(([0], np.cumsum(nas)))
    summed = np.array([xl[off[i]:off[i + 1]].sum(axis=0) for i in range(len(nas))])
    assert rel_err(summed, xg) < 1e-12


def test_large_molecules_eight_elements():
    nas, zs, coords = synth.large_like(3, 9, lo=60, hi=90)
    zl = synth.split_zs(nas, zs)
    mb = o.get_slatm_mbtypes(zl)
    kw = dict(sigmas=(.05, .05), dgrids=(.08, .08), rcut=4.8)
    for local in (True, False):
        ref = co.generate_slatm_batch(coords, zs, nas, mb, local, **kw)
        got = rcore.generate_slatm_batch(coords, zs, nas, mb, local=local, **kw)
        assert rel_err(got, ref) < TOL


def nan_rel_err(got, ref):
    """NaN pattern must be identical (the reference propagates 0/0 for coincident atoms); error on the rest."""
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    return rel_err(got[ok], ref[ok])


def test_edge_cases():
    kw = dict(sigmas=(.05, .05), dgrids=(.1, .1), rcut=4.8)
    # single atom; coincident atoms (excluded from 3-body by eps, counted by 2-body); pair exactly at rcut
    zs = np.array([6, 1, 1, 8, 1], dtype=np.int32)
    coords = np.array([[0, 0, 0], [1.0, 0, 0], [1.0, 0, 0], [0, 4.8, 0], [0, 0, 1.1]], dtype=np.float64)
    nas = np.array([1, 4], dtype=np.int32)
    mb = [[1], [6], [8], [1, 1], [1, 6], [1, 8], [6, 8], [1, 6, 1], [1, 8, 1], [1, 1, 8], [1, 1, 1], [1, 1, 6], [8, 1, 6]]
    for local in (True, False):
        ref = co.generate_slatm_batch(coords, zs, nas, mb, local, **kw)
        got = rcore.generate_slatm_batch(coords, zs, nas, mb, local=local, **kw)
        # two coincident atoms that are both neighbours of a third give 0/0 in calc_cos_angle
        # (utils.f90:96-98; no eps test on d(i,k), fslatm.f90:352): NaN in the reference, NaN here
        assert np.isnan(ref).any() and np.array_equal(np.isnan(got), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert rel_err(got[ok], ref[ok]) < TOL
    # same molecule without the duplicate atom: finite everywhere
    keep = np.array([0, 1, 3, 4])
    zs_f, coords_f, nas_f = zs[keep], coords[keep], np.array([1, 3], dtype=np.int32)
    for local in (True, False):
        ref = co.generate_slatm_batch(coords_f, zs_f, nas_f, mb, local, **kw)
        got = rcore.generate_slatm_batch(coords_f, zs_f, nas_f, mb, local=local, **kw)
        assert np.isfinite(ref).all() and rel_err(got, ref) < TOL
    # element present in the molecule but in no many-body type contributes nothing
    mb2 = [[1], [1, 1], [1, 1, 1]]
    ref = co.generate_slatm_batch(coords, zs, nas, mb2, True, **kw)
    got = rcore.generate_slatm_batch(coords, zs, nas, mb2, local=True, **kw)
    assert nan_rel_err(got, ref) < TOL
    # type listed with z1 > z3 (orientation flip) and arbitrary order of the list
    mb3 = [[6, 1], [8, 1, 6], [1], [1, 6, 1]]
    ref = co.generate_slatm_batch(coords, zs, nas, mb3, True, **kw)
    got = rcore.generate_slatm_batch(coords, zs, nas, mb3, local=True, **kw)
    assert nan_rel_err(got, ref) < TOL


def test_f2py_level_single_type_calls():
    nas, zs, coords = synth.qm9_like(2, 13)
    zs, coords = zs[:nas[0]], coords[:nas[0]]
    obj = [zs, coords, None]
    for mbt in ([1, 6], [6, 6], [1, 1]):
        for ia in (None, 0, 3):
            ref = o.get_sbop(mbt, obj, iloc=ia is not None, ia=ia, dgrid=0.05)
            got = rslatm.get_sbop(mbt, obj, iloc=ia is not None, ia=ia, dgrid=0.05)
            assert rel_err(got, ref) < TOL
    for mbt in ([1, 6, 1], [1, 6, 6], [6, 1, 6]):
        for ia in (None, 0, 3):
            ref = o.get_sbot(mbt, obj, iloc=ia is not None, ia=ia, dgrid=0.05)
            got = rslatm.get_sbot(mbt, obj, iloc=ia is not None, ia=ia, dgrid=0.05)
            assert rel_err(got, ref) < TOL
    assert rslatm.get_boa(6, zs)[0] == 6 * (zs == 6).sum()
    from cml.representation import fslatm
    ys = fslatm.calc_sbop_local(coords, zs, 0, None, False, 1, 6, 4.8, 95, 0.05, 0.05, 7.9788, 6)
    ref = o.calc_sbop(coords, zs, 1, 6, 4.8, 95, 0.05, 0.05, 7.9788, 6, ia=0)
    assert rel_err(ys, ref) < TOL
    with pytest.raises(ValueError):
        rslatm.get_sbop([1, 6], [zs[:-1], coords, None])


def test_random_tiny_molecules_all_branches():
    """Randomised small cases: 1-7 atoms, 1-3 elements, atoms in and out of the cut-off, every type
    orientation (z1==z2, z1!=z2, z1==z3, z1!=z3, centre not of type), local and global, both kernels paths."""
    rng = np.random.default_rng(2024)
    for trial in range(12):
        nmol = int(rng.integers(1, 6))
        nas = rng.integers(1, 8, size=nmol).astype(np.int32)
        pool = rng.choice([1, 6, 7, 8, 16], size=int(rng.integers(1, 4)), replace=False)
        zs = rng.choice(pool, size=int(nas.sum())).astype(np.int32)
        coords = rng.uniform(-2.2, 2.2, size=(int(nas.sum()), 3))
        # keep atoms apart (coincident atoms are covered by test_edge_cases)
        off = np.concatenate(([0], np.cumsum(nas)))
        for m in range(nmol):
            x = coords[off[m]:off[m + 1]]
            for i in range(1, len(x)):
                while np.min(np.linalg.norm(x[:i] - x[i], axis=1)) < 0.5:
                    x[i] = rng.uniform(-2.2, 2.2, size=3)
        mb = o.get_slatm_mbtypes(synth.split_zs(nas, zs))
        rng.shuffle(mb)  # arbitrary order of the type list = arbitrary column order
        cfg = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=float(rng.choice([1.5, 3.0, 4.8])))
        if trial % 3 == 2:
            cfg = dict(sigmas=(.1, .08), dgrids=(.25, .2), rcut=3.0, rpower=4, izeff=bool(set(pool) <= {1, 6, 7, 8, 16}))
        for local in (True, False):
            ref = co.generate_slatm_batch(coords, zs, nas, mb, local, **cfg)
            got = rcore.generate_slatm_batch(coords, zs, nas, mb, local=local, **cfg)
            assert rel_err(got, ref) < TOL, (trial, local, cfg)


def test_periodic_clusters():
    """pbc: one finite cluster of periodic images per query atom (slatm.py:134-210), query atom first."""
    zs = np.array([11, 17, 11, 17], dtype=np.int32)
    a = 4.2
    cell = np.array([[a, 0.0, 0.0], [0.3, a, 0.0], [0.0, 0.4, a]])
    frac = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5], [0.5, 0.5, 0.02], [0.02, 0.0, 0.5]])
    coords = frac @ cell
    mb = rcore.get_slatm_mbtypes([zs], pbc=True)
    assert mb == o.get_slatm_mbtypes([zs], pbc=True) and [11, 11, 11] in mb
    kw = dict(sigmas=[.05, .05], dgrids=[.1, .1], rcut=3.5)
    got = rcore.generate_slatm(coords, zs, mb, local=True, pbc='111', cell=cell, **kw)
    ref = o.generate_slatm_pbc(coords, zs, cell, mb, **kw)
    assert len(got) == 4
    assert rel_err(np.array(got), np.array(ref)) < TOL
    cz, cx = rslatm.MolPBC(zs, coords, cell, rcut=3.5).get_cluster(1)
    oz, ox = o.pbc_cluster(zs, coords, cell, 1, 3.5)
    assert np.array_equal(cz, oz) and np.array_equal(cx, ox) and len(cz) > 5
    y = rslatm.get_sbop([11, 17], [zs, coords, cell], iloc=True, ia=1, pbc='111', rcut=3.5, dgrid=0.1)
    yr = o.get_sbop([11, 17], [oz, ox, None], iloc=True, ia=0, rcut=3.5, dgrid=0.1)
    assert rel_err(y, yr) < TOL
