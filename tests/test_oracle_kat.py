"""Pin the oracle: reference known-answer test + NumPy-vs-C restatement agreement (CPU only)."""
import numpy as np

from conftest import mbtypes_from_array, rel_err
from oracle import c_oracle as co
from oracle import np_oracle as o
from aqml_b200 import synth


def test_readme_learning_curve(demo):
    """demo/example/README.md:57-63, four printed decimals, kcal/mol."""
    curve, readme = demo["curve"], demo["readme_curve"]
    assert np.array_equal(curve[:, :2], readme[:, :2])
    assert np.all(np.abs(curve[:, 2] - readme[:, 2]) < 5.1e-5)
    assert np.all(np.abs(curve[:, 3] - readme[:, 3]) < 5.1e-5)
    # 'DressedAtom' column: rows 2..7 match as printed; row 1 (rank-deficient free-atom
    # fallback) is printed with the opposite sign to what aqml.py:894 computes
    assert np.all(np.abs(curve[1:, 4] - readme[1:, 4]) < 5.1e-5)
    assert abs(abs(curve[0, 4]) - abs(readme[0, 4])) < 5.1e-5


def test_np_oracle_recomputes_golden(demo):
    n1 = int(demo["n1"])
    nas, zs, coords = demo["nas"], demo["zs"], demo["coords"]
    mbtypes = mbtypes_from_array(demo["mbtypes"])
    off = np.concatenate(([0], np.cumsum(nas)))
    zs_list = [zs[off[i]:off[i + 1]] for i in range(len(nas))]
    assert o.get_slatm_mbtypes(zs_list) == mbtypes
    m = 3  # one amon is enough for the pure-Python loops
    x = np.array(o.generate_slatm(coords[off[m]:off[m + 1]], zs_list[m], mbtypes, local=True,
                                  sigmas=[.05, .05], dgrids=[.04, .04], rcut=4.8))
    assert rel_err(x, demo["x_local"][off[m]:off[m + 1]]) < 1e-13
    dso = o.get_kernel_width(nas, zs, demo["x_local"])
    assert np.allclose(dso, demo["dso"], rtol=1e-14, atol=0)
    A1 = off[n1]
    k2 = o.calc_lk_gaussian(demo["x_local"][A1:], demo["x_local"][:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1],
                            8, False, demo["sigmas"])
    assert rel_err(k2, demo["k2"]) < 1e-13


def test_c_oracle_slatm_matches_golden(demo):
    nas, zs, coords = demo["nas"], demo["zs"], demo["coords"]
    mbtypes = mbtypes_from_array(demo["mbtypes"])
    kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
    xl = co.generate_slatm_batch(coords, zs, nas, mbtypes, True, **kw)
    xg = co.generate_slatm_batch(coords, zs, nas, mbtypes, False, **kw)
    assert xl.shape == demo["x_local"].shape
    assert rel_err(xl, demo["x_local"]) < 1e-12
    assert rel_err(xg, demo["x_global"]) < 1e-12
    # identity: global == sum over atoms of local (intended 3-body index semantics), except the
    # 1-body slots which hold Z*count either way
    off = np.concatenate(([0], np.cumsum(nas)))
    summed = np.array([xl[off[i]:off[i + 1]].sum(axis=0) for i in range(len(nas))])
    assert rel_err(summed, xg) < 1e-12


def test_c_oracle_kernels_match_golden(demo):
    n1 = int(demo["n1"])
    nas, zs, x = demo["nas"], demo["zs"], demo["x_local"]
    A1 = int(nas[:n1].sum())
    k1 = co.calc_lk_gaussian(x[:A1], x[:A1], zs[:A1], zs[:A1], nas[:n1], nas[:n1], 8, False, demo["sigmas"])
    k2 = co.calc_lk_gaussian(x[A1:], x[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], 8, False, demo["sigmas"])
    assert rel_err(k1, demo["k1"]) < 1e-12
    assert rel_err(k2, demo["k2"]) < 1e-12
    assert np.allclose(co.get_kernel_width(nas, zs, x), demo["dso"], rtol=1e-13)


def test_c_vs_np_on_synthetic():
    nas, zs, coords = synth.qm9_like(6, 7)
    zs_list = synth.split_zs(nas, zs)
    mbtypes = o.get_slatm_mbtypes(zs_list)
    kw = dict(sigmas=(.05, .05), dgrids=(.1, .1), rcut=4.8)
    xl_c = co.generate_slatm_batch(coords, zs, nas, mbtypes, True, **kw)
    xg_c = co.generate_slatm_batch(coords, zs, nas, mbtypes, False, **kw)
    off = np.concatenate(([0], np.cumsum(nas)))
    for m in range(2):
        sl = slice(off[m], off[m + 1])
        xl = np.array(o.generate_slatm(coords[sl], zs[sl], mbtypes, local=True, **kw))
        xg = o.generate_slatm(coords[sl], zs[sl], mbtypes, local=False, **kw)
        assert rel_err(xl_c[sl], xl) < 1e-12
        assert rel_err(xg_c[m], xg) < 1e-12
    # izeff + rpower variants on one molecule
    kw2 = dict(sigmas=(.07, .04), dgrids=(.2, .15), rcut=3.5, rpower=3, izeff=True)
    sl = slice(off[0], off[1])
    xl = np.array(o.generate_slatm(coords[sl], zs[sl], mbtypes, local=True, **kw2))
    xc = co.generate_slatm_batch(coords[sl], zs[sl], nas[:1], mbtypes, True, **kw2)
    assert rel_err(xc, xl) < 1e-12
    # kernels
    rng = np.random.default_rng(0)
    A = rng.random((9, 37))
    B = rng.random((5, 37))
    assert rel_err(co.gaussian_kernel(A, B, 1.3), o.gaussian_kernel_direct(A, B, 1.3)) < 1e-13
    assert rel_err(co.laplacian_kernel(A, B, 2.1), o.laplacian_kernel(A, B, 2.1)) < 1e-13
    assert rel_err(co.l2_distance(A, B), o.fl2_distance(A, B)) < 1e-13
    zmax = int(zs.max())
    sig = np.abs(rng.random((3, zmax))) + 0.5
    n1 = 4
    A1 = off[n1]
    kc = co.calc_lk_gaussian(xl_c[A1:], xl_c[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], zmax, False, sig)
    kn = o.calc_lk_gaussian(xl_c[A1:], xl_c[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], zmax, False, sig)
    assert rel_err(kc, kn) < 1e-12
    kc = co.calc_lk_gaussian(xl_c[A1:], xl_c[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], zmax, True, sig * 30)
    kn = o.calc_lk_gaussian(xl_c[A1:], xl_c[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], zmax, True, sig * 30)
    assert rel_err(kc, kn) < 1e-12
    kc = co.calc_lk_laplacian(xl_c[A1:], xl_c[:A1], nas[n1:], nas[:n1], [50., 200.])
    kn = o.calc_lk_laplacian(xl_c[A1:], xl_c[:A1], nas[n1:], nas[:n1], [50., 200.])
    assert rel_err(kc, kn) < 1e-12


# demo/example/README.md:75-94  (`aqml -train g7 -test target -p b3lypvdz -ieaq -prog orca`): atom, Z, E_A in kcal/mol
README_ATOMIC = [(6, -163.80), (6, -163.55), (6, -163.57), (6, -163.55), (6, -163.77), (6, -163.42), (6, -159.70),
                 (6, -161.60), (8, -90.76), (1, -59.83), (1, -59.98), (1, -59.99), (1, -59.98), (1, -59.78),
                 (1, -64.12), (1, -63.75), (1, -60.70)]


def test_readme_atomic_energies(demo):
    """The reference's second known-answer test: atomic contributions to the atomisation energy of the target molecule
    after training on all 16 amons (README two decimals).  Pins calc_ka (aqml.py:258-309), the dressed-atom baseline
    and the free-atom table on top of the aSLATM / width / KRR chain pinned by the learning curve."""
    n1 = int(demo["n1"])
    nas, zs, x = demo["nas"], demo["zs"], demo["x_local"]
    off = np.concatenate(([0], np.cumsum(nas)))
    A1 = off[n1]
    zs_list = [zs[off[i]:off[i + 1]] for i in range(len(nas))]
    zsu = np.unique(zs)
    ngs = np.array([[(z == zz).sum() for zz in zsu] for z in zs_list])
    ys = demo["energies"] * o.H2KC
    sig = demo["sigmas"][1:2]                                   # coefficient 1.0, as in the README run
    idx1, idx2 = np.arange(n1), np.arange(n1, n1 + 1)
    ys1, ys2, esb = o.calc_ae_dressed(ngs, ys, idx1, idx2, zsu)
    k1 = demo["k1"][1].copy()
    k1[np.diag_indices_from(k1)] += 1e-4
    alphas = np.linalg.solve(k1, ys1)
    ksa21 = o.calc_ka(x[A1:], x[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], 8, sig)[0]
    ea, easq = o.atomic_energies(ksa21, alphas, zs[A1:], zsu, esb)
    # aqml.py:909-910: the atomic contributions add up to the molecular prediction
    assert abs(easq.sum() - np.dot(demo["k2"][1], alphas)[0]) < 1e-9
    assert [int(z) for z in zs[A1:]] == [z for z, _ in README_ATOMIC]
    assert np.all(np.abs(ea - np.array([e for _, e in README_ATOMIC])) < 5.1e-3), ea


def test_docs_unit_conversion(demo):
    """docs/source/users_guide/amons_qml.rst:81,97-105: the b3lypvdz energies of the first five g7 amons in kcal/mol
    (Hartree x io2.Units().h2kc), as the reference's own reader prints them."""
    doc = np.array([-25401.85283769, -49280.26416003, -71817.46145015, -73935.93419366, -96479.43688458])
    assert np.allclose(demo["energies"][:5] * o.H2KC, doc, rtol=0, atol=5e-9)
