"""GPU parity: kernel matrices / distances through the C ABI vs the CPU oracle (bit-level goal:
<= 1e-10 relative per element, BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import kernel_rel_err, rel_err
from aqml_b200 import distance, kernels, synth
from aqml_b200.algo import aqml as algo
from oracle import c_oracle as co
from oracle import np_oracle as o

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def slatm_set():
    """60 QM9-shaped molecules, aSLATM + SLATM from the CPU oracle."""
    nas, zs, coords = synth.qm9_like(60, 20251019)
    mb = o.get_slatm_mbtypes(synth.split_zs(nas, zs))
    kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
    xl = co.generate_slatm_batch(coords, zs, nas, mb, True, **kw)
    xg = co.generate_slatm_batch(coords, zs, nas, mb, False, **kw)
    return dict(nas=nas, zs=zs, coords=coords, mb=mb, xl=xl, xg=xg)


@pytest.mark.parametrize("shape", [(1, 1, 1), (5, 3, 7), (37, 53, 129), (200, 131, 1000), (129, 257, 16)])
def test_global_gaussian_random(shape):
    n1, n2, D = shape
    rng = np.random.default_rng(n1 * 1000 + n2)
    A, B = rng.random((n1, D)), rng.random((n2, D))
    for sigma in (0.3 * np.sqrt(D), 2.0 * np.sqrt(D)):
        K = kernels.gaussian_kernel(A, B, sigma)
        assert K.shape == (n1, n2) and K.flags.f_contiguous
        assert kernel_rel_err(K, co.gaussian_kernel(A, B, sigma)) < TOL
        Kl = kernels.laplacian_kernel(A, B, sigma * 3)
        assert kernel_rel_err(Kl, co.laplacian_kernel(A, B, sigma * 3)) < TOL


def test_global_kernels_on_slatm(slatm_set):
    X = slatm_set["xg"]
    A, B = X[:45], X[45:]
    dmax = co.l2_distance(A, A).max()
    for c in (0.25, 1.0, 4.0):  # both ends of a sigma list
        sigma = c * dmax / np.sqrt(2 * np.log(2))
        assert kernel_rel_err(kernels.gaussian_kernel(B, A, sigma), co.gaussian_kernel(B, A, sigma)) < TOL
        K = kernels.gaussian_kernel(A, A, sigma)  # symmetric, exact-duplicate rows on the diagonal
        assert kernel_rel_err(K, co.gaussian_kernel(A, A, sigma)) < TOL
        assert np.allclose(np.diag(K), 1.0, rtol=0, atol=1e-12)
    l1max = co.l1_distance(A, A).max()
    assert kernel_rel_err(kernels.laplacian_kernel(B, A, l1max / np.log(2)), co.laplacian_kernel(B, A, l1max / np.log(2))) < TOL
    # near-duplicate rows
    A2 = A.copy()
    A2[1] = A2[0] * (1 + 1e-9)
    assert kernel_rel_err(kernels.gaussian_kernel(A2, A2, dmax), co.gaussian_kernel(A2, A2, dmax)) < TOL


def test_distances(slatm_set):
    X = slatm_set["xg"]
    A, B = X[:33], X[20:]
    D2 = distance.l2_distance(A, B)
    ref = co.l2_distance(A, B)
    assert D2.shape == ref.shape and D2.flags.f_contiguous
    # overlapping rows (d = 0 exactly) and everything else: small distances are recomputed by direct differences
    assert np.array_equal(D2 == 0, ref == 0) and (ref == 0).sum() == 13
    assert rel_err(D2, ref) < TOL
    D1 = distance.l1_distance(A, B)
    assert rel_err(D1, co.l1_distance(A, B)) < 1e-13
    with pytest.raises(ValueError):
        distance.l2_distance(A, B[:, :-1])
    with pytest.raises(ValueError):
        distance.l2_distance(A[0], B)


def test_multi_sigma_device_path(slatm_set):
    import torch
    X = slatm_set["xg"]
    A, B = X[:45], X[45:]
    sig = np.array([50., 100., 200., 400.])
    At, Bt = torch.as_tensor(A, device="cuda"), torch.as_tensor(B, device="cuda")
    for kern, ref in (("gaussian", co.gaussian_kernel), ("laplacian", co.laplacian_kernel)):
        s = sig if kern == "gaussian" else sig * 40
        K = kernels.global_kernels(At, Bt, s, kern).cpu().numpy()
        assert K.shape == (4, 45, 15)
        for i in range(4):
            assert kernel_rel_err(K[i], ref(A, B, s[i])) < TOL
    # row-block sharding == slicing rows of A
    Kb = kernels.global_kernels(At[10:30], Bt, sig, "gaussian").cpu().numpy()
    Kf = kernels.global_kernels(At, Bt, sig, "gaussian").cpu().numpy()
    assert np.array_equal(Kb, Kf[:, 10:30])


@pytest.mark.parametrize("iab", [False, True])
def test_local_gaussian_synthetic(slatm_set, iab):
    s = slatm_set
    n1 = 40
    off = np.concatenate(([0], np.cumsum(s["nas"])))
    A1 = off[n1]
    x1, x2 = s["xl"][:A1], s["xl"][A1:]
    zs1, zs2 = s["zs"][:A1], s["zs"][A1:]
    nas1, nas2 = s["nas"][:n1], s["nas"][n1:]
    zmax = 9
    dso = co.get_kernel_width(s["nas"], s["zs"], s["xl"])
    dso = np.concatenate((dso, np.full(zmax - len(dso), 1e-9)))
    sig = np.array([c * dso / np.sqrt(2 * np.log(2)) for c in (0.5, 1.0, 2.0, 4.0)])
    if iab:
        sig = np.maximum(sig, 5.0)
    for (a, za, na, b, zb, nb) in ((x2, zs2, nas2, x1, zs1, nas1), (x1, zs1, nas1, x1, zs1, nas1)):
        K = kernels.get_local_kernels_gaussian(a, b, za, zb, na, nb, iab, zmax, sig)
        ref = co.calc_lk_gaussian(a, b, za, zb, na, nb, zmax, iab, sig)
        assert K.shape == ref.shape == (4, len(na), len(nb))
        assert kernel_rel_err(K, ref) < TOL
        Knc = kernels.get_local_kernels_gaussian(a, b, za, zb, na, nb, iab, zmax, sig, center=False)
        assert kernel_rel_err(Knc, ref) < TOL


def test_local_gaussian_arbitrary_sigmas(slatm_set):
    """Sigmas that are NOT power-of-two multiples of each other (no squaring chain), per-element ratios differ."""
    s = slatm_set
    n1 = 25
    off = np.concatenate(([0], np.cumsum(s["nas"])))
    A1 = off[n1]
    x1, x2 = s["xl"][:A1], s["xl"][A1:off[n1 + 20]]
    zs1, zs2 = s["zs"][:A1], s["zs"][A1:off[n1 + 20]]
    nas1, nas2 = s["nas"][:n1], s["nas"][n1:n1 + 20]
    rng = np.random.default_rng(3)
    sig = 6.0 + 20.0 * rng.random((5, 9))
    for center in (True, False):
        K = kernels.get_local_kernels_gaussian(x2, x1, zs2, zs1, nas2, nas1, False, 9, sig, center=center)
        assert kernel_rel_err(K, co.calc_lk_gaussian(x2, x1, zs2, zs1, nas2, nas1, 9, False, sig)) < TOL
    # a ladder with a factor-sqrt(2) step (isig ratio 2: one squaring) and a repeated sigma
    base = 5.0 + rng.random(9)
    sig = np.array([base * 4, base * 2 * np.sqrt(2.0), base * 2, base * 2, base])
    K = kernels.get_local_kernels_gaussian(x2, x1, zs2, zs1, nas2, nas1, False, 9, sig, center=False)
    assert kernel_rel_err(K, co.calc_lk_gaussian(x2, x1, zs2, zs1, nas2, nas1, 9, False, sig)) < TOL
    Kl = kernels.get_local_kernels_laplacian(x2, x1, nas2, nas1, [300., 777., 1500.])
    assert kernel_rel_err(Kl, co.calc_lk_laplacian(x2, x1, nas2, nas1, [300., 777., 1500.])) < TOL


def test_local_gaussian_wide_sigma_mismatch_term(slatm_set):
    """Element-mismatched pairs carry d2 = 1e9 (fkernels.f90:73): visible for sigma >~ 1e3."""
    s = slatm_set
    n = 6
    A = int(s["nas"][:n].sum())
    x, zs, nas = s["xl"][:A], s["zs"][:A], s["nas"][:n]
    sig = np.full((2, 9), 3000.0)
    sig[1] = 9000.0
    K = kernels.get_local_kernels_gaussian(x, x, zs, zs, nas, nas, False, 9, sig)
    assert kernel_rel_err(K, co.calc_lk_gaussian(x, x, zs, zs, nas, nas, 9, False, sig)) < TOL


def test_local_gaussian_golden_demo(demo):
    n1 = int(demo["n1"])
    nas, zs, x = demo["nas"], demo["zs"], demo["x_local"]
    A1 = int(nas[:n1].sum())
    k1 = kernels.get_local_kernels_gaussian(x[:A1], x[:A1], zs[:A1], zs[:A1], nas[:n1], nas[:n1], False, 8, demo["sigmas"])
    k2 = kernels.get_local_kernels_gaussian(x[A1:], x[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], False, 8, demo["sigmas"])
    assert kernel_rel_err(k1, demo["k1"]) < TOL
    assert kernel_rel_err(k2, demo["k2"]) < TOL


def test_local_laplacian(slatm_set):
    s = slatm_set
    n1 = 12
    off = np.concatenate(([0], np.cumsum(s["nas"])))
    x1, x2 = s["xl"][:off[n1]], s["xl"][off[n1]:off[n1 + 9]]
    nas1, nas2 = s["nas"][:n1], s["nas"][n1:n1 + 9]
    sig = [200., 800., 3200.]
    K = kernels.get_local_kernels_laplacian(x2, x1, nas2, nas1, sig)
    assert kernel_rel_err(K, co.calc_lk_laplacian(x2, x1, nas2, nas1, sig)) < TOL


def test_kernel_width(demo, slatm_set):
    dso = algo.get_kernel_width(demo["nas"], demo["zs"], demo["x_local"])
    assert np.allclose(dso, demo["dso"], rtol=1e-9, atol=0)
    s = slatm_set
    ref = co.get_kernel_width(s["nas"], s["zs"], s["xl"])
    assert np.allclose(algo.get_kernel_width(s["nas"], s["zs"], s["xl"]), ref, rtol=1e-10, atol=0)
    ref1 = co.get_kernel_width(s["nas"], s["zs"], s["xl"], kernel='l')
    assert np.allclose(algo.get_kernel_width(s["nas"], s["zs"], s["xl"], kernel='l'), ref1, rtol=1e-12, atol=0)


def test_ragged_and_empty(slatm_set):
    s = slatm_set
    x = s["xl"][:int(s["nas"][:3].sum())]
    zs, nas = s["zs"][:len(x)], s["nas"][:3]
    sig = np.full((1, 9), 10.0)
    # molecule with zero atoms in the middle
    nas0 = np.array([nas[0], 0, nas[1] + nas[2]], dtype=np.int32)
    K = kernels.get_local_kernels_gaussian(x, x, zs, zs, nas0, nas, False, 9, sig)
    ref = co.calc_lk_gaussian(x, x, zs, zs, nas0, nas, 9, False, sig)
    assert kernel_rel_err(K, ref) < TOL and np.all(K[0, 1] == 0)
    with pytest.raises(AssertionError):
        kernels.get_local_kernels_gaussian(x, x, zs, zs, nas, nas[:2], False, 9, sig)
    from aqml_b200._lib import AqmlError
    with pytest.raises(AqmlError):
        kernels.get_local_kernels_gaussian(x, x, zs + 20, zs, nas, nas, False, 9, sig)


def test_linear_matern_sargan(slatm_set):
    """cml.kernels.{linear,matern,sargan}_kernel (kernels.py:69-191; fkernels.f90:390-518) vs the restated formulas."""
    X = slatm_set["xg"]
    A, B = X[:21], X[10:47]
    d1, d2 = co.l1_distance(A, B), co.l2_distance(A, B)
    K = kernels.linear_kernel(A, B)
    ref = A @ B.T
    assert K.shape == (21, 37) and K.flags.f_contiguous and kernel_rel_err(K, ref) < 1e-13
    s2, s1 = d2.max(), d1.max()
    for order in (0, 1, 2):
        c = [1.0, np.sqrt(3.0), np.sqrt(5.0)][order]
        poly2 = [1.0, 1 + c * d2 / s2, 1 + c * d2 / s2 + 5.0 / (3 * s2 * s2) * d2 * d2][order]
        assert kernel_rel_err(kernels.matern_kernel(A, B, s2, order, "l2"), np.exp(-c * d2 / s2) * poly2) < 1e-12
        x = c * d1 / s1
        poly1 = [1.0, 1 + x, 1 + x + x * x / 3.0][order]
        assert kernel_rel_err(kernels.matern_kernel(A, B, s1, order, "l1"), np.exp(-x) * poly1) < 1e-12
    g = [0.3, 0.05, 0.7]
    x = d1 / s1
    assert kernel_rel_err(kernels.sargan_kernel(A, B, s1, g), np.exp(-x) * (1 + g[0] * x + g[1] * x ** 2 + g[2] * x ** 3)) < 1e-12
    with pytest.raises(SystemExit):
        kernels.matern_kernel(A, B, 1.0, 0, "l3")


def test_vector_kernels_on_padded_arrays(slatm_set):
    """fkernels.f90:185-324 (f2py-level names): zero-padded (nm, maxatoms, D) inputs, all atom pairs."""
    s = slatm_set
    off = np.concatenate(([0], np.cumsum(s["nas"])))
    nas1, nas2 = s["nas"][:7], s["nas"][7:12]
    x1, x2 = s["xl"][:off[7]], s["xl"][off[7]:off[12]]

    def pad(x, nas):
        out = np.zeros((len(nas), int(nas.max()), x.shape[1]))
        o2 = np.concatenate(([0], np.cumsum(nas)))
        for i, n in enumerate(nas):
            out[i, :n] = x[o2[i]:o2[i + 1]]
        return out

    p1, p2 = pad(x1, nas1), pad(x2, nas2)
    sig = [20.0, 40.0]
    Kg = kernels.calc_vector_kernels_gaussian(p1.T, p2.T, nas1, nas2, sig)  # .T = what f2py receives
    z1, z2 = np.ones(len(x1), np.int32), np.ones(len(x2), np.int32)
    ref = co.calc_lk_gaussian(x1, x2, z1, z2, nas1, nas2, 1, True, np.array(sig).reshape(-1, 1))
    assert kernel_rel_err(Kg, ref) < TOL
    Kl = kernels.calc_vector_kernels_laplacian(p1, p2, nas1, nas2, [500., 900.])
    assert kernel_rel_err(Kl, co.calc_lk_laplacian(x1, x2, nas1, nas2, [500., 900.])) < TOL


def test_int8_engine_hostile_rows():
    """Rows that stress the digit-plane representation of the tcgen05 engine: 16 orders of magnitude inside one row,
    rows of very different scale, denormal-sized and exactly-zero rows, negative values, exact powers of two and
    half-integers at digit boundaries, duplicated rows (d2 = 0).  Global and local kernels, both engines vs the oracle."""
    import os
    rng = np.random.default_rng(99)
    n1, n2, D = 70, 150, 300
    def rows(n):
        X = rng.normal(size=(n, D)) * 10.0 ** rng.integers(-8, 3, size=(n, D))        # wide dynamic range within rows
        X[:, ::7] = 0.0
        X[1] *= 1e-6; X[2] *= 1e5; X[3] = 0.0; X[4] = 1e-310; X[5, :40] = 2.0 ** rng.integers(-30, 10, size=40)
        X[6, :40] = (rng.integers(-64, 65, size=40) + 0.5) / 128.0                      # ties in the first digit
        return X
    A, B = rows(n1), rows(n2)
    B[10] = A[0]; B[11] = A[2]
    dmax = co.l2_distance(A, B).max()
    for sig in (0.05 * dmax, dmax):
        ref = co.gaussian_kernel(A, B, sig)
        for engine in ("i8", "dmma"):
            os.environ["AQML_PAIR_ENGINE"] = engine
            try:
                K = kernels.gaussian_kernel(A, B, sig)
            finally:
                del os.environ["AQML_PAIR_ENGINE"]
            assert kernel_rel_err(K, ref) < TOL, (engine, sig)
    # local kernel on the same rows: 3 "elements", ragged molecules
    zs1 = rng.choice([1, 6, 8], size=n1); zs2 = rng.choice([1, 6, 8], size=n2)
    nas1 = np.array([3, 7, 1, 9, 20, 30]); nas2 = np.array([50, 1, 2, 40, 57])
    sigmas = np.array([[c * dmax] * 8 for c in (0.07, 0.5, 3.0)])
    ref = co.calc_lk_gaussian(A, B, zs1, zs2, nas1, nas2, 8, False, sigmas)
    for engine in ("i8", "dmma"):
        os.environ["AQML_PAIR_ENGINE"] = engine
        try:
            K = kernels.get_local_kernels_gaussian(A, B, zs1, zs2, nas1, nas2, False, 8, sigmas)
            K0 = kernels.get_local_kernels_gaussian(A, B, zs1, zs2, nas1, nas2, False, 8, sigmas, center=False)
        finally:
            del os.environ["AQML_PAIR_ENGINE"]
        assert kernel_rel_err(K, ref) < TOL and kernel_rel_err(K0, ref) < TOL, engine
    # more than 8 sigmas: served by the FP64 engine automatically
    many = np.array([[c * dmax] * 8 for c in np.linspace(0.1, 2.0, 11)])
    K = kernels.get_local_kernels_gaussian(A, B, zs1, zs2, nas1, nas2, False, 8, many)
    assert kernel_rel_err(K, co.calc_lk_gaussian(A, B, zs1, zs2, nas1, nas2, 8, False, many)) < TOL
    # NaN / Inf inputs fall back to the FP64 engine and propagate as in the reference
    A2 = A.copy(); A2[7, 3] = np.nan
    K = kernels.gaussian_kernel(A2, B, dmax)
    assert np.isnan(K[7]).all() and not np.isnan(np.delete(K, 7, axis=0)).any()


def test_global_gaussian_very_long_rows():
    """More 16-column k-chunks than the tile metadata's chunk list holds (D > 16384): the int8 engine walks them all."""
    rng = np.random.default_rng(5)
    A, B = rng.random((70, 20011)), rng.random((45, 20011))
    sigma = 0.4 * np.sqrt(20011)
    assert kernel_rel_err(kernels.gaussian_kernel(A, B, sigma), co.gaussian_kernel(A, B, sigma)) < TOL


def test_wide_power_of_two_sigma_ladder():
    """sigma = [s, 256 s, 65536 s] are exact power-of-two multiples, but chaining them by squarings would need
    32 squarings (error 2^32 ulp); plan_sigma_chain must refuse (<= 8 squarings in all) and give every sigma its own exp.
    Also values so close to 1 that a squaring chain would collapse them to exactly 1."""
    rng = np.random.default_rng(3)
    A, B = rng.random((70, 300)), rng.random((90, 300))
    s0 = 3.0
    for sig in ([s0, 256 * s0, 65536 * s0], [s0, 4 * s0, 64 * s0, 1024 * s0], [s0, 2 * s0, 4 * s0, 8 * s0, 16 * s0, 32 * s0]):
        K = kernels.gaussian_kernels(A, B, sig)
        for i, s in enumerate(sig):
            assert kernel_rel_err(K[i], co.gaussian_kernel(A, B, s)) < TOL, (sig, i)
    zs1, zs2 = np.full(70, 6), np.full(90, 6)
    nas1, nas2 = np.array([30, 40]), np.array([45, 45])
    for ladder in ([1.0, 256.0, 65536.0], [1.0, 2.0, 4.0, 8.0, 16.0, 32.0, 64.0]):
        sig = np.array([[c * s0] * 6 for c in ladder])
        K = kernels.get_local_kernels_gaussian(A, B, zs1, zs2, nas1, nas2, False, 6, sig)
        assert kernel_rel_err(K, co.calc_lk_gaussian(A, B, zs1, zs2, nas1, nas2, 6, False, sig)) < TOL
