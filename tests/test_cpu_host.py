"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, fails loudly without a device, and the host logic (mbtypes, grids) matches the oracle."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, mbtypes_from_array
from aqml_b200 import _lib as L
from aqml_b200 import synth
from aqml_b200.representation import core as rcore
from oracle import np_oracle as o


def header_symbols():
    src = open(os.path.join(ROOT, "include", "aqml_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aqml_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    lib = L.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/aqml_b200.h but not exported"
        assert s in L.SIGNATURES, f"{s} has no ctypes signature"
    assert set(L.SIGNATURES) == set(syms)
    assert lib.aqml_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from aqml_b200 import kernels
    with pytest.raises(L.AqmlError, match="no CUDA device"):
        kernels.gaussian_kernel(np.ones((3, 4)), np.ones((2, 4)), 1.0)
    with pytest.raises(L.AqmlError, match="no CUDA device"):
        rcore.generate_slatm(np.zeros((2, 3)), [1, 1], [[1], [1, 1]])


def test_product_does_not_import_oracle():
    import subprocess
    import sys
    code = ("import sys; import aqml_b200, aqml_b200.kernels, aqml_b200.distance, aqml_b200.representation, "
            "aqml_b200.fused, aqml_b200.algo.aqml, aqml_b200.parallel, cml.kernels; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "aqml_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_mbtypes_match_oracle_and_golden(demo):
    nas, zs = demo["nas"], demo["zs"]
    zs_list = synth.split_zs(nas, zs)
    assert rcore.get_slatm_mbtypes(zs_list) == mbtypes_from_array(demo["mbtypes"])
    nas2, zs2, _ = synth.qm9_like(40, 3)
    zl = synth.split_zs(nas2, zs2)
    assert rcore.get_slatm_mbtypes(zl) == o.get_slatm_mbtypes(zl)
    assert rcore.get_slatm_mbtypes(zl, pbc=True) == o.get_slatm_mbtypes(zl, pbc=True)


def test_slatm_params_grids():
    mb = [[1], [6], [1, 1], [1, 6], [6, 6], [1, 6, 1]]
    p = L.slatm_params(mb, sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
    assert (p.nx2, p.nx3) == (118, 96) == (o.grid2(4.8, .04), o.grid3(.04))
    p = L.slatm_params(mb)
    assert (p.nx2, p.nx3) == (157, 128)
    assert L.load().aqml_slatm_dim(p) == 2 + 3 * 157 + 128
    with pytest.raises(L.AqmlError):
        L.slatm_params([[1, 2, 3, 4]])


def test_synth_is_deterministic():
    a = synth.qm9_like_fast(700, 5, pool=64)
    b = synth.qm9_like_fast(700, 5, pool=64)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    nas, zs, co = a
    assert nas.sum() == len(zs) == len(co) and nas.max() <= 29
    assert set(np.unique(zs)) <= {1, 6, 7, 8, 9}


def test_tile_shape_and_counters_without_device():
    import ctypes as C
    lib = L.load()
    bm, bn, bk = C.c_int(0), C.c_int(0), C.c_int(0)
    lib.aqml_tile_shape(C.c_void_p(C.addressof(bm)), C.c_void_p(C.addressof(bn)), C.c_void_p(C.addressof(bk)))
    assert (bm.value, bn.value, bk.value) == (64, 64, 16)
    assert lib.aqml_launch_count() >= 0 and lib.aqml_device_count() >= 0


def test_krr_helpers_on_golden(demo):
    """Host-side consumers (aqml.py:129-181,806-897; rkrr.py:203-228) against the golden kernels: the README curve."""
    from aqml_b200.algo import krr
    n1 = int(demo["n1"])
    nas, zs = demo["nas"], demo["zs"]
    zs_list = synth.split_zs(nas, zs)
    zsu = np.unique(zs)
    ngs = np.array([[(z == zz).sum() for zz in zsu] for z in zs_list])
    nheav = np.array([(z > 1).sum() for z in zs_list[:n1]])
    ys = demo["energies"] * krr.H2KC
    free = [krr.H2KC * o.ORCA_B3LYPVDZ[int(z)] for z in zsu]
    rows = krr.learning_curve(demo["k1"][1], demo["k2"][1], ngs, ys, nheav, free, 1e-4)
    for r, ref in zip(rows, demo["readme_curve"]):
        assert (r[0], r[1]) == (ref[0], ref[1]) and abs(r[2] - ref[2]) < 5.1e-5


def test_reference_file_formats(tmp_path, monkeypatch):
    """Same file names / npz keys as `aqml -savex -savek` (aqml.py:626-665,719,736,795)."""
    from aqml_b200 import io as aio
    monkeypatch.chdir(tmp_path)
    assert aio.stem("g7/") == "data/g7" and aio.stem("a/b", rcut=3.0, z=6) == "data/a-b-rc3.0-z6"
    x1 = np.arange(12.0).reshape(3, 4)
    aio.save_x1("g7/", x1)
    assert os.path.exists("data/g7-x1.npz") and np.array_equal(np.load("data/g7-x1.npz")["x1"], x1)
    assert np.array_equal(aio.load_x1("g7"), x1)
    ks1, ks2 = np.ones((2, 3, 3)), np.zeros((2, 1, 3))
    aio.save_kernels("g7/", ks1, ks2, kernel="l")
    assert os.path.exists("data/g7-kernels-l.npz")
    a, b = aio.load_kernels("g7", kernel="l")
    assert np.array_equal(a, ks1) and np.array_equal(b, ks2)


def test_block_driver_job_ids():
    """slatm_x.py:147-184: job enumeration order and sub-job selection."""
    from aqml_b200.representation.slatm_x import all_jobs, select_jobs, MolSet
    assert all_jobs(2, 1) == ['11ii001001', '11ij001002', '12ij001001', '11ii002002', '12ij002001']
    jobs, wko = select_jobs('a', 2, 1)
    assert jobs == all_jobs(2, 1) and not wko
    jobs, wko = select_jobs(['12ij', '11ii002002'], 2, 1)
    assert jobs == ['12ij001001', '12ij002001', '11ii002002'] and wko
    ms = MolSet([1, 1, 8, 6, 1, 1, 1, 1], np.zeros((8, 3)), [3, 5], 1)
    assert ms.nzs.tolist() == [[2, 0, 1], [4, 1, 0]] and len(ms.nas1) == 1 and len(ms.nas2) == 1


def _digit_planes(X):
    """i8_rowexp_kernel + i8_slice_kernel (aqml_b200/csrc/i8pair.cuh) in NumPy: six signed 8-bit digits per element."""
    m = np.abs(X).max(axis=1)
    f, e0 = np.frexp(m)
    e = np.where(m > 0, e0 + np.where(f < 0.996, 1, 2), 0)          # |x| 2^-e < 0.498
    t = X * np.exp2(-e.astype(float))[:, None]
    qs = []
    for _ in range(6):
        u = t * 256.0
        q = np.clip(np.floor(u + 128.0 / 255.0), -128, 127)
        t = u - q                                                    # exact, in [-128/255, 127/255)
        qs.append(q.astype(np.int64))
    return qs, e, t


def _digit_dot(qa, ea, qb, eb, full6=False):
    """The 22 (26) plane products, exact integer accumulation per level, the epilogue's hi / lo recombination."""
    keep = (lambda i, j: i + j <= 6) if full6 else (lambda i, j: i + j <= 5 or (i == 3 and j == 3))
    P = [sum(qa[i] @ qb[g - i].T for i in range(6) if 0 <= g - i < 6 and keep(i, g - i)) for g in range(7)]
    assert all(np.abs(p).max() < 2 ** 31 for p in P)
    hi = (P[0] << 16) + (P[1] << 8) + P[2]
    lo = (P[3] << 24) + (P[4] << 16) + (P[5] << 8) + P[6]
    d = lo.astype(np.float64) * 2.0 ** -32 + hi.astype(np.float64)
    return d * np.exp2((ea[:, None] + eb[None, :] - 32).astype(float))


def test_digit_plane_algebra():
    """The arithmetic behind the int8 tensor engine (aqml_b200/csrc/i8pair.cuh), restated in NumPy: 6 signed 8-bit digit
    planes per row, 22 plane products (levels i + j <= 5 and the square (3,3)), exact integer accumulation, hi / lo
    recombination -- against the FP64 dot product; every digit must fit an int8."""
    rng = np.random.default_rng(11)
    n, D = 40, 1500
    X = rng.normal(size=(n, D)) * 10.0 ** rng.integers(-6, 2, size=(n, D))
    X[:, ::5] = 0.0
    X[3] = 0.0
    X[4, :64] = (rng.integers(-128, 128, size=64) + 0.5) / 256.0 * 3.9   # ties at digit boundaries
    X[5] = X[6] * (1 + 1e-9)                                              # a ~ b: the level-6 square does not average out
    qs, e, resid = _digit_planes(X)
    assert all(q.min() >= -128 and q.max() <= 127 for q in qs)
    assert np.all(resid >= -128.0 / 255.0 - 1e-12) and np.all(resid < 127.0 / 255.0 + 1e-12)
    recon = sum(q * 256.0 ** -(s + 1) for s, q in enumerate(qs)) * np.exp2(e.astype(float))[:, None]
    assert np.all(np.abs(recon - X) <= 0.503 * np.exp2(e.astype(float) - 48)[:, None])
    exact = X @ X.T
    scale = np.sqrt(np.outer((X * X).sum(1), (X * X).sum(1)))       # |a||b|
    err22 = np.abs(_digit_dot(qs, e, qs, e) - exact) / np.maximum(scale, 1e-300)
    err26 = np.abs(_digit_dot(qs, e, qs, e, full6=True) - exact) / np.maximum(scale, 1e-300)
    assert err22.max() <= 2e-13 and err26.max() <= 2e-14, (err22.max(), err26.max())
    # without the (3,3) product the diagonal (a = b) is biased: that is why it is kept
    P33 = (qs[3] * qs[3]).sum(axis=1) * 256.0 ** -8 * np.exp2(2.0 * e)
    assert np.median(P33[scale.diagonal() > 0] / scale.diagonal()[scale.diagonal() > 0]) > 1e-14


def test_shape_validation_before_any_device_call():
    """mismatched coords / zs / nas must raise in Python -- the C ABI only sees raw pointers."""
    from aqml_b200 import fused, kernels
    x = np.zeros((5, 8))
    with pytest.raises(ValueError):
        kernels.get_local_kernels_gaussian(x, x, [1, 1, 1], [1, 1, 1, 1, 1], [5], [5], False, 1, np.ones((1, 1)))
    with pytest.raises(ValueError):
        fused.PackedRep(np.zeros((4, 3)), [1, 1, 1, 1, 1], [5], [[1], [1, 1]])
    with pytest.raises(ValueError):
        fused.PackedRep(np.zeros((5, 3)), [1, 1, 1, 1], [5], [[1], [1, 1]])


def test_pairwise_oracle_internal_consistency():
    """oracle/np_pairwise.py (fslatm_pairwise.f90, parity unpinned): with every atom a listed neighbour and equal cut-offs /
    grids, an atom's two- and three-body blocks are the sums of its bond rows' blocks -- the relation between
    calc_sbop_local / calc_sbot_local with ib = 0 and ib = j that the file's two code paths must satisfy."""
    from aqml_b200 import synth
    from oracle import np_pairwise as o
    nas, zs, coords = synth.qm9_like(2, 3)
    nas, zs, coords = np.asarray(nas), np.asarray(zs), np.asarray(coords)
    z, c = zs[:nas[0]], coords[:nas[0]]
    mb = o.get_slatm_mbtypes(zs, nas)
    rc = 3.0
    _, nbrs = o.get_neighbors(100.0, c)
    nsx = o.grid_sizes(rc, rc, 0.05)
    ys, ys2 = o.calc_local_spectrum(z, c, nbrs, mb[0], mb[1], mb[2], nsx, [rc, rc], 0.05, 0.06, 1.0, 6, 1 / 3, 3)
    n1 = len(mb[0])
    assert np.array_equal(ys[:, :n1] != 0, z[:, None] == mb[0][None, :]) and np.all(ys2[:, :, :n1] == 0)
    assert np.max(np.abs(ys[:, n1:] - ys2[:, :, n1:].sum(axis=1))) <= 1e-14 * np.max(np.abs(ys))
    assert abs(o.R0 - 0.1) > 1e-9 and abs(o.R0 - 0.1) < 2e-9   # r0 is the single-precision 0.1


def test_integer_exp_of_the_int8_epilogue(tmp_path):
    """aqml_b200/csrc/i8exp.cuh (host-callable): 2^-T by table + polynomial on integers, the sigma ladder as shifts, aligned
    64-bit sums -- against exp2l in extended precision (tools/micro/i8exp_check.cpp: worst relative error < 2e-12, specials)."""
    import subprocess
    exe = str(tmp_path / "i8exp_check")
    subprocess.check_call(["g++", "-O2", "-o", exe, os.path.join(ROOT, "tools", "micro", "i8exp_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "T = inf -> 0" in out.stdout and "T = 0 -> 1" in out.stdout


def test_fixed_point_d2_of_the_int8_epilogue():
    """The integer d^2 / T arithmetic of i8_pair_kernel's epilogue (i8pair.cuh: metadata warp + `M.fix` path), restated with
    Python integers: per tile the norms become 64-bit integers in units of 2^(E - 60) (E: exponent of the largest norm),
    2 a.b = W 2^(ea + eb - 47) is shifted into the same units, d^2 = NA + NB - Ws is clamped at 0, and
    T 2^48 = mulhi64(d^2, Kq) >> kr with K = c 2^(E - 12) = Kq 2^-(64 + kr).  Checked against float64 d^2 * c for rows of very
    different scale, identical rows (d^2 = 0 up to the digits' truncation) and the range conditions of the metadata warp."""
    import struct
    rng = np.random.default_rng(5)

    def bits(x):
        return struct.unpack("<Q", struct.pack("<d", float(x)))[0]

    def norm_fix(v, E):          # the metadata warp: m 2^(e - 1075) -> m 2^(e - 1015 - E), truncated
        b = bits(v)
        e, m = b >> 52, (b & ((1 << 52) - 1)) | (1 << 52)
        if e == 0:
            return 0
        s = e - 1015 - E
        return (m << s) if s >= 0 else (m >> min(-s, 63))

    worst = 0.0
    for trial in range(300):
        n, D = 6, 40
        scale = 10.0 ** rng.integers(-3, 4)
        X = rng.normal(size=(n, D)) * scale * 10.0 ** rng.integers(-2, 2, size=(n, 1))
        X[1] = X[0]                                   # identical rows
        na = (X * X).sum(axis=1)
        # exact integers W = round(a.b 2^(48 - ea - eb)) stand for the digit-plane products (their own truncation is tested elsewhere)
        ex = [int(np.frexp(np.abs(r).max())[1]) + 1 for r in X]
        E = ((bits(na.max()) >> 52) & 0x7ff) - 1022
        c = 1.4426950408889634 / (2.0 * (0.3 * np.sqrt(na.max())) ** 2)
        K = c * 2.0 ** (E - 12)
        eK = (bits(c) >> 52) + (E - 12)
        assert 1022 - 63 <= eK <= 1021 and 0.0 < K < 0.5
        kr = 1022 - eK
        Kq = ((bits(c) & ((1 << 52) - 1)) | (1 << 52)) << 11
        assert abs(Kq * 2.0 ** -(64 + kr) - K) <= K * 2.0 ** -52
        NA = [norm_fix(v, E) for v in na]
        for i in range(n):
            for j in range(n):
                ab = float(np.dot(X[i], X[j]))
                W = int(round(ab * 2.0 ** (48 - ex[i] - ex[j])))
                sh = ex[i] + ex[j] + 13 - E
                Ws = (W << sh) if sh >= 0 else (W >> min(-sh, 63))
                Dq = max(NA[i] + NA[j] - Ws, 0)
                assert Dq < 2 ** 63
                F = ((Dq * Kq) >> 64) >> kr
                T = F * 2.0 ** -48
                d2 = max(na[i] + na[j] - 2.0 * ab, 0.0)
                ref = d2 * c
                # error budget: the norms' own rounding (as in the FP64 path) ~ 1e-16 (na + nb) c, the fixed-point units 2^-60 of the largest norm
                tol = 4e-16 * (na[i] + na[j]) * c + 2.0 ** -47 + 2.0 ** (ex[i] + ex[j] - 47) * c   # + the rounding of W itself
                assert abs(T - ref) <= tol, (trial, i, j, T, ref)
                worst = max(worst, abs(T - ref))
    assert worst < 1e-12
