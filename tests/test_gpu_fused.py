"""GPU: fused device-resident path and the KRR known-answer test through the CUDA code."""
import numpy as np
import pytest

from conftest import kernel_rel_err, mbtypes_from_array, rel_err
from aqml_b200 import fused, kernels, synth
from aqml_b200.algo import aqml as algo
from aqml_b200.representation import core as rcore
from oracle import c_oracle as co
from oracle import np_oracle as o

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_fused_equals_two_step_and_oracle():
    import torch
    nas, zs, coords = synth.qm9_like(70, 31)
    n1 = 50
    off = np.concatenate(([0], np.cumsum(nas)))
    A1 = off[n1]
    mb = o.get_slatm_mbtypes(synth.split_zs(nas, zs))
    kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
    xl = co.generate_slatm_batch(coords, zs, nas, mb, True, **kw)
    zmax = 9
    dso = co.get_kernel_width(nas, zs, xl)
    dso = np.concatenate((dso, np.full(zmax - len(dso), 1e-9)))
    sig = np.array([c * dso / np.sqrt(2 * np.log(2)) for c in (0.5, 1.0, 2.0, 4.0)])
    ref = co.calc_lk_gaussian(xl[A1:], xl[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], zmax, False, sig)
    train = fused.PackedRep(coords[:A1], zs[:A1], nas[:n1], mb, **kw)
    test = fused.PackedRep(torch.as_tensor(coords[A1:], device="cuda"), zs[A1:], nas[n1:], mb, **kw)
    K = fused.local_gaussian(test, train, sig, zmax).cpu().numpy()
    assert kernel_rel_err(K, ref) < TOL
    # row-block sharding: rows 5..12 of the test set
    Kb = fused.local_gaussian(test, train, sig, zmax, row0=5, nrows=7).cpu().numpy()
    assert kernel_rel_err(Kb, ref[:, 5:12]) < TOL
    # symmetric K(train, train): half the contraction + mirror == the rectangular route == the oracle
    ref11 = co.calc_lk_gaussian(xl[:A1], xl[:A1], zs[:A1], zs[:A1], nas[:n1], nas[:n1], zmax, False, sig)
    K11 = fused.local_gaussian(train, train, sig, zmax).cpu().numpy()
    K11s = fused.local_gaussian(train, train, sig, zmax, symmetric=True).cpu().numpy()
    assert kernel_rel_err(K11, ref11) < TOL and kernel_rel_err(K11s, ref11) < TOL
    assert np.array_equal(K11s, np.transpose(K11s, (0, 2, 1)))
    # width heuristic on packed rows == dense get_kernel_width on train+test
    both = fused.PackedRep(coords, zs, nas, mb, **kw)
    assert np.allclose(fused.kernel_width(both, both, zmax), dso, rtol=1e-10)
    # device two-step route
    xd = rcore.generate_slatm_batch(torch.as_tensor(coords, device="cuda"), zs, nas, mb, local=True,
                                    sigmas=[.05, .05], dgrids=[.04, .04], rcut=4.8)
    K2 = kernels.get_local_kernels_gaussian(xd[A1:], xd[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], False, zmax, sig)
    assert kernel_rel_err(K2.cpu().numpy(), ref) < TOL


def test_demo_learning_curve_through_gpu(demo):
    """The reference's only KAT (demo/example/README.md:57-63), every stage on the GPU path:
    aSLATM -> per-element dmax -> sigma -> local Gaussian kernels; KRR solve on the host as upstream."""
    n1 = int(demo["n1"])
    nas, zs, coords = demo["nas"], demo["zs"], demo["coords"]
    mb = mbtypes_from_array(demo["mbtypes"])
    off = np.concatenate(([0], np.cumsum(nas)))
    zs_list = synth.split_zs(nas, zs)
    x = np.concatenate([np.array(rcore.generate_slatm(coords[off[i]:off[i + 1]], zs_list[i], mb, local=True,
                                                       sigmas=[.05, .05], dgrids=[.04, .04], rcut=4.8))
                        for i in range(len(nas))])
    A1 = off[n1]
    dsmax = algo.get_kernel_width(nas, zs, x)
    fun = algo.fobj('g')
    sigmas = np.array([fun._factor * dsmax * 1.0])
    zmax = int(zs.max())
    ks1 = fun.k(x[:A1], x[:A1], zs[:A1], zs[:A1], nas[:n1], nas[:n1], False, zmax, sigmas)
    ks2 = fun.k(x[A1:], x[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], False, zmax, sigmas)
    zsu = np.unique(zs)
    ngs = np.array([[(z == zz).sum() for zz in zsu] for z in zs_list])
    nheav = np.array([(z > 1).sum() for z in zs_list[:n1]])
    ys = demo["energies"] * o.H2KC
    rows = o.learning_curve(ks1[0], ks2[0], ngs, ys, nheav, zsu)
    readme = demo["readme_curve"]
    for r, ref in zip(rows, readme):
        assert (r[0], r[1]) == (ref[0], ref[1])
        assert abs(r[2] - ref[2]) < 5.1e-5 and abs(r[3] - ref[3]) < 5.1e-5
    # predictions within 1e-8 Ha of the oracle's own pipeline
    rows_ref = o.learning_curve(demo["k1"][1], demo["k2"][1], ngs, ys, nheav, zsu)
    for r, rr in zip(rows, rows_ref):
        assert np.max(np.abs(r[5] - rr[5])) < 1e-8 * o.H2KC


def test_int8_tensor_engine_equals_fp64_engine_and_oracle():
    """The tcgen05 int8 digit-plane engine (default) and the FP64 DMMA engine (AQML_PAIR_ENGINE=dmma) on the SAME packed
    representation: both within 1e-10 of the oracle, each other within 1e-11; arbitrary (non power-of-two) sigma ratios,
    row ranges and the symmetric path included; values spanning many orders of magnitude (narrow sigma)."""
    import os
    nas, zs, coords = synth.qm9_like(90, 55)
    n1 = 64
    off = np.concatenate(([0], np.cumsum(nas)))
    A1 = off[n1]
    mb = o.get_slatm_mbtypes(synth.split_zs(nas, zs))
    kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
    xl = co.generate_slatm_batch(coords, zs, nas, mb, True, **kw)
    zmax = 9
    dso = co.get_kernel_width(nas, zs, xl)
    dso = np.concatenate((dso, np.full(zmax - len(dso), 1e-9)))
    for coeffs in ((0.5, 1.0, 2.0, 4.0), (0.05, 0.3, 1.7)):
        sig = np.array([c * dso / np.sqrt(2 * np.log(2)) for c in coeffs])
        ref = co.calc_lk_gaussian(xl[A1:], xl[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], zmax, False, sig)
        ref11 = co.calc_lk_gaussian(xl[:A1], xl[:A1], zs[:A1], zs[:A1], nas[:n1], nas[:n1], zmax, False, sig)
        train = fused.PackedRep(coords[:A1], zs[:A1], nas[:n1], mb, **kw)
        test = fused.PackedRep(coords[A1:], zs[A1:], nas[n1:], mb, **kw)
        out = {}
        for engine in ("i8", "dmma"):
            os.environ["AQML_PAIR_ENGINE"] = engine
            try:
                out[engine] = (fused.local_gaussian(test, train, sig, zmax).cpu().numpy(),
                               fused.local_gaussian(test, train, sig, zmax, row0=3, nrows=11).cpu().numpy(),
                               fused.local_gaussian(train, train, sig, zmax, symmetric=True).cpu().numpy())
            finally:
                del os.environ["AQML_PAIR_ENGINE"]
            assert kernel_rel_err(out[engine][0], ref) < TOL
            assert kernel_rel_err(out[engine][1], ref[:, 3:14]) < TOL
            assert kernel_rel_err(out[engine][2], ref11) < TOL
        for a, b in zip(out["i8"], out["dmma"]):
            assert rel_err(a, b) < 1e-11
        # the int8 engine's FP64 epilogue (AQML_I8_EPILOGUE=fp64: table exp, squaring chain, FP64 sums) against its default
        # integer epilogue (i8exp.cuh) on the same accumulators
        os.environ["AQML_I8_EPILOGUE"] = "fp64"
        try:
            f64 = (fused.local_gaussian(test, train, sig, zmax).cpu().numpy(),
                   fused.local_gaussian(train, train, sig, zmax, symmetric=True).cpu().numpy())
        finally:
            del os.environ["AQML_I8_EPILOGUE"]
        assert kernel_rel_err(f64[0], ref) < TOL and kernel_rel_err(f64[1], ref11) < TOL
        assert kernel_rel_err(out["i8"][0], f64[0]) < 2e-11 and kernel_rel_err(out["i8"][2], f64[1]) < 2e-11


def test_chunked_host_delivery_equals_one_shot():
    """fused.local_gaussian_to_host (row blocks, D2H overlapped on a second stream) == one contraction + one copy."""
    import torch
    nas, zs, coords = synth.qm9_like(83, 77)
    n1 = 50
    off = np.concatenate(([0], np.cumsum(nas)))
    A1 = off[n1]
    mb = o.get_slatm_mbtypes(synth.split_zs(nas, zs))
    kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
    train = fused.PackedRep(coords[:A1], zs[:A1], nas[:n1], mb, **kw)
    test = fused.PackedRep(coords[A1:], zs[A1:], nas[n1:], mb, **kw)
    sig = np.array([[c * 12.0] * 9 for c in (0.5, 1.0, 2.0)])
    ref = fused.local_gaussian(test, train, sig, 9).cpu()
    for chunks in (1, 4, 5, 200):
        out = torch.full((3, 83 - n1, n1), -1.0, dtype=torch.float64).pin_memory()
        fused.local_gaussian_to_host(coords[A1:], zs[A1:], nas[n1:], train, mb, sig, 9, out, chunks=chunks, **kw)
        torch.cuda.synchronize()
        assert rel_err(out.numpy(), ref.numpy()) < 1e-12


def test_eight_sigmas_on_the_int8_engine():
    """The int8 engine's limit of 8 sigmas per launch, as a power-of-two ladder (squaring chain) and as arbitrary widths."""
    nas, zs, coords = synth.qm9_like(40, 91)
    n1 = 28
    off = np.concatenate(([0], np.cumsum(nas)))
    A1 = off[n1]
    mb = o.get_slatm_mbtypes(synth.split_zs(nas, zs))
    kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
    xl = co.generate_slatm_batch(coords, zs, nas, mb, True, **kw)
    dso = co.get_kernel_width(nas, zs, xl)
    dso = np.concatenate((dso, np.full(9 - len(dso), 1e-9)))
    train = fused.PackedRep(coords[:A1], zs[:A1], nas[:n1], mb, **kw)
    test = fused.PackedRep(coords[A1:], zs[A1:], nas[n1:], mb, **kw)
    for coeffs in ([2.0 ** k for k in range(-3, 5)], [0.11, 0.37, 0.5, 0.93, 1.4, 2.2, 3.1, 5.0]):
        sig = np.array([c * dso / np.sqrt(2 * np.log(2)) for c in coeffs])
        ref = co.calc_lk_gaussian(xl[A1:], xl[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], 9, False, sig)
        K = fused.local_gaussian(test, train, sig, 9).cpu().numpy()
        assert K.shape == ref.shape and kernel_rel_err(K, ref) < TOL


def test_digit_planes_written_by_the_aslatm_kernel(monkeypatch):
    """AQML_FUSE_PLANES=1: the aSLATM kernel writes the int8 digit planes of a row itself (phase 4) instead of the two split
    kernels: same digits from the same FP64 rows, so K agrees to the order of the REDs."""
    import torch
    nas, zs, coords = synth.qm9_like(60, 77)
    mb = rcore.get_slatm_mbtypes(synth.split_zs(nas, zs))
    kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
    out = []
    for fuse in ("0", "1"):
        monkeypatch.setenv("AQML_FUSE_PLANES", fuse)
        r = fused.PackedRep(coords, zs, nas, mb, **kw)
        dso = fused.kernel_width(r, r, 9)
        sig = np.array([c * dso / np.sqrt(2 * np.log(2)) for c in (0.5, 1.0, 2.0)])
        out.append(fused.local_gaussian(r, r, sig, 9).cpu())
        r.close()
    assert torch.allclose(out[0], out[1], rtol=1e-13, atol=0)
