"""GPU parity at BASELINE.json's full size (config 2: 20k train x 2k test, local aSLATM Gaussian kernel,
4 sigmas) through checks that do not need the oracle on the whole problem: a random sample of kernel
elements recomputed by the CPU oracle, row-block sharding consistency, and rotation invariance."""
import numpy as np
import pytest

from conftest import kernel_rel_err, rel_err
from aqml_b200 import fused, synth
from aqml_b200.representation import core as rcore
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
KW = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
ZMAX = 9


def _mols(nas, zs, coords, idx):
    off = np.concatenate(([0], np.cumsum(nas)))
    sel = np.concatenate([np.arange(off[i], off[i + 1]) for i in idx])
    return nas[idx], zs[sel], coords[sel]


def test_config2_sampled_elements_match_oracle():
    import torch
    n_train, n_test = 20000, 2000
    train = synth.qm9_like_fast(n_train, 20251021, pool=512)
    test = synth.qm9_like_fast(n_test, 20251021 + 1000, pool=256)
    zl = synth.split_zs(train[0][:1500], train[1][:int(train[0][:1500].sum())])
    zl.append(np.array([1, 1, 1, 6, 6, 6, 7, 7, 7, 8, 8, 8, 9, 9, 9]))
    mb = rcore.get_slatm_mbtypes(zl)
    r1 = fused.PackedRep(torch.as_tensor(train[2], device="cuda"), train[1], train[0], mb, **KW)
    r2 = fused.PackedRep(torch.as_tensor(test[2], device="cuda"), test[1], test[0], mb, **KW)
    dso = fused.kernel_width(fused.PackedRep(train[2][:int(train[0][:300].sum())], train[1][:int(train[0][:300].sum())],
                                             train[0][:300], mb, **KW), r2, ZMAX)
    sig = np.array([c * dso / np.sqrt(2 * np.log(2)) for c in (0.5, 1.0, 2.0, 4.0)])
    K = fused.local_gaussian(r2, r1, sig, ZMAX)
    assert K.shape == (4, n_test, n_train)
    # (1) 48 sampled (test, train) molecule pairs against the CPU oracle, all four sigmas
    rng = np.random.default_rng(5)
    ia = rng.choice(n_test, 8, replace=False)
    ib = rng.choice(n_train, 6, replace=False)
    na_, za_, ca_ = _mols(*test, ia)
    nb_, zb_, cb_ = _mols(*train, ib)
    xa = co.generate_slatm_batch(ca_, za_, na_, mb, True, **KW)
    xb = co.generate_slatm_batch(cb_, zb_, nb_, mb, True, **KW)
    ref = co.calc_lk_gaussian(xa, xb, za_, zb_, na_, nb_, ZMAX, False, sig)
    got = K[:, torch.as_tensor(ia, device="cuda")][:, :, torch.as_tensor(ib, device="cuda")].cpu().numpy()
    assert kernel_rel_err(got, ref) < 1e-10
    # (2) row-block sharding: rows [700, 1200) computed on their own == the same rows of the full K
    Kb = fused.local_gaussian(r2, r1, sig, ZMAX, row0=700, nrows=500)
    assert torch.allclose(Kb, K[:, 700:1200], rtol=1e-13, atol=0)
    # (3) K is a sum of positive terms bounded by the number of same-element atom pairs
    cnt1 = np.stack([np.bincount(np.repeat(np.arange(n_train), train[0]), weights=(train[1] == z), minlength=n_train)
                     for z in (1, 6, 7, 8, 9)])
    cnt2 = np.stack([np.bincount(np.repeat(np.arange(n_test), test[0]), weights=(test[1] == z), minlength=n_test)
                     for z in (1, 6, 7, 8, 9)])
    bound = torch.as_tensor(cnt2.T @ cnt1, device="cuda")
    assert bool((K >= 0).all()) and bool((K <= bound[None] * (1 + 1e-12)).all())
    assert bool((K[3] >= K[0]).all())  # wider sigma, larger kernel
    r1.close()
    r2.close()


def test_rotation_invariance_at_scale():
    """aSLATM depends on distances and angles only: rigidly rotated / translated molecules give the same rows."""
    import torch
    nas, zs, coords = synth.qm9_like_fast(3000, 77, pool=256)
    rng = np.random.default_rng(1)
    q, r = np.linalg.qr(rng.normal(size=(len(nas), 3, 3)))
    q = q * np.sign(np.diagonal(r, axis1=1, axis2=2))[:, None, :]
    mol = np.repeat(np.arange(len(nas)), nas)
    rot = np.einsum('aij,aj->ai', q[mol], coords) + rng.normal(size=(len(nas), 3))[mol]
    zl = synth.split_zs(nas[:800], zs[:int(nas[:800].sum())])
    zl.append(np.array([1, 1, 1, 6, 6, 6, 7, 7, 7, 8, 8, 8, 9, 9, 9]))
    mb = rcore.get_slatm_mbtypes(zl)
    kw = dict(sigmas=[.05, .05], dgrids=[.04, .04], rcut=4.8)
    x0 = rcore.generate_slatm_batch(torch.as_tensor(coords, device="cuda"), zs, nas, mb, local=True, **kw)
    x1 = rcore.generate_slatm_batch(torch.as_tensor(rot, device="cuda"), zs, nas, mb, local=True, **kw)
    scale = float(x0.abs().max())
    # rotation changes coordinates by O(1e-16) relative; acos near +-1 amplifies to ~1e-8 rad for collinear triples
    assert float((x0 - x1).abs().max()) < 1e-6 * scale
    med = float(((x0 - x1).abs().max(dim=1).values).median())
    assert med < 1e-11 * scale
    g0 = rcore.generate_slatm_batch(torch.as_tensor(coords, device="cuda"), zs, nas, mb, local=False, **kw)
    off = torch.as_tensor(np.concatenate(([0], np.cumsum(nas))), device="cuda")
    summed = torch.zeros_like(g0)
    summed.index_add_(0, torch.as_tensor(mol, device="cuda"), x0)
    n1 = sum(1 for t in mb if len(t) == 1)
    assert float((summed[:, n1:] - g0[:, n1:]).abs().max()) < 1e-11 * scale  # SLATM == sum of aSLATM rows


def test_symmetric_multi_tile_both_engines():
    """K(X, X) at a size where every element group spans many 128 x 64 tiles (diagonal-crossing tiles, ragged last tiles):
    the symmetric route of the int8 engine == its rectangular route == the FP64 engine."""
    import os
    from aqml_b200 import fused
    from aqml_b200.representation import core as rcore
    nas, zs, coords = synth.qm9_like_fast(700, 9, pool=256)
    mb = rcore.get_slatm_mbtypes(synth.split_zs(nas, zs))
    kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
    rep = fused.PackedRep(coords, zs, nas, mb, **kw)
    dso = fused.kernel_width(rep, rep, 9)
    sig = np.array([c * dso / np.sqrt(2 * np.log(2)) for c in (0.5, 1.0, 2.0, 4.0)])
    Ks = fused.local_gaussian(rep, rep, sig, 9, symmetric=True)
    Kr = fused.local_gaussian(rep, rep, sig, 9)
    assert float((Ks - Kr).abs().max() / Kr.abs().max()) < 1e-12
    assert bool((Ks == Ks.transpose(1, 2)).all())
    os.environ["AQML_PAIR_ENGINE"] = "dmma"
    try:
        Kd = fused.local_gaussian(rep, rep, sig, 9, symmetric=True)
    finally:
        del os.environ["AQML_PAIR_ENGINE"]
    rel = ((Ks - Kd).abs() / Kd.abs().clamp_min(1e-10 * float(Kd.abs().max()))).max()
    assert float(rel) < 1e-11
