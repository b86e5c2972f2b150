"""GPU: the reference workflow's consumers on top of the CUDA kernels."""
import numpy as np
import pytest

from conftest import kernel_rel_err, rel_err
from aqml_b200 import kernels, synth
from aqml_b200.algo import aqml as algo
from aqml_b200.algo import krr
from oracle import c_oracle as co
from oracle import np_oracle as o

pytestmark = pytest.mark.gpu


def _ref_calc_ka(x2, x1, zs2, zs1, nas1, sigmas, kernel):
    """aqml.py:258-309 restated with the oracle distances."""
    dist = co.l2_distance if kernel == 'g' else co.l1_distance
    ksa = np.zeros((len(sigmas), len(zs2), len(zs1)))
    for z in np.unique(zs2):
        i1, i2 = np.nonzero(zs1 == z)[0], np.nonzero(zs2 == z)[0]
        if len(i1) == 0:
            continue
        ds = dist(x2[i2], x1[i1])
        for k, s in enumerate(sigmas):
            ksa[k][np.ix_(i2, i1)] = np.exp(-0.5 * (ds / s[z - 1]) ** 2)
    off = np.concatenate(([0], np.cumsum(nas1)))
    return np.stack([ksa[:, :, off[i]:off[i + 1]].sum(axis=2) for i in range(len(nas1))], axis=2), ksa


def test_calc_ka_matches_restated_reference():
    nas, zs, coords = synth.qm9_like(14, 42)
    mb = o.get_slatm_mbtypes(synth.split_zs(nas, zs))
    x = co.generate_slatm_batch(coords, zs, nas, mb, True, sigmas=(.05, .05), dgrids=(.08, .08), rcut=4.8)
    n1 = 10
    A1 = int(nas[:n1].sum())
    zmax = 9
    dso = co.get_kernel_width(nas, zs, x)
    dso = np.concatenate((dso, np.full(zmax - len(dso), 1e-9)))
    sig = np.array([c * dso / np.sqrt(2 * np.log(2)) for c in (1.0, 2.0)])
    for kern in ('g', 'l'):
        s = sig if kern == 'g' else sig * 30
        ref_ks, ref_ksa = _ref_calc_ka(x[A1:], x[:A1], zs[A1:], zs[:A1], nas[:n1], s, kern)
        ks = algo.calc_ka(x[A1:], x[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], False, zmax, s, kernel=kern)
        assert ks.shape == ref_ks.shape
        assert kernel_rel_err(ks, ref_ks) < 1e-10
        kz = algo.calc_ka(x[A1:], x[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], False, zmax, s, zaq=6, kernel=kern)
        assert kernel_rel_err(kz, ref_ksa[:, zs[A1:] == 6][:, :, zs[:A1] == 6]) < 1e-10
    # the sum of the atomic contributions over test-molecule atoms is the local kernel (aqml.py:913)
    ks = algo.calc_ka(x[A1:], x[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], False, zmax, sig)
    kl = kernels.get_local_kernels_gaussian(x[A1:], x[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], False, zmax, sig)
    off2 = np.concatenate(([0], np.cumsum(nas[n1:])))
    summed = np.stack([ks[:, off2[i]:off2[i + 1]].sum(axis=1) for i in range(len(nas) - n1)], axis=1)
    assert kernel_rel_err(summed, kl) < 1e-10


def test_learning_curve_and_multifidelity(demo):
    n1 = int(demo["n1"])
    nas, zs = demo["nas"], demo["zs"]
    zs_list = synth.split_zs(nas, zs)
    zsu = np.unique(zs)
    ngs = np.array([[(z == zz).sum() for zz in zsu] for z in zs_list])
    nheav = np.array([(z > 1).sum() for z in zs_list[:n1]])
    ys = demo["energies"] * krr.H2KC
    free = [krr.H2KC * o.ORCA_B3LYPVDZ[int(z)] for z in zsu]
    for device in (False, True):
        rows = krr.learning_curve(demo["k1"][1], demo["k2"][1], ngs, ys, nheav, free, 1e-4, device=device)
        readme = demo["readme_curve"]
        for r, ref in zip(rows, readme):
            assert (r[0], r[1]) == (ref[0], ref[1]) and abs(r[2] - ref[2]) < 5.1e-5
    # multi-fidelity: a two-level problem whose high level differs from the low level by a smooth term
    rng = np.random.default_rng(0)
    k1, k2 = demo["k1"][1], demo["k2"][1]
    y0 = ys[:n1]
    y1 = y0 + 3.0 + 0.01 * ngs[:n1] @ rng.random(len(zsu))
    yl = np.stack([y0, y1], axis=1)
    got = krr.multi_fidelity_predict(k1, k2, yl, [n1, 10], ngs[:n1].astype(float), ngs[n1:].astype(float), 1e-6)
    # restated rkrr.py:203-228
    pred = np.zeros(1)
    for l, n in enumerate([n1, 10]):
        tgt = yl[:n, l] if l == 0 else yl[:n, l] - yl[:n, l - 1]
        esb = np.linalg.lstsq(ngs[:n].astype(float), tgt, rcond=None)[0]
        kk = k1[:n, :n].copy()
        kk[np.diag_indices_from(kk)] += 1e-6
        pred += k2[:, :n] @ np.linalg.solve(kk, tgt - ngs[:n] @ esb) + ngs[n1:] @ esb
    assert np.allclose(got, pred, rtol=1e-12, atol=1e-9)


def test_amons_workflow_per_query_sigmas():
    """every query is trained on its own amon subset (maps, -1 padded) with its own
    per-element dmax -> sigmas; compared with the restated reference iteration for each query."""
    from aqml_b200.algo import amons as wf
    rng = np.random.default_rng(9)
    gen = np.random.Generator(np.random.PCG64(101))
    heavy = ([6, 7, 8], [0.6, 0.2, 0.2])  # every element occurs many times: dmax > 0 for all of them

    def make(n, lo, hi):
        ms = [synth.make_molecule(gen, int(gen.integers(lo, hi + 1)), heavy=heavy) for _ in range(n)]
        return (np.array([len(z) for z, _ in ms], dtype=np.int32), np.concatenate([z for z, _ in ms]),
                np.concatenate([x for _, x in ms]))

    pool, queries = make(40, 2, 5), make(4, 7, 9)
    maps = np.full((4, 14), -1, dtype=np.int64)
    for q in range(4):
        k = int(rng.integers(6, 14))
        maps[q, :k] = rng.choice(40, k, replace=False)
    zl = synth.split_zs(pool[0], pool[1]) + synth.split_zs(queries[0], queries[1])
    mb = o.get_slatm_mbtypes(zl)
    kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
    res = wf.query_kernels(pool, queries, maps, coeffs=(0.5, 1.0), mbtypes=mb)
    zmax = int(max(pool[1].max(), queries[1].max()))
    for q, (idx, sig, ks1, ks2) in enumerate(res):
        assert idx == sorted(int(i) for i in maps[q][maps[q] > -1])
        a = wf._subset(*pool, idx)
        t = wf._subset(*queries, [q])
        x1 = co.generate_slatm_batch(a[2], a[1], a[0], mb, True, **kw)
        x2 = co.generate_slatm_batch(t[2], t[1], t[0], mb, True, **kw)
        dso = co.get_kernel_width(np.concatenate((a[0], t[0])), np.concatenate((a[1], t[1])), np.concatenate((x1, x2)))
        dso = np.concatenate((dso, np.full(zmax - len(dso), 1e-9)))
        ref_sig = np.array([o.width_factor() * dso * c for c in (0.5, 1.0)])
        assert np.allclose(sig, ref_sig, rtol=1e-10, atol=0)
        r1 = co.calc_lk_gaussian(x1, x1, a[1], a[1], a[0], a[0], zmax, False, ref_sig)
        r2 = co.calc_lk_gaussian(x2, x1, t[1], a[1], t[0], a[0], zmax, False, ref_sig)
        assert kernel_rel_err(ks1, r1) < 1e-10 and kernel_rel_err(ks2, r2) < 1e-10


def test_sharded_krr_single_rank_matches_direct_solve():
    """ShardedKRR (block lower-triangle assembly + block Cholesky, world 1) == full K + LU solve (aqml.py:874-876)."""
    import torch
    from aqml_b200 import fused
    from aqml_b200.algo.krr_sharded import ShardedKRR
    from aqml_b200.representation import core as rcore
    train = synth.qm9_like(150, 6, max_heavy=6)   # seed: every element has >= 2 atoms (a single atom gives d_max = 0)
    test = synth.qm9_like(21, 7, max_heavy=6)
    mb = rcore.get_slatm_mbtypes(synth.split_zs(train[0], train[1]) + synth.split_zs(test[0], test[1]))
    kw = dict(sigmas=(.05, .05), dgrids=(.08, .08), rcut=4.8)
    zmax = 9
    model = ShardedKRR(train[0], train[1], train[2], mb, zmax, nb=32, **kw)   # 5 blocks, ragged last one
    rep_train = fused.PackedRep(train[2], train[1], train[0], mb, **kw)
    rep_test = fused.PackedRep(test[2], test[1], test[0], mb, **kw)
    dso = model.kernel_width()
    assert kernel_rel_err(dso, fused.kernel_width(rep_train, rep_train, zmax)) < 1e-12
    sig = dso / np.sqrt(2 * np.log(2))
    K_loc = model.assemble(sig)
    K = fused.local_gaussian(rep_train, rep_train, sig[None], zmax)[0]
    low = torch.tril(torch.ones_like(K, dtype=torch.bool))
    assert kernel_rel_err(torch.where(low, K_loc, 0).cpu().numpy(), torch.where(low, K, 0).cpu().numpy()) < 1e-12
    y = torch.as_tensor(np.random.default_rng(0).normal(size=150), device=K.device)
    # lambda = 10 keeps cond(K + lambda I) ~ 1e4, so that rounding-order noise in K (FP64 atomics, ~1e-16) stays
    # far below the tolerance in alpha
    alpha = model.fit(K_loc, y, 10.0)
    Kh = K.cpu().numpy() + 10.0 * np.eye(150)
    a_ref = np.linalg.solve(Kh, y.cpu().numpy())
    assert rel_err(alpha.cpu().numpy(), a_ref) < 1e-9
    y_ref = fused.local_gaussian(rep_test, rep_train, sig[None], zmax)[0].cpu().numpy() @ a_ref
    assert rel_err(model.predict(rep_test).cpu().numpy(), y_ref) < 1e-9


def test_readme_atomic_energies_through_gpu(demo):
    """The reference's second KAT (demo/example/README.md:75-94, `-ieaq -prog orca`) with every kernel on the GPU:
    aSLATM -> widths -> K_train -> alphas (host LAPACK as upstream) -> calc_ka -> per-atom energies, two printed decimals."""
    import torch
    from aqml_b200.representation import core as rcore
    from conftest import mbtypes_from_array
    readme = [-163.80, -163.55, -163.57, -163.55, -163.77, -163.42, -159.70, -161.60, -90.76, -59.83, -59.98, -59.99,
              -59.98, -59.78, -64.12, -63.75, -60.70]
    n1 = int(demo["n1"])
    nas, zs, coords = demo["nas"], demo["zs"], demo["coords"]
    mb = mbtypes_from_array(demo["mbtypes"])
    off = np.concatenate(([0], np.cumsum(nas)))
    A1 = off[n1]
    x = rcore.generate_slatm_batch(torch.as_tensor(coords, device="cuda"), zs, nas, mb, local=True, sigmas=[.05, .05],
                                   dgrids=[.04, .04], rcut=4.8)
    dso = algo.get_kernel_width(nas, zs, x)
    sig = (dso / np.sqrt(2 * np.log(2)))[None, :]
    k1 = kernels.get_local_kernels_gaussian(x[:A1], x[:A1], zs[:A1], zs[:A1], nas[:n1], nas[:n1], False, 8, sig)[0].cpu().numpy()
    zs_list = synth.split_zs(nas, zs)
    zsu = np.unique(zs)
    ngs = np.array([[(z == zz).sum() for zz in zsu] for z in zs_list])
    ys = demo["energies"] * krr.H2KC
    free = np.array([krr.H2KC * o.ORCA_B3LYPVDZ[int(z)] for z in zsu])
    _, ys1, ys2, esb = krr.calc_ae_dressed(ngs, ys, np.arange(n1), np.arange(n1, n1 + 1), free)
    alphas = krr.solve(k1, ys1, 1e-4)
    ksa21 = algo.calc_ka(x[A1:], x[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], False, 8, sig)[0].cpu().numpy()
    easq = ksa21 @ alphas
    edct = dict(zip([int(z) for z in zsu], esb))
    fdct = dict(zip([int(z) for z in zsu], free))
    ea = np.array([easq[j] + edct[int(z)] - fdct[int(z)] for j, z in enumerate(zs[A1:])])
    assert np.all(np.abs(ea - np.array(readme)) < 5.1e-3), ea
