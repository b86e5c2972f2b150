"""GPU parity on the BASELINE.json configurations that round 1 left without a CUDA-path check ("next" item 1):
config 1 end to end (global SLATM -> Gaussian KRR predictions), config 3 at its per-GPU shape (Gaussian AND Laplacian,
sampled elements), config 4 (amons x large queries, sampled molecule pairs) and config 5 (multi-fidelity KRR on kernels
from the GPU, against the restated rkrr.py:203-228 on oracle kernels)."""
import numpy as np
import pytest

from conftest import kernel_rel_err
from aqml_b200 import fused, kernels, synth
from aqml_b200.algo import krr
from aqml_b200.representation import core as rcore
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
KW = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)


def _mols(nas, zs, coords, idx):
    off = np.concatenate(([0], np.cumsum(nas)))
    sel = np.concatenate([np.arange(off[i], off[i + 1]) for i in idx])
    return nas[idx], zs[sel], coords[sel]


def _labels(nas, zs, coords, seed):
    """Synthetic energies (Ha): dressed atoms + a smooth function of the geometry."""
    rng = np.random.default_rng(seed)
    e_z = {1: -0.5, 6: -37.8, 7: -54.5, 8: -75.0, 9: -99.7, 15: -341.2, 16: -398.0, 17: -460.1}
    off = np.concatenate(([0], np.cumsum(nas)))
    w = rng.random(3)
    ys = np.zeros(len(nas))
    for i in range(len(nas)):
        z, x = zs[off[i]:off[i + 1]], coords[off[i]:off[i + 1]]
        d = np.linalg.norm(x[:, None] - x[None], axis=2)[np.triu_indices(len(z), 1)]
        ys[i] = sum(e_z[int(v)] for v in z) - 0.02 * np.exp(-d * w[0]).sum() + 0.003 * np.cos(d * (1 + w[1])).sum()
    return ys


def _krr_predict(k1, k2, ngs1, ngs2, y1, llambda):
    """aqml.py:129-181 (dressed-atom baseline by least squares) + aqml.py:874-876."""
    esb = np.linalg.lstsq(ngs1, y1, rcond=None)[0]
    kk = k1.copy()
    kk[np.diag_indices_from(kk)] += llambda
    return k2 @ np.linalg.solve(kk, y1 - ngs1 @ esb) + ngs2 @ esb


def test_config1_global_slatm_krr_predictions():
    """BASELINE configs[0]: 1000 QM9-shaped molecules, global SLATM at the default dgrid 0.03, Gaussian kernel, 800 / 200
    split: predictions from (GPU SLATM, GPU kernel) vs the oracle pipeline, <= 1e-8 Ha."""
    nas, zs, coords = synth.qm9_like_fast(1000, 20251019 + 1, pool=1000)
    mb = rcore.get_slatm_mbtypes(synth.split_zs(nas, zs))
    kw = dict(sigmas=(.05, .05), dgrids=(.03, .03), rcut=4.8)
    x_ref = co.generate_slatm_batch(coords, zs, nas, mb, False, **kw)
    x = rcore.generate_slatm_batch(coords, zs, nas, mb, local=False, sigmas=[.05, .05], dgrids=[.03, .03], rcut=4.8)
    assert x.shape == x_ref.shape and x.shape[1] == sum(1 if len(t) == 1 else (157 if len(t) == 2 else 128) for t in mb)
    scale = np.abs(x_ref).max()
    assert np.max(np.abs(x - x_ref) / np.maximum(np.abs(x_ref), 1e-10 * scale)) < 1e-10
    n1 = 800
    dmax = co.l2_distance(x_ref[:n1], x_ref[:n1]).max()
    sigma = dmax / np.sqrt(2 * np.log(2))                     # coeff 1 (aqml.py:195,775)
    k1_ref, k2_ref = co.gaussian_kernel(x_ref[:n1], x_ref[:n1], sigma), co.gaussian_kernel(x_ref[n1:], x_ref[:n1], sigma)
    k1, k2 = kernels.gaussian_kernel(x[:n1], x[:n1], sigma), kernels.gaussian_kernel(x[n1:], x[:n1], sigma)
    assert kernel_rel_err(k1, k1_ref) < 1e-10 and kernel_rel_err(k2, k2_ref) < 1e-10
    zsu = np.unique(zs)
    ngs = np.array([[(z == zz).sum() for zz in zsu] for z in synth.split_zs(nas, zs)], dtype=float)
    ys = _labels(nas, zs, coords, 3)
    p_ref = _krr_predict(np.asarray(k1_ref), np.asarray(k2_ref), ngs[:n1], ngs[n1:], ys[:n1], 1e-4)
    p_gpu = _krr_predict(np.asarray(k1), np.asarray(k2), ngs[:n1], ngs[n1:], ys[:n1], 1e-4)
    assert np.max(np.abs(p_gpu - p_ref)) < 1e-8, np.max(np.abs(p_gpu - p_ref))
    assert np.mean(np.abs(p_ref - ys[n1:])) < np.mean(np.abs(ys[n1:] - ngs[n1:] @ np.linalg.lstsq(ngs[:n1], ys[:n1], rcond=None)[0]))


@pytest.mark.parametrize("kernel", ["gaussian", "laplacian"])
def test_config3_per_gpu_shape_sampled(kernel):
    """BASELINE configs[2] at the shape one of 8 GPUs holds: 12 500 rows x 100 000 columns, global SLATM (D = 8975),
    4 sigmas: 64 sampled (row, column) pairs per sigma against the oracle's kernel on oracle SLATM rows."""
    import torch
    n_rows, n_cols = 12500, 100000
    cols = synth.qm9_like_fast(n_cols, 20251019 + 3, pool=512)
    rows = synth.qm9_like_fast(n_rows, 20251019 + 3 + 1000, pool=256)
    zl = synth.split_zs(cols[0][:1500], cols[1][:int(cols[0][:1500].sum())])
    zl.append(np.array([1, 1, 1, 6, 6, 6, 7, 7, 7, 8, 8, 8, 9, 9, 9]))
    mb = rcore.get_slatm_mbtypes(zl)
    kwl = dict(sigmas=[.05, .05], dgrids=[.04, .04], rcut=4.8)
    xc = rcore.generate_slatm_batch(torch.as_tensor(cols[2], device="cuda"), cols[1], cols[0], mb, local=False, **kwl)
    xr = rcore.generate_slatm_batch(torch.as_tensor(rows[2], device="cuda"), rows[1], rows[0], mb, local=False, **kwl)
    assert xc.shape == (n_cols, 8975)
    rng = np.random.default_rng(17)
    ir, ic = rng.choice(n_rows, 64, replace=False), rng.choice(n_cols, 64, replace=False)
    mr, mc = _mols(*rows, ir), _mols(*cols, ic)
    xr_ref = co.generate_slatm_batch(mr[2], mr[1], mr[0], mb, False, **KW)
    xc_ref = co.generate_slatm_batch(mc[2], mc[1], mc[0], mb, False, **KW)
    if kernel == "gaussian":
        dmax = co.l2_distance(xc_ref, xc_ref).max()
        sig = np.array([c * dmax / np.sqrt(2 * np.log(2)) for c in (0.5, 1.0, 2.0, 4.0)])
        ref = np.stack([np.exp(-0.5 * ((xr_ref - xc_ref) ** 2).sum(axis=1) / s ** 2) for s in sig])      # fkernels.f90:344-355
    else:
        dmax = co.l1_distance(xc_ref, xc_ref).max()
        sig = np.array([c * dmax / np.log(2) for c in (0.5, 1.0, 2.0, 4.0)])
        ref = np.stack([np.exp(-np.abs(xr_ref - xc_ref).sum(axis=1) / s) for s in sig])                   # fkernels.f90:377-385
    K = kernels.global_kernels(xr, xc, sig, kernel)
    assert K.shape == (4, n_rows, n_cols)
    got = K[:, torch.as_tensor(ir, device="cuda"), torch.as_tensor(ic, device="cuda")].cpu().numpy()
    assert kernel_rel_err(got, ref) < 1e-10, kernel_rel_err(got, ref)


def test_config4_amons_times_large_queries_sampled():
    """BASELINE configs[3]: 10 000 amon conformers x 500 query molecules of 100-200 atoms, 8 elements (D = 31 904), fused
    aSLATM -> local Gaussian kernel; 3 x 8 sampled molecule pairs, 4 sigmas, against the oracle."""
    import torch
    n_amons, n_query = 10000, 500
    rng = np.random.Generator(np.random.PCG64(4))
    pool = [synth.make_molecule(rng, int(rng.integers(1, 8)), heavy=synth.BIG_HEAVY, max_atoms=23) for _ in range(600)]
    base = (np.array([len(z) for z, _ in pool], dtype=np.int32), np.concatenate([z for z, _ in pool]),
            np.concatenate([x for _, x in pool]))
    amons = synth.replicate(base, n_amons, 11)
    qpool = synth.large_like(40, 12)
    query = synth.replicate(qpool, n_query, 13)
    mb = rcore.get_slatm_mbtypes(synth.split_zs(base[0], base[1]) + synth.split_zs(qpool[0], qpool[1]))
    assert sum(1 if len(t) == 1 else (118 if len(t) == 2 else 96) for t in mb) == 31904
    zmax = 17
    ra = fused.PackedRep(torch.as_tensor(amons[2], device="cuda"), amons[1], amons[0], mb, **KW)
    rq = fused.PackedRep(torch.as_tensor(query[2], device="cuda"), query[1], query[0], mb, **KW)
    dso = fused.kernel_width(ra, ra, zmax)
    sig = np.array([c * dso / np.sqrt(2 * np.log(2)) for c in (0.5, 1.0, 2.0, 4.0)])
    K = fused.local_gaussian(rq, ra, sig, zmax)
    assert K.shape == (4, n_query, n_amons)
    ia, ib = rng.choice(n_query, 3, replace=False), rng.choice(n_amons, 8, replace=False)
    na_, za_, ca_ = _mols(*query, ia)
    nb_, zb_, cb_ = _mols(*amons, ib)
    xa = co.generate_slatm_batch(ca_, za_, na_, mb, True, **KW)
    xb = co.generate_slatm_batch(cb_, zb_, nb_, mb, True, **KW)
    ref = co.calc_lk_gaussian(xa, xb, za_, zb_, na_, nb_, zmax, False, sig)
    got = K[:, torch.as_tensor(ia, device="cuda")][:, :, torch.as_tensor(ib, device="cuda")].cpu().numpy()
    assert kernel_rel_err(got, ref) < 1e-10, kernel_rel_err(got, ref)
    ra.close()
    rq.close()


def _multi_fidelity_ref(k1, k2, yl, n1s, ngs1, ngs2, llambda):
    """rkrr.py:203-228 restated: level l is trained on y_l - y_(l-1) over the nested prefix n1s[l] of the same kernel."""
    pred = np.zeros(k2.shape[0])
    for l, n in enumerate(n1s):
        tgt = yl[:n, l] if l == 0 else yl[:n, l] - yl[:n, l - 1]
        esb = np.linalg.lstsq(ngs1[:n], tgt, rcond=None)[0]
        kk = k1[:n, :n].copy()
        kk[np.diag_indices_from(kk)] += llambda
        pred += k2[:, :n] @ np.linalg.solve(kk, tgt - ngs1[:n] @ esb) + ngs2 @ esb
    return pred


def test_config5_multi_fidelity_on_gpu_kernels():
    """BASELINE configs[4], scaled to what the oracle can check: (a) 240 low- + 60 high-fidelity nested training molecules,
    40 queries: every kernel from the GPU (symmetric route + K_test), multi_fidelity_predict vs the restated rkrr.py on the
    ORACLE's kernels, <= 1e-8 Ha; (b) 2000 + 200 nested training molecules: GPU kernels, 64 sampled elements of K_train and
    K_test against the oracle, and the GPU-kernel predictions against the restated reference on the same kernels."""
    import torch
    zmax = 9
    nas, zs, coords = synth.qm9_like_fast(2400, 20251019 + 5, pool=512)
    mb = rcore.get_slatm_mbtypes(synth.split_zs(nas[:600], zs[:int(nas[:600].sum())]) + [np.array([1, 1, 1, 6, 6, 6, 7, 7, 7, 8, 8, 8, 9, 9, 9])])
    zsu = np.array([1, 6, 7, 8, 9])
    ngs = np.array([[(z == zz).sum() for zz in zsu] for z in synth.split_zs(nas, zs)], dtype=float)
    y_hi = _labels(nas, zs, coords, 11)
    y_lo = y_hi - 0.004 * ngs @ np.array([1.0, 3.0, 2.5, 2.0, 1.5]) - 0.001 * np.sin(np.arange(len(nas)))
    yl = np.stack([y_lo, y_hi], axis=1)

    # (a) small: full oracle pipeline
    n1, n2 = 240, 40
    tr, te = _mols(nas, zs, coords, np.arange(n1)), _mols(nas, zs, coords, np.arange(n1, n1 + n2))
    x1 = co.generate_slatm_batch(tr[2], tr[1], tr[0], mb, True, **KW)
    x2 = co.generate_slatm_batch(te[2], te[1], te[0], mb, True, **KW)
    dso = co.get_kernel_width(np.concatenate((tr[0], te[0])), np.concatenate((tr[1], te[1])), np.concatenate((x1, x2)))
    dso = np.concatenate((dso, np.full(zmax - len(dso), 1e-9)))
    sig = (dso / np.sqrt(2 * np.log(2)))[None]
    k1_ref = co.calc_lk_gaussian(x1, x1, tr[1], tr[1], tr[0], tr[0], zmax, False, sig)[0]
    k2_ref = co.calc_lk_gaussian(x2, x1, te[1], tr[1], te[0], tr[0], zmax, False, sig)[0]
    r1 = fused.PackedRep(tr[2], tr[1], tr[0], mb, **KW)
    r2 = fused.PackedRep(te[2], te[1], te[0], mb, **KW)
    dso_gpu = fused.kernel_width([r1, r2], [r1, r2], zmax)
    assert np.allclose(dso_gpu, dso, rtol=1e-10, atol=0)
    k1 = fused.local_gaussian(r1, r1, sig, zmax, symmetric=True)[0].cpu().numpy()
    k2 = fused.local_gaussian(r2, r1, sig, zmax)[0].cpu().numpy()
    assert kernel_rel_err(k1, k1_ref) < 1e-10 and kernel_rel_err(k2, k2_ref) < 1e-10
    n1s = [n1, 60]
    ref = _multi_fidelity_ref(k1_ref, k2_ref, yl[:n1], n1s, ngs[:n1], ngs[n1:n1 + n2], 1e-4)
    got = krr.multi_fidelity_predict(k1, k2, yl[:n1], n1s, ngs[:n1], ngs[n1:n1 + n2], 1e-4)
    assert np.max(np.abs(got - ref)) < 1e-8, np.max(np.abs(got - ref))
    r1.close()
    r2.close()

    # (b) 2000 low + 200 high fidelity, 200 queries: kernels on the GPU, sampled against the oracle
    n1, n2 = 2000, 200
    tr, te = _mols(nas, zs, coords, np.arange(n1)), _mols(nas, zs, coords, np.arange(n1 + 200, n1 + 200 + n2))
    r1 = fused.PackedRep(torch.as_tensor(tr[2], device="cuda"), tr[1], tr[0], mb, **KW)
    r2 = fused.PackedRep(torch.as_tensor(te[2], device="cuda"), te[1], te[0], mb, **KW)
    dso = fused.kernel_width([r1, r2], [r1, r2], zmax)
    sig = (dso / np.sqrt(2 * np.log(2)))[None]
    K1 = fused.local_gaussian(r1, r1, sig, zmax, symmetric=True)[0]
    K2 = fused.local_gaussian(r2, r1, sig, zmax)[0]
    assert torch.equal(K1, K1.T)
    rng = np.random.default_rng(23)
    ia, ib = rng.choice(n1, 8, replace=False), rng.choice(n1, 8, replace=False)
    ma, mbm = _mols(*tr, ia), _mols(*tr, ib)
    xa = co.generate_slatm_batch(ma[2], ma[1], ma[0], mb, True, **KW)
    xb = co.generate_slatm_batch(mbm[2], mbm[1], mbm[0], mb, True, **KW)
    ref11 = co.calc_lk_gaussian(xa, xb, ma[1], mbm[1], ma[0], mbm[0], zmax, False, sig)[0]
    assert kernel_rel_err(K1[torch.as_tensor(ia, device="cuda")][:, torch.as_tensor(ib, device="cuda")].cpu().numpy(), ref11) < 1e-10
    it = rng.choice(n2, 8, replace=False)
    mt = _mols(*te, it)
    xt = co.generate_slatm_batch(mt[2], mt[1], mt[0], mb, True, **KW)
    ref21 = co.calc_lk_gaussian(xt, xb, mt[1], mbm[1], mt[0], mbm[0], zmax, False, sig)[0]
    assert kernel_rel_err(K2[torch.as_tensor(it, device="cuda")][:, torch.as_tensor(ib, device="cuda")].cpu().numpy(), ref21) < 1e-10
    k1, k2 = K1.cpu().numpy(), K2.cpu().numpy()
    idx = np.concatenate((np.arange(n1), np.arange(n1 + 200, n1 + 200 + n2)))
    got = krr.multi_fidelity_predict(k1, k2, yl[:n1], [n1, 200], ngs[:n1], ngs[idx[n1:]], 1e-4)
    ref = _multi_fidelity_ref(k1, k2, yl[:n1], [n1, 200], ngs[:n1], ngs[idx[n1:]], 1e-4)
    assert np.max(np.abs(got - ref)) < 1e-8
    assert np.mean(np.abs(got - y_hi[idx[n1:]])) < 0.5 * np.mean(np.abs(y_lo[idx[n1:]] - y_hi[idx[n1:]]) + 1e-3) + 0.05
    r1.close()
    r2.close()
