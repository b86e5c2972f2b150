#!/usr/bin/env python
"""Config-5-shaped run (BASELINE.json configs[4]) on ONE B200: multi-fidelity KRR with 50k low-fidelity + 5k
high-fidelity training molecules and 2k queries.  One aSLATM kernel matrix serves both levels (nested prefixes,
rkrr.py:203-228): K_train (55k x 55k, symmetric route) + K_test (2k x 55k), then two GPU solves.
Usage: python tools/bench_config5.py [n_low n_high n_test]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aqml_b200 import fused, synth  # noqa: E402
from aqml_b200.representation import core as rcore  # noqa: E402

n_low, n_high, n_test = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (50000, 5000, 2000)
n_train = n_low + n_high  # high-fidelity molecules first (prefix of the low-fidelity set, which uses all n_train)
dev = torch.device("cuda")
train = synth.qm9_like_fast(n_train, 20251023, pool=512)
test = synth.qm9_like_fast(n_test, 20251024, pool=256)
zl = synth.split_zs(train[0][:1500], train[1][:int(train[0][:1500].sum())])
zl.append(np.array([1, 1, 1, 6, 6, 6, 7, 7, 7, 8, 8, 8, 9, 9, 9]))
mb = rcore.get_slatm_mbtypes(zl)
kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
zmax = 9
els = np.array([1, 6, 7, 8, 9])


def counts(s):
    mol = np.repeat(np.arange(len(s[0])), s[0])
    return np.stack([np.bincount(mol, weights=(s[1] == z), minlength=len(s[0])) for z in els], axis=1)


def ev():
    return torch.cuda.Event(enable_timing=True)


t0 = time.time()
e = [ev() for _ in range(5)]
e[0].record()
r1 = fused.PackedRep(torch.as_tensor(train[2], device=dev), train[1], train[0], mb, **kw)
r2 = fused.PackedRep(torch.as_tensor(test[2], device=dev), test[1], test[0], mb, **kw)
e[1].record()
sub = fused.PackedRep(train[2][:int(train[0][:300].sum())], train[1][:int(train[0][:300].sum())], train[0][:300], mb, **kw)
dso = fused.kernel_width(sub, sub, zmax)
sig = np.array([1.0 * dso / np.sqrt(2 * np.log(2))])          # one sigma: K_train is 24 GB at 55k
K1 = fused.local_gaussian(r1, r1, sig, zmax, symmetric=True)[0]
e[2].record()
K2 = fused.local_gaussian(r2, r1, sig, zmax)[0]
e[3].record()
# synthetic two-level labels: dressed atoms + a smooth function of the composition; level 1 = level 0 + correction
rng = np.random.default_rng(0)
n1c, n2c = counts(train), counts(test)
w0, w1 = rng.normal(size=5) * 50, rng.normal(size=5)
f = lambda c: np.sin(c @ np.array([.11, .07, .23, .19, .31]))  # noqa: E731
y0_tr, y0_te = n1c @ w0 + 3 * f(n1c), n2c @ w0 + 3 * f(n2c)
y1_tr, y1_te = y0_tr + n1c @ w1 + 0.3 * f(2 * n1c), y0_te + n2c @ w1 + 0.3 * f(2 * n2c)
lam = 1e-6
pred = torch.zeros(n_test, dtype=torch.float64, device=dev)
res = []
for level, (n, tgt) in enumerate(((n_train, y0_tr), (n_high, y1_tr - y0_tr))):
    esb = np.linalg.lstsq(n1c[:n], tgt[:n], rcond=None)[0]
    y = torch.as_tensor(tgt[:n] - n1c[:n] @ esb, device=dev)
    A = K1[:n, :n].clone()
    A.diagonal().add_(lam)
    alpha = torch.linalg.solve(A, y)
    res.append(float((A @ alpha - y).abs().max() / y.abs().max()))
    del A
    pred += K2[:, :n] @ alpha + torch.as_tensor(n2c @ esb, device=dev)
e[4].record()
torch.cuda.synchronize()
err_mf = float(np.mean(np.abs(pred.cpu().numpy() - y1_te)))
base = float(np.mean(np.abs(y1_te - np.mean(y1_tr))))
print(json.dumps({"workload": "config5-shaped multi-fidelity KRR: %d low + %d high fidelity, %d queries, 1 GPU, 1 sigma"
                              % (n_low, n_high, n_test),
                  "atoms": int(train[0].sum() + test[0].sum()),
                  "aslatm_ms": e[0].elapsed_time(e[1]), "K_train_symmetric_ms": e[1].elapsed_time(e[2]),
                  "K_test_ms": e[2].elapsed_time(e[3]), "two_gpu_solves_ms": e[3].elapsed_time(e[4]),
                  "solve_rel_residuals": res, "mae_multifidelity": err_mf, "mae_mean_predictor": base,
                  "K_train_GB": K1.numel() * 8 / 1e9, "wall_s": time.time() - t0}))
