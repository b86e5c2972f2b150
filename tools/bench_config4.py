#!/usr/bin/env python
"""Config-4-shaped check (BASELINE.json configs[3]): amon conformers (<= 7 heavy atoms, 8 elements) x large query
molecules (100-200 atoms): fused aSLATM -> local Gaussian kernel, timing + sampled parity against the CPU oracle."""
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
from oracle import c_oracle as co  # noqa: E402

n_amons, n_query = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, int(sys.argv[2]) if len(sys.argv) > 2 else 500
rng = np.random.Generator(np.random.PCG64(4))
pool = []
for _ in range(600):
    z, x = synth.make_molecule(rng, int(rng.integers(1, 8)), heavy=synth.BIG_HEAVY, max_atoms=23)
    pool.append((z, x))
base = (np.array([len(z) for z, _ in pool], dtype=np.int32), np.concatenate([z for z, _ in pool]),
        np.concatenate([x for _, x in pool]))
amons = synth.replicate(base, n_amons, 11)
qpool = synth.large_like(40, 12)
query = synth.replicate(qpool, n_query, 13)
zl = synth.split_zs(base[0], base[1]) + synth.split_zs(qpool[0], qpool[1])
mb = rcore.get_slatm_mbtypes(zl)
kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
zmax = 17
dev = torch.device("cuda")
ca, cq = torch.as_tensor(amons[2], device=dev), torch.as_tensor(query[2], device=dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


def step():
    e = [ev() for _ in range(3)]
    e[0].record()
    ra = fused.PackedRep(ca, amons[1], amons[0], mb, **kw)
    rq = fused.PackedRep(cq, query[1], query[0], mb, **kw)
    e[1].record()
    K = fused.local_gaussian(rq, ra, sig, zmax)
    e[2].record()
    torch.cuda.synchronize()
    return ra, rq, K, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])


sub = fused.PackedRep(amons[2][:int(amons[0][:300].sum())], amons[1][:int(amons[0][:300].sum())], amons[0][:300], mb, **kw)
dso = fused.kernel_width(sub, sub, zmax)
sig = np.array([c * dso / np.sqrt(2 * np.log(2)) for c in (0.5, 1.0, 2.0, 4.0)])
for _ in range(2):
    ra, rq, K, t1, t2 = step()
    ra.close()
    rq.close()
ra, rq, K, t1, t2 = step()
# sampled parity
ia = rng.choice(n_query, 2, replace=False)
ib = rng.choice(n_amons, 5, replace=False)


def pick(s, idx):
    off = np.concatenate(([0], np.cumsum(s[0])))
    sel = np.concatenate([np.arange(off[i], off[i + 1]) for i in idx])
    return s[0][idx], s[1][sel], s[2][sel]


na_, za_, xa_ = pick(query, ia)
nb_, zb_, xb_ = pick(amons, ib)
xa = co.generate_slatm_batch(xa_, za_, na_, mb, True, **kw)
xb = co.generate_slatm_batch(xb_, zb_, nb_, mb, True, **kw)
ref = co.calc_lk_gaussian(xa, xb, za_, zb_, na_, nb_, zmax, False, sig)
got = K[:, torch.as_tensor(ia, device=dev)][:, :, torch.as_tensor(ib, device=dev)].cpu().numpy()
err = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-10 * np.abs(ref).max())))
D = sum(1 if len(t) == 1 else (118 if len(t) == 2 else 96) for t in mb)
print(json.dumps({"workload": "config4-shaped: %d amons x %d queries, 8 elements" % (n_amons, n_query), "D": D,
                  "mbtypes": len(mb), "atoms": [int(amons[0].sum()), int(query[0].sum())],
                  "aslatm_ms": t1, "kernel_ms": t2, "atoms_per_s": (amons[0].sum() + query[0].sum()) / t1 * 1e3,
                  "kernel_elements_per_s": 4.0 * n_amons * n_query / t2 * 1e3, "sampled_max_rel_err": err,
                  "rep_GB": (ra.nbytes + rq.nbytes) / 1e9}))
