#!/usr/bin/env python
"""The sharded KRR feed (aqml_b200/sharded.py) on N GPUs against the same problem done directly on one GPU:
kernel widths (distributed pruned max-distance pass), every K(train,train) tile a rank owns, its K(test,train)
columns and the all-reduced prediction.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aqml_b200 import fused, sharded, synth  # noqa: E402
from aqml_b200.representation import core as rcore  # noqa: E402

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
n_train, n_test, zmax = int(os.environ.get("N_TRAIN", 1500)), 160, 9
train = synth.qm9_like_fast(n_train, 11, pool=256)
test = synth.qm9_like_fast(n_test, 12, pool=64)
mb = rcore.get_slatm_mbtypes(synth.split_zs(train[0], train[1]) + synth.split_zs(test[0], test[1]))
kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
coeffs = (0.5, 1.0, 2.0, 4.0)
alpha = torch.as_tensor(np.random.default_rng(0).normal(size=(len(coeffs), n_train)), device=dev)

sk = sharded.ShardedKernels((train[0], train[1]), (test[0], test[1]), mb, zmax, rank, world, **kw)
c1, c2 = torch.as_tensor(train[2], device=dev), torch.as_tensor(test[2], device=dev)
for it in range(2):   # twice: the second pass re-uses communicators / pools
    sk.pack(c1, c2)
    sk.share_begin()
    dso = sk.width()
    sk.share_end()
    sig = np.array([c * dso / np.sqrt(2 * np.log(2)) for c in coeffs])
    kt, ks = sk.alloc_out(len(coeffs))
    sk.k_train(sig, kt)
    sk.k_test(sig, ks)
    y = sk.predict(ks, alpha)
torch.cuda.synchronize()

# the same problem directly on this GPU
r1 = fused.PackedRep(c1, train[1], train[0], mb, **kw)
r2 = fused.PackedRep(c2, test[1], test[0], mb, **kw)
dso_ref = fused.kernel_width([r1, r2], [r1, r2], zmax)
K1 = fused.local_gaussian(r1, r1, sig, zmax, symmetric=True)
K2 = fused.local_gaussian(r2, r1, sig, zmax)
y_ref = torch.einsum("stn,sn->st", K2, alpha)
e_w = float(np.max(np.abs(dso - dso_ref) / np.abs(dso_ref)))
e_t = 0.0
for (b, c), t in kt.items():
    (lo, hi), (clo, chi) = sk.bounds[b], sk.bounds[c]
    ref = K1[:, lo:hi, clo:chi]
    e_t = max(e_t, float(((t - ref).abs() / ref.abs()).max()))
e_s = 0.0
for (_, tb, b), t in ks.items():
    (lo, hi), (tlo, thi) = sk.bounds[b], sk.tbounds[tb]
    ref = K2[:, tlo:thi, lo:hi]
    e_s = max(e_s, float(((t - ref).abs() / ref.abs()).max()))
e_y = float(((y - y_ref).abs() / y_ref.abs()).max())
ok = e_w < 1e-12 and e_t < 1e-12 and e_s < 1e-12 and e_y < 1e-11
flags = torch.tensor([int(ok)], device=dev)
if world > 1:
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
print("sharded check rank %d/%d: width rel.err %.2e, K_train tiles %.2e, K_test columns %.2e, prediction %.2e -> %s"
      % (rank, world, e_w, e_t, e_s, e_y, "OK" if ok else "FAILED"), flush=True)
sk.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if int(flags.item()) else 1)
