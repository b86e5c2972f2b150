#!/usr/bin/env python
"""Multi-GPU check (torchrun, NCCL): K_test sharded (a) by test rows -> gather, (b) by training columns ->
partial predictions + one all-reduce.  Both must equal the single-GPU result computed on rank 0.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multigpu.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aqml_b200 import fused, parallel, synth  # noqa: E402
from aqml_b200.representation import core as rcore  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
n_train, n_test = 600, 90
train = synth.qm9_like_fast(n_train, 5, pool=128)
test = synth.qm9_like_fast(n_test, 6, pool=64)
mb = rcore.get_slatm_mbtypes(synth.split_zs(train[0], train[1]) + synth.split_zs(test[0], test[1]))
kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
sig = np.array([c * np.full(9, 16.0) / np.sqrt(2 * np.log(2)) for c in (1.0, 2.0)])
alpha = torch.as_tensor(np.random.default_rng(0).normal(size=n_train), device=dev)


def take(s, lo, hi):
    off = np.concatenate(([0], np.cumsum(s[0])))
    return s[0][lo:hi], s[1][off[lo]:off[hi]], s[2][off[lo]:off[hi]]


rep_train = fused.PackedRep(train[2], train[1], train[0], mb, **kw)
# (a) row blocks of the test set
lo, hi = parallel.shard_range(n_test, rank, world)
t = take(test, lo, hi)
rep_rows = fused.PackedRep(t[2], t[1], t[0], mb, **kw)
K_rows = fused.local_gaussian(rep_rows, rep_train, sig, 9)
K_a = parallel.gather_rows(K_rows)
# (b) column blocks of the training set, partial predictions + all-reduce
clo, chi = parallel.shard_range(n_train, rank, world)
c = take(train, clo, chi)
rep_cols = fused.PackedRep(c[2], c[1], c[0], mb, **kw)
rep_test = fused.PackedRep(test[2], test[1], test[0], mb, **kw)
K_cols = fused.local_gaussian(rep_test, rep_cols, sig, 9)
y_b = parallel.predict_allreduce(K_cols, alpha[clo:chi])
# single-GPU reference
K_full = fused.local_gaussian(rep_test, rep_train, sig, 9)
y_full = K_full @ alpha
ea = float((K_a - K_full).abs().max() / K_full.abs().max())
eb = float((y_b - y_full).abs().max() / y_full.abs().max())
ok = ea < 1e-13 and eb < 1e-12
flags = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print("multigpu check: world=%d rows-gather rel.err %.2e, all-reduce prediction rel.err %.2e -> %s"
          % (world, ea, eb, "OK" if int(flags.item()) else "FAIL"))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flags.item()) else 1)
