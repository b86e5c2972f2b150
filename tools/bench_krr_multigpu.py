#!/usr/bin/env python
"""KRR training + prediction with K_train sharded over GPUs, one process per GPU:

  molecules -> packed aSLATM (every rank) -> my row blocks of K_train (block-cyclic, no collective)
  -> distributed block Cholesky of K + lambda I (NCCL: one broadcast + one panel exchange per block column)
  -> alpha -> K(train rows of mine, test)^T alpha partial sums -> one all-reduce.

With --check rank 0 also solves the whole problem alone (symmetric assembly + torch.linalg.solve) and the
relative differences of alpha and of the predictions are reported.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_krr_multigpu.py --check
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aqml_b200 import fused, synth  # noqa: E402
from aqml_b200.algo.krr_sharded import ShardedKRR  # noqa: E402
from aqml_b200.representation import core as rcore  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n-train", type=int, default=16384)
ap.add_argument("--n-test", type=int, default=1024)
ap.add_argument("--nb", type=int, default=1024)
ap.add_argument("--llambda", type=float, default=1e-4)
ap.add_argument("--check", action="store_true")
args = ap.parse_args()

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    so = os.dup(1)
    os.dup2(2, 1)   # NCCL prints its version banner on stdout
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
    os.dup2(so, 1)

train = synth.qm9_like_fast(args.n_train, 11, pool=512)
test = synth.qm9_like_fast(args.n_test, 12, pool=128)
mb = rcore.get_slatm_mbtypes(synth.split_zs(train[0], train[1]) + synth.split_zs(test[0], test[1]))
kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
zmax = 9
y = torch.as_tensor(np.random.default_rng(3).normal(size=args.n_train), device=dev)


def timed(fn):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, float(t.item())


rep_test = fused.PackedRep(test[2], test[1], test[0], mb, **kw)
model, t_rep = timed(lambda: ShardedKRR(train[0], train[1], train[2], mb, zmax, nb=args.nb, rank=rank, world=world, **kw))
dso, t_width = timed(model.kernel_width)
sig = dso / np.sqrt(2 * np.log(2))
K_loc, t_asm = timed(lambda: model.assemble(sig))
alpha, t_fit = timed(lambda: model.fit(K_loc, y, args.llambda))
yp, t_pred = timed(lambda: model.predict(rep_test))

res = {"n_gpus": world, "n_train": args.n_train, "n_test": args.n_test, "nb": args.nb, "llambda": args.llambda,
       "ms": {"pack_aslatm": t_rep, "kernel_width": t_width, "assemble_lower": t_asm, "cholesky_and_solve": t_fit, "predict": t_pred},
       "total_ms": t_rep + t_width + t_asm + t_fit + t_pred}
if args.check:
    del K_loc
    model.L = None
    rep_train = fused.PackedRep(train[2], train[1], train[0], mb, **kw) if rank == 0 else None
    if rank == 0:
        sig = sig[None, :]
        Kf = fused.local_gaussian(rep_train, rep_train, sig, zmax, symmetric=True)[0]
        Kf.diagonal().add_(args.llambda)
        a_ref = torch.linalg.solve(Kf, y)
        y_ref = fused.local_gaussian(rep_test, rep_train, sig, zmax)[0] @ a_ref
        # alpha is ill-conditioned (cond(K + lambda I) ~ n / lambda); the prediction is the quantity with a tolerance
        res["check"] = {"alpha_rel": float((alpha - a_ref).abs().max() / a_ref.abs().max()),
                        "pred_rel": float((yp - y_ref).abs().max() / y_ref.abs().max()),
                        "residual_rel": float(((Kf @ alpha) - y).abs().max() / y.abs().max()),
                        "residual_rel_lu_single_gpu": float(((Kf @ a_ref) - y).abs().max() / y.abs().max()),
                        "dso_rel": float(np.abs(dso - fused.kernel_width(rep_train, rep_train, zmax)).max() / dso.max())}
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
