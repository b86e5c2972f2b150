"""Multi-GPU plumbing (one process per GPU, torch.distributed: NCCL over NVLink on GPUs, gloo in
the CPU tests).

The kernel matrix shards by row blocks of the molecule set: every rank builds
K[rows_r, :] independently -- there is NO collective on the kernel-assembly path.  The only
exchange step of the workflow is the prediction y = K_test . alpha when K_test is sharded along
the training (column) axis: one all-reduce(sum) of an (nsigmas, n_test) vector.
"""
from __future__ import annotations

import numpy as np


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) of ``n`` items for ``rank``."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_cost(costs, world):
    """Contiguous partition of items with the given costs into ``world`` ranges of ~equal total cost.
    Used to balance SLATM generation (cost ~ neighbour pairs) and local-kernel rows (cost ~ atoms).
    Returns a list of (lo, hi)."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    csum = np.concatenate(([0.0], np.cumsum(costs)))
    total = csum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(csum, target, side="left"))
        i = min(max(i, cuts[-1]), n)
        cuts.append(i)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def predict_allreduce(K_cols, alpha_cols, group=None):
    """y[s, t] = sum over ALL ranks' training columns of K_cols[s, t, :] @ alpha_cols[s or :, :].

    K_cols (nsigmas, n_test, n_train_local), alpha_cols (n_train_local,) or (nsigmas, n_train_local);
    krr.py:299 / aqml.py:876 ``np.dot(k2, alphas)`` with k2 sharded by columns."""
    import torch
    import torch.distributed as dist
    if alpha_cols.dim() == 1:
        y = torch.matmul(K_cols, alpha_cols)
    else:
        y = torch.einsum("stn,sn->st", K_cols, alpha_cols)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y


def gather_rows(K_rows, group=None):
    """Concatenate row blocks K[:, rows_r, :] from all ranks along the row axis (ragged allowed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return K_rows
    world = dist.get_world_size(group)
    n = torch.tensor([K_rows.shape[1]], device=K_rows.device, dtype=torch.int64)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    nmax = int(max(int(v.item()) for v in ns))
    pad = torch.zeros((K_rows.shape[0], nmax, K_rows.shape[2]), dtype=K_rows.dtype, device=K_rows.device)
    pad[:, :K_rows.shape[1]] = K_rows
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:, :int(v.item())] for o, v in zip(outs, ns)], dim=1)
