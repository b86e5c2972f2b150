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


def allgather_rows(x_local, counts, group=None):
    """Row blocks of a 2-D tensor from all ranks, concatenated in rank order on every rank (rank r contributes
    ``counts[r]`` rows; ``counts`` is known everywhere).  SURVEY 8e: SLATM generation is sharded by molecules, then ONE
    all-gather gives every GPU the full column-side representation (7.2 GB for 100 k x 8975 over NVLink) -- the only
    collective of the global-kernel path."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x_local
    world = dist.get_world_size(group)
    counts = [int(c) for c in counts]
    assert len(counts) == world and x_local.shape[0] == counts[dist.get_rank(group)]
    nmax, D = max(counts), x_local.shape[1]
    out = torch.empty((world * nmax, D), dtype=x_local.dtype, device=x_local.device)
    mine = x_local.contiguous()
    if mine.shape[0] != nmax:
        pad = torch.zeros((nmax, D), dtype=x_local.dtype, device=x_local.device)
        pad[:mine.shape[0]] = mine
        mine = pad
    dist.all_gather([out[r * nmax:(r + 1) * nmax] for r in range(world)], mine, group=group)
    if all(c == nmax for c in counts):
        return out
    return torch.cat([out[r * nmax:r * nmax + counts[r]] for r in range(world)], dim=0)


def share_reps(reps, owners, group=None, rows=False):
    """Packed representations computed 1/N per rank, made complete on every rank ("SLATM: molecules are
    independent -> ranges per GPU, then one all-gather so every GPU holds the full column-side representation").
    reps[b] is a PackedRep on every rank -- computed where owners[b] == rank, laid out only (compute=False) elsewhere;
    its storage is one buffer whose layout depends only on the molecules, so block b is one broadcast from its owner.
    The broadcasts are enqueued together (NCCL runs them back to back on its stream) and waited for at the end."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    src = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    # keep the views alive until the collectives have been waited for; rows=True also ships the packed FP64 rows (needed
    # only by the FP64 engine and by a width pass over foreign blocks)
    views = [(r.slab_tensor(w), o) for r, o in zip(reps, owners) for w in ((0, 1) if rows else (0,))]
    works = [dist.broadcast(v, src=src(o), group=group, async_op=True) for v, o in views]
    for w in works:
        w.wait()


# ----------------------------------------------------------------------------------------------------------
# Distributed solve of (K + lambda I) alpha = y for a row-sharded K ("multi-GPU block
# Cholesky for 100 k").  The reference solves on one host with LAPACK (`np.linalg.solve`, aqml.py:875,
# krr.py:296); K(100k x 100k) is 80 GB per sigma, so here the matrix never leaves the GPUs that assembled it.
#
# Layout: row blocks of `nb` molecules dealt round-robin to the ranks (block b -> rank b % world), each rank
# holding FULL rows of its blocks -- exactly what `fused.local_gaussian(rep, rep, row0=, nrows=)` produces, so
# the assembly still needs no collective.  Factorisation: right-looking block Cholesky; per block column one
# broadcast of the nb x nb diagonal factor and one exchange of the (n - k) x nb panel (the path's only "real
# exchange steps" besides the prediction all-reduce), everything else is local cuSOLVER/cuBLAS through torch.

class RowBlockCyclic(object):
    """Which rows of an n x n matrix a rank owns: blocks of `nb` rows, block b on rank b % world."""

    def __init__(self, n, nb, rank, world):
        self.n, self.nb, self.rank, self.world = int(n), int(nb), int(rank), int(world)
        self.nblk = (self.n + self.nb - 1) // self.nb
        self.blocks = [b for b in range(self.nblk) if b % self.world == self.rank]
        self.local_start = {}
        o = 0
        for b in self.blocks:
            self.local_start[b] = o
            o += self.size(b)
        self.nloc = o

    def lo(self, b):
        return b * self.nb

    def hi(self, b):
        return min((b + 1) * self.nb, self.n)

    def size(self, b):
        return self.hi(b) - self.lo(b)

    def owner(self, b):
        return b % self.world

    def rows(self):
        """Global row indices of the local rows, in local order."""
        return np.concatenate([np.arange(self.lo(b), self.hi(b)) for b in self.blocks]) if self.blocks else np.zeros(0, dtype=int)

    def loc(self, b):
        """Local row range of an owned block."""
        s = self.local_start[b]
        return s, s + self.size(b)


def _world(group):
    import torch.distributed as dist
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def cholesky_cyclic_(A_loc, layout, group=None, lam=0.0):
    """In place: the lower triangle of the local rows of A + lam I becomes the local rows of its Cholesky
    factor L (A = L L^T).  A_loc (layout.nloc, n); only columns <= the row's own block are read / written."""
    import torch
    import torch.distributed as dist
    lay = layout
    multi = _world(group) > 1
    if lam:
        for b in lay.blocks:
            s, e = lay.loc(b)
            A_loc[s:e, lay.lo(b):lay.hi(b)].diagonal().add_(lam)
    for k in range(lay.nblk):
        k0, k1 = lay.lo(k), lay.hi(k)
        nk = k1 - k0
        # 1. diagonal block -> L_kk on its owner, broadcast
        Lkk = torch.empty((nk, nk), dtype=A_loc.dtype, device=A_loc.device)
        if lay.owner(k) == lay.rank:
            s, e = lay.loc(k)
            Lkk, info = torch.linalg.cholesky_ex(A_loc[s:e, k0:k1])
            A_loc[s:e, k0:k1] = Lkk
            Lkk = Lkk.contiguous()
            if int(info.item()) != 0:
                Lkk.fill_(float("nan"))  # tell every rank (they all raise, nobody is left waiting in a collective)
        if multi:
            dist.broadcast(Lkk, src=dist.get_global_rank(group, lay.owner(k)) if group is not None else lay.owner(k), group=group)
        if bool(torch.isnan(Lkk[0, 0])):
            raise np.linalg.LinAlgError(f"K + lambda I is not positive definite (block column {k})")
        if k1 >= lay.n:
            break
        # 2. my rows below block k: panel L[b, k] = A[b, k] L_kk^-T
        below = [b for b in lay.blocks if b > k]
        panel = torch.zeros((lay.n - k1, nk), dtype=A_loc.dtype, device=A_loc.device)
        if below:
            s = lay.loc(below[0])[0]
            mine = torch.linalg.solve_triangular(Lkk, A_loc[s:, k0:k1].T, upper=False).T  # X L^T = A
            A_loc[s:, k0:k1] = mine
            for b in below:
                bs, be = lay.loc(b)
                panel[lay.lo(b) - k1:lay.hi(b) - k1] = mine[bs - s:be - s]
        # 3. exchange: every rank needs the whole panel (rows of the others arrive through the sum)
        if multi:
            dist.all_reduce(panel, op=dist.ReduceOp.SUM, group=group)
        # 4. trailing update of my rows, lower triangle only: A[b, j] -= L[b, k] L[j, k]^T for k < j <= b
        for b in below:
            bs, be = lay.loc(b)
            A_loc[bs:be, k1:lay.hi(b)].addmm_(panel[lay.lo(b) - k1:lay.hi(b) - k1], panel[:lay.hi(b) - k1].T, alpha=-1.0)
    return A_loc


def cholesky_solve_cyclic(L_loc, layout, Y, group=None):
    """alpha = (L L^T)^-1 Y with L row-block-cyclic (output of `cholesky_cyclic_`).  Y (n,) or (n, m), the same on
    every rank; returns alpha with Y's shape on every rank."""
    import torch
    import torch.distributed as dist
    lay = layout
    multi = _world(group) > 1
    src = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    vec = Y.dim() == 1
    Z = (Y[:, None] if vec else Y).clone()
    # forward: z_k = L_kk^-1 (y_k - L[k, :k] z[:k]) on the owner of block k, then broadcast
    for k in range(lay.nblk):
        k0, k1 = lay.lo(k), lay.hi(k)
        zk = torch.empty((k1 - k0, Z.shape[1]), dtype=Z.dtype, device=Z.device)
        if lay.owner(k) == lay.rank:
            s, e = lay.loc(k)
            rhs = Z[k0:k1] - L_loc[s:e, :k0] @ Z[:k0]
            zk = torch.linalg.solve_triangular(L_loc[s:e, k0:k1], rhs, upper=False).contiguous()
        if multi:
            dist.broadcast(zk, src=src(lay.owner(k)), group=group)
        Z[k0:k1] = zk
    # backward: alpha_k = L_kk^-T (z_k - sum_{b>k} L[b, k]^T alpha_b); the sum runs over rows owned by all ranks
    X = torch.zeros_like(Z)
    acc = torch.zeros_like(Z)          # my share of  sum_b L[b, :]^T alpha_b  over my finished blocks
    for k in range(lay.nblk - 1, -1, -1):
        k0, k1 = lay.lo(k), lay.hi(k)
        part = acc[k0:k1].contiguous()
        if multi:
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
        xk = torch.empty((k1 - k0, Z.shape[1]), dtype=Z.dtype, device=Z.device)
        if lay.owner(k) == lay.rank:
            s, e = lay.loc(k)
            xk = torch.linalg.solve_triangular(L_loc[s:e, k0:k1].T, Z[k0:k1] - part, upper=True).contiguous()
        if multi:
            dist.broadcast(xk, src=src(lay.owner(k)), group=group)
        X[k0:k1] = xk
        if lay.owner(k) == lay.rank:
            s, e = lay.loc(k)
            acc[:k0] += L_loc[s:e, :k0].T @ xk
    return X[:, 0] if vec else X


def krr_solve_cyclic(K_loc, layout, Y, llambda, group=None):
    """alpha = (K + llambda I)^-1 Y (aqml.py:874-875) for a row-block-cyclic K; K_loc is overwritten by L."""
    cholesky_cyclic_(K_loc, layout, group=group, lam=llambda)
    return cholesky_solve_cyclic(K_loc, layout, Y, group=group)
