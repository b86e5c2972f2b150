"""KRR feed on N GPUs: everything between the coordinates and the solver of the reference's local-kernel workflow
(algo/aqml.py:711-789: aSLATM of train + test, kernel width over all atoms, ks1 = K(train, train), ks2 = K(test,
train)) plus the prediction ks2 . alpha (aqml.py:876), for ONE fixed problem cut over the ranks.

One process per GPU.  The training set is cut into 2 N blocks of molecules; rank r owns blocks r and 2 N - 1 - r
("folded" assignment: block row b of the lower triangle of K_train has b + 1 tiles, so every rank gets exactly
2 N + 1 tiles, two of them diagonal).

  phase       what a rank does                                                        exchange
  aslatm      packs its two training blocks (+ the test set)                          --
  allgather   --                                                                      each block's storage once (NCCL broadcast
                                                                                      of one buffer per block, `parallel.share_reps`)
  width       max distance per element over (its blocks) x (all blocks + test)        all-reduce(MAX) of (zmax,) doubles
  k_train     tiles (b, c), c <= b, of its block rows; diagonal tiles symmetric       --
  k_test      K(test, its training blocks): K_test sharded by training columns        --
  predict     partial K_test . alpha over its columns                                  all-reduce(SUM) of (nsigmas, n_test)

The kernel assembly itself needs no communication (north_star: "each GPU builds its K tile independently, only the
prediction is combined with an all-reduce"); the all-gather replaces the N-fold recomputation of the training
representation.  With world = 1 the same code runs without any collective.
"""
from __future__ import annotations

import numpy as np

from . import fused, parallel


def fold_owner(b, world):
    """Owner of training block b of 2 * world blocks."""
    return b if b < world else 2 * world - 1 - b


def block_bounds(n, nblk):
    """nblk contiguous molecule ranges of ~equal size: list of (lo, hi)."""
    return [parallel.shard_range(n, b, nblk) for b in range(nblk)]


def my_tiles(rank, world):
    """Lower-triangle tiles (b, c), c <= b, of the block rows this rank owns."""
    return [(b, c) for b in range(2 * world) if fold_owner(b, world) == rank for c in range(b + 1)]


class ShardedKernels(object):

    def __init__(self, train, test, mbtypes, zmax, rank=0, world=1, group=None, **slatm_kw):
        """train / test: (nas, zs) host arrays of the whole problem, the same on every rank."""
        self.rank, self.world, self.group = int(rank), int(world), group
        self.mbtypes, self.zmax, self.kw = mbtypes, int(zmax), slatm_kw
        self.nas1, self.zs1 = np.asarray(train[0]), np.asarray(train[1])
        self.nas2, self.zs2 = np.asarray(test[0]), np.asarray(test[1])
        self.n1, self.n2 = len(self.nas1), len(self.nas2)
        self.nblk = 2 * self.world
        self.bounds = block_bounds(self.n1, self.nblk)
        self.off1 = np.concatenate(([0], np.cumsum(self.nas1)))
        self.owners = [fold_owner(b, self.world) for b in range(self.nblk)]
        self.mine = [b for b in range(self.nblk) if self.owners[b] == self.rank]
        self.tiles = my_tiles(self.rank, self.world)
        self.reps, self.test = [None] * self.nblk, None

    # ---- phases ---------------------------------------------------------------------------------------------
    def _block(self, b):
        lo, hi = self.bounds[b]
        return self.nas1[lo:hi], self.zs1[self.off1[lo]:self.off1[hi]], slice(int(self.off1[lo]), int(self.off1[hi]))

    def pack(self, coords_train, coords_test):
        """aSLATM: my training blocks are computed, the others only laid out; the (small) test set is packed by every
        rank.  coords_*: (A, 3) host arrays (copied to the device inside) or CUDA tensors."""
        self.close()
        for b in range(self.nblk):
            nas, zs, sl = self._block(b)
            own = self.owners[b] == self.rank
            self.reps[b] = fused.PackedRep(coords_train[sl] if own else None, zs, nas, self.mbtypes, compute=own, **self.kw)
        self.test = fused.PackedRep(coords_test, self.zs2, self.nas2, self.mbtypes, **self.kw)

    def share(self):
        """Every training block travels once from its owner to the other ranks."""
        if self.world > 1:
            parallel.share_reps(self.reps, self.owners, group=self.group)

    def width(self):
        """dmax per element over all training + test atoms (get_kernel_width with dmxby='a', aqml.py:744-749)."""
        import torch
        import torch.distributed as dist
        everything = self.reps + [self.test]
        mine = [self.reps[b] for b in self.mine] + ([self.test] if self.rank == self.world - 1 else [])
        if self.world == 1:
            return fused.kernel_width(everything, everything, self.zmax)
        dso = torch.as_tensor(fused.kernel_width(mine, everything, self.zmax), device=self.test.device)
        dist.all_reduce(dso, op=dist.ReduceOp.MAX, group=self.group)
        return dso.cpu().numpy()

    def alloc_out(self, nsig):
        """Device buffers for my tiles of K_train and my column blocks of K_test."""
        import torch
        dev = self.test.device if self.test is not None else torch.device("cuda", torch.cuda.current_device())
        size = lambda b: self.bounds[b][1] - self.bounds[b][0]  # noqa: E731
        kt = {(b, c): torch.empty((nsig, size(b), size(c)), dtype=torch.float64, device=dev) for b, c in self.tiles}
        ks = {b: torch.empty((nsig, self.n2, size(b)), dtype=torch.float64, device=dev) for b in self.mine}
        return kt, ks

    def k_train(self, sig, out, after=None):
        """My tiles of the lower triangle of K(train, train); `after(key, tensor)` is called once a tile is enqueued."""
        for b, c in self.tiles:
            if b == c:
                fused.local_gaussian(self.reps[b], self.reps[b], sig, self.zmax, out=out[(b, c)], symmetric=True)
            else:
                fused.local_gaussian(self.reps[b], self.reps[c], sig, self.zmax, out=out[(b, c)])
            if after is not None:
                after((b, c), out[(b, c)])
        return out

    def k_test(self, sig, out, after=None):
        """K(test, my training blocks): K_test sharded by training columns."""
        for b in self.mine:
            fused.local_gaussian(self.test, self.reps[b], sig, self.zmax, out=out[b])
            if after is not None:
                after(b, out[b])
        return out

    def predict(self, k_test, alpha):
        """y[s, t] = sum_j K_test[s, t, j] alpha[(s,) j] over ALL training molecules: partial sums over my columns and
        one all-reduce(SUM) (aqml.py:876 with ks2 sharded by columns).  alpha: (n_train,) or (nsigmas, n_train)."""
        import torch
        y = None
        for b in self.mine:
            lo, hi = self.bounds[b]
            a = alpha[..., lo:hi]
            part = torch.matmul(k_test[b], a) if a.dim() == 1 else torch.einsum("stn,sn->st", k_test[b], a)
            y = part if y is None else y + part
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
        return y

    def close(self):
        for r in self.reps:
            if r is not None:
                r.close()
        if self.test is not None:
            self.test.close()
        self.reps, self.test = [None] * self.nblk, None

    # ---- bookkeeping for the bench ---------------------------------------------------------------------------
    def unique_elements(self, nsig):
        """Kernel elements this rank delivers per step: tiles of the lower triangle (diagonal tiles: their lower
        triangle incl. the diagonal) + its K_test columns."""
        size = lambda b: self.bounds[b][1] - self.bounds[b][0]  # noqa: E731
        n = 0
        for b, c in self.tiles:
            n += size(b) * (size(b) + 1) // 2 if b == c else size(b) * size(c)
        n += sum(self.n2 * size(b) for b in self.mine)
        return nsig * n
