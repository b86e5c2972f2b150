"""KRR on the local aSLATM kernel with K_train sharded over GPUs (the reference's
single-host equivalent is aqml.py:744-789 + :874-876).

One process per GPU.  The training set is cut into blocks of ``nb`` molecules; block row b of the LOWER
triangle of K_train lives on rank b % world (`parallel.RowBlockCyclic`), which balances both the assembly of
the triangle and the Cholesky factorisation.  Every rank packs the aSLATM of every block itself (cheap next
to K), so kernel-width pass and kernel assembly need no data exchange at all; the collectives are

  * kernel width: none (the pruned max-distance pass is cheap enough to run redundantly on every rank),
  * Cholesky: per block column one broadcast (nb x nb) + one panel exchange ((n - k) x nb),
  * prediction: one all-reduce(SUM) of the (n_test,) partial predictions.

The solve uses the symmetric positive definite structure (Cholesky) where the reference calls LU
(`np.linalg.solve`): same alphas up to the conditioning of K + lambda I.
"""
from __future__ import annotations

import numpy as np

from .. import fused, parallel


def _take(nas, zs, coords, lo, hi):
    off = np.concatenate(([0], np.cumsum(nas)))
    return nas[lo:hi], zs[off[lo]:off[hi]], coords[off[lo]:off[hi]]


class ShardedKRR(object):

    def __init__(self, nas, zs, coords, mbtypes, zmax, nb=1024, rank=0, world=1, group=None, **slatm_kw):
        """nas/zs/coords: the training molecules (host arrays, the same on every rank)."""
        self.n = len(nas)
        self.zmax = int(zmax)
        self.group = group
        self.lay = parallel.RowBlockCyclic(self.n, nb, rank, world)
        self.mbtypes, self.slatm_kw = mbtypes, slatm_kw
        nas, zs, coords = np.asarray(nas), np.asarray(zs), np.asarray(coords)
        # every block on every rank: column operands of my rows, and the (pruned, hence cheap) width pass
        self.reps = []
        for c in range(self.lay.nblk):
            t = _take(nas, zs, coords, self.lay.lo(c), self.lay.hi(c))
            self.reps.append(fused.PackedRep(t[2], t[1], t[0], mbtypes, **slatm_kw))
        self.dso = None
        self.L = None

    def _multi(self):
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def kernel_width(self):
        """Per-element max aSLATM distance over all training atom pairs (get_kernel_width, aqml.py:204-255).
        The triangle-inequality pruning leaves ~1 % of the pairs, so every rank simply does the whole pass."""
        self.dso = fused.kernel_width(self.reps, self.reps, self.zmax)
        return self.dso

    def assemble(self, sigmas):
        """My rows of the lower triangle of K_train for ONE sigma set (zmax,) -> (nloc, n) device tensor
        (columns right of a row's own block are left uninitialised; the factorisation never reads them)."""
        import torch
        sig = np.asarray(sigmas, dtype=np.float64).reshape(1, self.zmax)
        dev = self.reps[0].device if self.reps else torch.device("cuda", torch.cuda.current_device())
        K = torch.empty((self.lay.nloc, self.n), dtype=torch.float64, device=dev)
        for b in self.lay.blocks:
            s, e = self.lay.loc(b)
            for c in range(b):
                K[s:e, self.lay.lo(c):self.lay.hi(c)] = fused.local_gaussian(self.reps[b], self.reps[c], sig, self.zmax)[0]
            K[s:e, self.lay.lo(b):self.lay.hi(b)] = fused.local_gaussian(self.reps[b], self.reps[b], sig, self.zmax,
                                                                         symmetric=True)[0]
        self.sig = sig
        return K

    def fit(self, K_loc, y, llambda=1e-4):
        """alpha = (K + lambda I)^-1 y; K_loc is overwritten by its Cholesky factor.  y: (n,) device tensor."""
        parallel.cholesky_cyclic_(K_loc, self.lay, group=self.group, lam=llambda)
        self.L = K_loc
        self.alpha = parallel.cholesky_solve_cyclic(K_loc, self.lay, y, group=self.group)
        return self.alpha

    def predict(self, rep_test, alpha=None):
        """y_test = K(test, train) alpha, K sharded by MY training blocks (columns) + one all-reduce."""
        import torch
        import torch.distributed as dist
        alpha = self.alpha if alpha is None else alpha
        yp = torch.zeros(rep_test.nmol, dtype=torch.float64, device=alpha.device)
        for b in self.lay.blocks:
            kb = fused.local_gaussian(rep_test, self.reps[b], self.sig, self.zmax)[0]  # (n_test, block)
            yp += kb @ alpha[self.lay.lo(b):self.lay.hi(b)]
        if self._multi():
            dist.all_reduce(yp, op=dist.ReduceOp.SUM, group=self.group)
        return yp
