"""KRR feed on N GPUs: everything between the coordinates and the solver of the reference's local-kernel workflow
(algo/aqml.py:711-789: aSLATM of train + test, kernel width over all atoms, ks1 = K(train, train), ks2 = K(test,
train)) plus the prediction ks2 . alpha (aqml.py:876), for ONE fixed problem cut over the ranks.

One process per GPU.  The training set is cut into 2 N blocks of molecules; rank r owns blocks r and 2 N - 1 - r
("folded" assignment: block row b of the lower triangle of K_train has b + 1 tiles, so every rank gets exactly
2 N + 1 tiles, two of them diagonal).

  phase       what a rank does                                                        exchange
  aslatm      packs its two training blocks (+ the test set)                          --
  allgather   --                                                                      each block's pair-kernel operands (digit planes,
                                                                                      norms, masks: 6 of the 14 bytes per packed
                                                                                      element) once: one NCCL broadcast per block
                                                                                      (`parallel.share_reps`), enqueued BEFORE ...
  width       ... the max distance per element over all train + test atoms, computed   per element: all-reduce of the column sums,
              where the FP64 rows are (every rank sweeps only its own rows):           two row broadcasts, small all-gathers
              distances to the common mean, two farthest-point sweeps for an           (candidates: a few % of the rows)
              attained lower bound, then only the rows the triangle inequality         -- on a second communicator, so that it
              leaves are gathered and contracted exactly (same pruning as on one GPU)   overlaps the broadcasts of the blocks
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
        self.group2, self._works, self._views = None, [], []

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
            self.reps[b] = fused.PackedRep(coords_train[sl] if own else None, zs, nas, self.mbtypes, compute=own, rows=own, **self.kw)
        self.test = fused.PackedRep(coords_test, self.zs2, self.nas2, self.mbtypes, **self.kw)

    def share_begin(self):
        """Enqueue the broadcasts: every training block's pair-kernel operands travel once from their owner to the other
        ranks (the packed FP64 rows stay where they were computed)."""
        self._works = []
        if self.world > 1:
            import torch.distributed as dist
            src = (lambda r: dist.get_global_rank(self.group, r)) if self.group is not None else (lambda r: r)
            self._views = [r.slab_tensor(0) for r in self.reps]
            self._works = [dist.broadcast(v, src=src(o), group=self.group, async_op=True) for v, o in zip(self._views, self.owners)]

    def share_end(self):
        for w in self._works:
            w.wait()
        self._works, self._views = [], []

    def share(self):
        self.share_begin()
        self.share_end()

    def width(self):
        """dmax per element over all training + test atoms (get_kernel_width with dmxby='a', aqml.py:744-749)."""
        if self.world == 1:
            everything = self.reps + [self.test]
            return fused.kernel_width(everything, everything, self.zmax)
        return self._width_distributed()

    def _width_distributed(self):
        """The pruned max-distance pass of `aqml_rep_kernel_width_multi` over rows that live on different GPUs: nobody ever
        holds all rows.  Exact: the candidates are contracted by the same EPI_MAX pair kernel."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib as L
        if self.group2 is None:   # a second communicator: these small collectives must not queue behind the block broadcasts
            self.group2 = dist.new_group(list(range(self.world))) if self.group is None else self.group
        g, dev = self.group2, self.test.device
        mine = [self.reps[b] for b in self.mine] + ([self.test] if self.rank == self.world - 1 else [])
        rows = [r.element_rows() for r in mine]
        labels = sorted(int(t[0]) for t in self.mbtypes if len(t) == 1)
        st = L.current_stream(dev)
        dso = np.full(self.zmax, 1.e-9)

        def sweep(segs, point):   # distance of every local row to one point
            out = []
            for x in segs:
                r = torch.empty(x.shape[0], dtype=torch.float64, device=dev)
                if x.shape[0]:
                    L.call("aqml_rows_dist_dev", 2, L.ptr(x), x.stride(0), x.shape[0], L.ptr(point), x.shape[1], L.ptr(r), st)
                out.append(r)
            return torch.cat(out) if out else torch.empty(0, dtype=torch.float64, device=dev)

        def global_max(v):   # (max over all ranks, owner rank, local index on the owner); NaN anywhere -> NaN
            loc = torch.full((3,), -1.0, dtype=torch.float64, device=dev)
            if v.numel():
                m, i = torch.max(v, dim=0)
                loc[0], loc[1] = m, i.to(torch.float64)
                loc[2] = torch.isnan(v).any().to(torch.float64)
            allv = [torch.empty_like(loc) for _ in range(self.world)]
            dist.all_gather(allv, loc, group=g)
            tab = torch.stack(allv).cpu().numpy()
            if (tab[:, 2] > 0).any():
                return float("nan"), 0, 0
            owner = int(np.argmax(tab[:, 0]))
            return float(tab[owner, 0]), owner, int(tab[owner, 1])

        def row_from(owner, idx, allrows, ncols):
            p = torch.empty(ncols, dtype=torch.float64, device=dev)
            if owner == self.rank:
                p.copy_(allrows(idx))
            dist.broadcast(p, src=owner if self.group is None else dist.get_global_rank(g, owner), group=g)
            return p

        for z in labels:
            segs = [d[z] for d in rows if z in d]
            if not segs:
                raise RuntimeError("internal: element %d missing from a packed block" % z)
            ncols = int(segs[0].shape[1])
            nloc = sum(int(x.shape[0]) for x in segs)
            offs = np.concatenate(([0], np.cumsum([int(x.shape[0]) for x in segs])))

            def local_row(i):
                k = int(np.searchsorted(offs, i, side="right") - 1)
                return segs[k][i - offs[k]]

            stat = torch.zeros(ncols + 1, dtype=torch.float64, device=dev)
            for x in segs:
                if x.shape[0]:
                    L.call("aqml_rows_colsum_dev", L.ptr(x), x.stride(0), x.shape[0], ncols, L.ptr(stat), st)
            stat[ncols] = nloc
            dist.all_reduce(stat, group=g)
            ntot = int(round(float(stat[ncols].item())))
            if ntot == 0:
                continue
            centre = (stat[:ncols] / ntot).contiguous()
            r = sweep(segs, centre)
            rmax, o1, i1 = global_max(r)
            if rmax != rmax:
                dso[z - 1] = float("nan")
                continue
            t1 = sweep(segs, row_from(o1, i1, local_row, ncols))
            l1, o2, i2 = global_max(t1)
            t2 = sweep(segs, row_from(o2, i2, local_row, ncols))
            l2, _, _ = global_max(t2)
            low = max(l1, l2) * (1.0 - 1e-9)    # an attained distance: only rows with r_i + max r >= low can take part in the maximum
            sel = torch.nonzero(r + rmax >= low).flatten()
            cand = torch.empty((0, ncols), dtype=torch.float64, device=dev)
            if nloc and sel.numel():   # gather the survivors segment by segment (never a copy of all local rows)
                cut = torch.searchsorted(sel, torch.as_tensor(offs, device=dev)).tolist()
                cand = torch.cat([x[sel[cut[k]:cut[k + 1]] - int(offs[k])] for k, x in enumerate(segs)])
            cnt = torch.tensor([cand.shape[0]], dtype=torch.int64, device=dev)
            cnts = [torch.empty_like(cnt) for _ in range(self.world)]
            dist.all_gather(cnts, cnt, group=g)
            cnts = [int(c.item()) for c in cnts]
            kmax = max(cnts)
            pad = torch.zeros((kmax, ncols), dtype=torch.float64, device=dev)
            pad[:cand.shape[0]] = cand
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad, group=g)
            allc = torch.cat([p[:c] for p, c in zip(parts, cnts)]).contiguous()
            dm = C.c_double(0.0)
            with torch.cuda.device(dev):
                L.call("aqml_max_distance_dev", 2, L.ptr(allc), allc.shape[0], allc.stride(0), L.ptr(allc), allc.shape[0], allc.stride(0),
                       ncols, C.c_void_p(C.addressof(dm)), st)
            dso[z - 1] = dm.value
        return dso

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
