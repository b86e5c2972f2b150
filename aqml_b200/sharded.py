"""KRR feed on N GPUs: everything between the coordinates and the solver of the reference's local-kernel workflow
(algo/aqml.py:711-789: aSLATM of train + test, kernel width over all atoms, ks1 = K(train, train), ks2 = K(test,
train)) plus the prediction ks2 . alpha (aqml.py:876), for ONE fixed problem cut over the ranks.

One process per GPU.  The training set is cut into 2 N blocks of molecules; rank r owns blocks r and 2 N - 1 - r
("folded" assignment: block row b of the lower triangle of K_train has b + 1 tiles, so every rank gets exactly
2 N + 1 tiles, two of them diagonal).

  phase       what a rank does                                                        exchange
  aslatm      packs its two training blocks and its 1/N of the test molecules (own     --
              blocks first; the others are only laid out, by a few host threads)
  allgather   copies the arrived operands into the blocks laid out for them            each block's pair-kernel operands (digit planes,
                                                                                      norms, masks: 6 of the 14 bytes per packed
                                                                                      element) once: three NCCL all-gathers into
                                                                                      equal slots, enqueued BEFORE ...
  width       ... the max distance per element over all train + test atoms, computed   all elements per round together: all-reduce of
              where the FP64 rows are (every rank sweeps only its own rows):           the column sums, two all-gathers per sweep,
              distances to the common mean, two farthest-point sweeps for an           candidate counts and rows, all-reduce MAX of the
              attained lower bound, then only the rows the triangle inequality         results -- on a second communicator, so that it
              leaves are gathered and contracted exactly, 1/N of them per rank          overlaps the exchange of the blocks
  k_train     tiles (b, c), c <= b, of its block rows; diagonal tiles symmetric       --
  k_test      K(test block t, its training blocks) for all N test blocks: K_test      --
              sharded by training columns
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
        # the test set is cut into N blocks, block t packed by rank t (at N = 8 the 2 k test molecules would otherwise cost every
        # rank almost as much aSLATM time as its 2.5 k training molecules)
        self.tbounds = block_bounds(self.n2, self.world)
        self.off2 = np.concatenate(([0], np.cumsum(self.nas2)))
        self.reps, self.tests = [None] * self.nblk, [None] * self.world
        self.group2, self._works, self._slots, self._pool = None, [], [], None

    # ---- phases ---------------------------------------------------------------------------------------------
    def _block(self, b):
        lo, hi = self.bounds[b]
        return self.nas1[lo:hi], self.zs1[self.off1[lo]:self.off1[hi]], slice(int(self.off1[lo]), int(self.off1[hi]))

    def pack(self, coords_train, coords_test):
        """aSLATM: my training blocks and my test block are computed, the others only laid out.
        coords_*: (A, 3) host arrays (copied to the device inside) or CUDA tensors."""
        self.close()

        def train(b):
            nas, zs, sl = self._block(b)
            own = self.owners[b] == self.rank
            self.reps[b] = fused.PackedRep(coords_train[sl] if own else None, zs, nas, self.mbtypes, compute=own, rows=own, **self.kw)

        def test(t):
            lo, hi = self.tbounds[t]
            sl = slice(int(self.off2[lo]), int(self.off2[hi]))
            own = t == self.rank
            self.tests[t] = fused.PackedRep(coords_test[sl] if own else None, self.zs2[sl], self.nas2[lo:hi], self.mbtypes, compute=own,
                                            rows=own, **self.kw)

        # my own blocks first: their aSLATM kernels run while the host lays out the 3 N - 3 blocks that will arrive over NCCL
        for b in self.mine:
            train(b)
        test(self.rank)
        others = [(train, b) for b in range(self.nblk) if self.reps[b] is None] + [(test, t) for t in range(self.world) if self.tests[t] is None]
        if len(others) > 4:
            # pure host work (index building, two stream-ordered allocations each): a few threads, the C side releases the GIL
            import torch
            from concurrent.futures import ThreadPoolExecutor
            dev = self.test.device

            def run(job):
                with torch.cuda.device(dev):
                    job[0](job[1])
            if self._pool is None:
                self._pool = ThreadPoolExecutor(max_workers=4)
            list(self._pool.map(run, others))
        else:
            for f, i in others:
                f(i)

    @property
    def test(self):
        """My block of the test set (the whole test set with one rank)."""
        return self.tests[self.rank]

    def _slab_bound(self, zs, like):
        """Upper bound of the operand storage (aqml_rep_slab(..., 0)) of a block with element numbers `zs`, from the row
        pitches of a representation `like` built with the same many-body types -- so that the exchange can start before the
        other ranks' blocks have been laid out here."""
        import ctypes as C
        from . import _lib as L
        total = 4096
        for e in range(int(L.load().aqml_rep_elements(like._h))):
            label, n, ld, x = C.c_int(), C.c_int(), C.c_int(), C.c_void_p()
            L.call("aqml_rep_element", like._h, e, C.byref(label), C.byref(n), C.byref(ld), C.byref(x))
            nz = int(np.count_nonzero(zs == label.value)) + 128          # packed rows incl. the padding at the element's end
            total += (nz // 128 + 2) * (ld.value // 16 + 1) * 12288        # digit planes: 6 x 2 KB per 128-row block and 16-column chunk
            total += nz * 16 + (nz // 64 + 2) * ((ld.value // 16 + 31) // 32) * 4 + 5 * 256
        return total

    def _exchange_start(self):
        """Enqueue the exchange: every block's pair-kernel operands travel once from their owner to the other ranks (the packed
        FP64 rows stay where they were computed).  Three all-gathers -- the ranks' first training blocks, their second ones, their
        test blocks -- into buffers of N equal slots (a block's operand storage is one buffer whose size depends only on its
        molecules; the slot size is a bound every rank can compute from the element counts): one all-gather per class instead of
        3 N broadcasts.  `share_end` waits and copies the slots into the blocks."""
        self._works, self._slots = [], []
        if self.world == 1:
            return
        import torch
        import torch.distributed as dist
        W = self.world
        tz = lambda t: self.zs2[int(self.off2[self.tbounds[t][0]]):int(self.off2[self.tbounds[t][1]])]  # noqa: E731
        classes = [([r for r in range(W)], self.reps, lambda b: self._block(b)[1]),                    # block r of rank r
                   ([2 * W - 1 - r for r in range(W)], self.reps, lambda b: self._block(b)[1]),        # block 2N-1-r of rank r
                   ([t for t in range(W)], self.tests, tz)]                                              # test block r of rank r
        for ids, store, zs_of in classes:
            mine = store[ids[self.rank]]
            size = max(self._slab_bound(np.asarray(zs_of(i)), mine) for i in ids)
            size = (size + 255) // 256 * 256
            view = mine.slab_tensor(0)
            if view.numel() > size:
                raise RuntimeError("internal: operand storage larger than its bound")
            buf = torch.empty((W, size), dtype=torch.uint8, device=view.device)
            buf[self.rank, :view.numel()].copy_(view)
            work = dist.all_gather_into_tensor(buf.view(-1), buf[self.rank], group=self.group, async_op=True)
            self._works.append(work)
            self._slots.append((buf, ids, store))

    def share_begin(self):
        """Enqueue the exchange (after `pack`).  Starting it inside `pack`, right after this rank's own blocks and before the host
        lays out the others, was measured slower (N = 4: 278 vs 264 ms per step: the width pass's small collectives then queue
        behind the all-gathers for longer)."""
        if self.world > 1 and not self._works and not self._slots:
            self._exchange_start()

    def share_end(self):
        for w in self._works:
            w.wait()
        for buf, ids, store in self._slots:
            for r, i in enumerate(ids):
                if r != self.rank:
                    v = store[i].slab_tensor(0)
                    if v.numel() > buf.shape[1]:
                        raise RuntimeError("internal: operand storage larger than its bound")
                    v.copy_(buf[r, :v.numel()])
        self._works, self._slots = [], []

    def share(self):
        self.share_begin()
        self.share_end()

    def width(self):
        """dmax per element over all training + test atoms (get_kernel_width with dmxby='a', aqml.py:744-749)."""
        if self.world == 1:
            everything = self.reps + self.tests
            return fused.kernel_width(everything, everything, self.zmax)
        return self._width_distributed()

    def _width_distributed(self):
        """The pruned max-distance pass of `aqml_rep_kernel_width_multi` over rows that live on different GPUs: nobody ever
        holds all rows.  Exact: the candidates are contracted by the same EPI_MAX pair kernel.  All elements go through every
        round together, so the pass costs seven small collectives (not six per element) and as many host synchronisations; the
        candidate contraction itself is cut over the ranks."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib as L
        if self.group2 is None:   # a second communicator: these small collectives must not queue behind the block broadcasts
            self.group2 = dist.new_group(list(range(self.world))) if self.group is None else self.group
        g, dev, W = self.group2, self.test.device, self.world
        mine = [self.reps[b] for b in self.mine] + [self.test]
        rows = [r.element_rows() for r in mine]
        labels = sorted(int(t[0]) for t in self.mbtypes if len(t) == 1)
        st = L.current_stream(dev)
        dso = np.full(self.zmax, 1.e-9)
        f64 = dict(dtype=torch.float64, device=dev)
        segs = {z: [d[z] for d in rows if z in d] for z in labels}           # my row segments per element
        for z in labels:
            if not segs[z]:
                raise RuntimeError("internal: element %d missing from a packed block" % z)
        ncol = {z: int(segs[z][0].shape[1]) for z in labels}
        coff = np.concatenate(([0], np.cumsum([ncol[z] for z in labels])))    # offsets into a concatenated "one row per element"
        ctot = int(coff[-1])

        def gather(t):   # all ranks' copies of a small tensor, stacked: (world, ...)
            out = [torch.empty_like(t) for _ in range(W)]
            dist.all_gather(out, t.contiguous(), group=g)
            return torch.stack(out)

        def sweep(points):   # distance of every local row of every element to that element's point: {z: [per segment (n,)]}
            res = {}
            for k, z in enumerate(labels):
                p = points[coff[k]:coff[k + 1]].contiguous()
                out = []
                for x in segs[z]:
                    r = torch.empty(x.shape[0], **f64)
                    if x.shape[0]:
                        L.call("aqml_rows_dist_dev", 2, L.ptr(x), x.stride(0), x.shape[0], L.ptr(p), x.shape[1], L.ptr(r), st)
                    out.append(r)
                res[z] = out
            return res

        def farthest(dists):
            """Per element: the global maximum of `dists` (device tensor (nel,), NaN if any rank saw a NaN) and the row that attains
            it, known to every rank -- without a host synchronisation: one all-gather, everything else stays on the device."""
            msg = torch.full((len(labels), 2), -1.0, **f64)
            best = torch.zeros(ctot, **f64)
            for k, z in enumerate(labels):
                ms, rws, nan = [], [], None
                for x, v in zip(segs[z], dists[z]):
                    if not v.numel():
                        continue
                    m, i = torch.max(v, dim=0)
                    ms.append(m)
                    rws.append(x.index_select(0, i.view(1)))
                    f = torch.isnan(v).any()
                    nan = f if nan is None else (nan | f)
                if ms:
                    mm = torch.stack(ms)
                    j = torch.argmax(mm).view(1)
                    msg[k, 0] = mm.index_select(0, j)[0]
                    msg[k, 1] = nan.to(torch.float64)
                    best[coff[k]:coff[k + 1]] = torch.cat(rws).index_select(0, j)[0]
            both = gather(torch.cat([msg.view(-1), best]))                   # one collective for the values and the rows
            allm, allb = both[:, :2 * len(labels)].reshape(W, len(labels), 2), both[:, 2 * len(labels):]
            owner = torch.argmax(allm[:, :, 0], dim=0)                        # (nel,)
            vmax = allm[:, :, 0].max(dim=0).values
            vmax = torch.where(allm[:, :, 1].max(dim=0).values > 0, torch.full_like(vmax, float("nan")), vmax)
            point = torch.cat([allb.index_select(0, owner[k:k + 1])[0, coff[k]:coff[k + 1]] for k in range(len(labels))])
            return vmax, point

        import os
        import time
        dbg = bool(os.environ.get("AQML_WIDTH_DEBUG"))
        tms = []

        def tick(name):
            if dbg:
                torch.cuda.synchronize()
                tms.append((name, time.perf_counter()))
        tick("start")
        # round 1: the common mean of every element's rows (the row counts of the whole problem are known on the host)
        ntot = np.array([int((self.zs1 == z).sum() + (self.zs2 == z).sum()) for z in labels])
        stat = torch.zeros(ctot, **f64)
        for k, z in enumerate(labels):
            for x in segs[z]:
                if x.shape[0]:
                    L.call("aqml_rows_colsum_dev", L.ptr(x), x.stride(0), x.shape[0], ncol[z], L.ptr(stat[coff[k]:coff[k + 1]]), st)
        dist.all_reduce(stat, group=g)
        scale = torch.as_tensor(np.repeat(1.0 / np.maximum(ntot, 1), [ncol[z] for z in labels]), **f64)
        centre = stat * scale
        tick("mean")
        # rounds 2-4: distances to the mean, two farthest-point sweeps for an attained pair distance
        r = sweep(centre)
        rmax, p1 = farthest(r)
        l1, p2 = farthest(sweep(p1))
        l2, _ = farthest(sweep(p2))
        low = torch.maximum(l1, l2) * (1.0 - 1e-9)
        tick("sweeps")
        # rounds 5-6: only rows with r_i + max r >= the attained distance can take part in the maximum: gather and contract them
        cands = []
        for k, z in enumerate(labels):
            parts = [x[(v + rmax[k]) >= low[k]] for x, v in zip(segs[z], r[z]) if v.numel()]   # NaN compares false: no candidates
            cands.append(torch.cat(parts) if parts else torch.empty((0, ncol[z]), **f64))
        cnt = torch.tensor([c.shape[0] for c in cands], dtype=torch.int64, device=dev)
        info = gather(torch.cat([cnt.to(torch.float64), rmax])).cpu().numpy()   # the one host synchronisation before the contraction
        cnts, rmax_h = info[:, :len(labels)].astype(np.int64), info[0, len(labels):]
        kmax = cnts.max(axis=0)
        poff = np.concatenate(([0], np.cumsum([int(kmax[k]) * ncol[z] for k, z in enumerate(labels)])))
        pad = torch.zeros(int(poff[-1]), **f64)
        for k, z in enumerate(labels):
            if cands[k].shape[0]:
                pad[poff[k]:poff[k] + cands[k].numel()] = cands[k].reshape(-1)
        allp = gather(pad)                                                    # (world, sum_k kmax_k ncol_k)
        tick("candidates")
        for k, z in enumerate(labels):
            if ntot[k] == 0:
                continue
            if rmax_h[k] != rmax_h[k]:
                dso[z - 1] = float("nan")
                continue
            blk = allp[:, poff[k]:poff[k + 1]].reshape(W, int(kmax[k]), ncol[z])
            allc = torch.cat([blk[w, :int(cnts[w, k])] for w in range(W)]).contiguous()
            # every rank contracts its 1/N of the candidate rows against all of them; the maxima meet in one all-reduce below
            part = allc[self.rank::W].contiguous()
            dm = C.c_double(0.0)
            if part.shape[0]:
                with torch.cuda.device(dev):
                    L.call("aqml_max_distance_dev", 2, L.ptr(part), part.shape[0], part.stride(0), L.ptr(allc), allc.shape[0], allc.stride(0),
                           ncol[z], C.c_void_p(C.addressof(dm)), st)
            dso[z - 1] = dm.value
        tick("contraction")
        if dbg and self.rank == 0:
            print("width: " + " ".join("%s=%.2f ms" % (n, (t - tms[i][1]) * 1e3) for i, (n, t) in enumerate(tms[1:])) +
                  " candidates " + str(cnts.sum(axis=0).tolist()), flush=True)
        red = torch.as_tensor(np.where(np.isnan(dso), np.inf, dso), **f64)    # MAX does not propagate NaN: send it as +inf
        dist.all_reduce(red, op=dist.ReduceOp.MAX, group=g)
        dso = red.cpu().numpy()
        dso[np.isinf(dso)] = np.nan
        return dso

    def alloc_out(self, nsig):
        """Device buffers for my tiles of K_train and my column blocks of K_test."""
        import torch
        dev = self.test.device if self.test is not None else torch.device("cuda", torch.cuda.current_device())  # noqa: E501
        size = lambda b: self.bounds[b][1] - self.bounds[b][0]  # noqa: E731
        tsize = lambda t: self.tbounds[t][1] - self.tbounds[t][0]  # noqa: E731
        kt = {(b, c): torch.empty((nsig, size(b), size(c)), dtype=torch.float64, device=dev) for b, c in self.tiles}
        ks = {("test", t, b): torch.empty((nsig, tsize(t), size(b)), dtype=torch.float64, device=dev)
              for b in self.mine for t in range(self.world)}
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
        """K(test block t, my training blocks), t = 0 .. N-1: K_test sharded by training columns."""
        for b in self.mine:
            for t in range(self.world):
                key = ("test", t, b)
                fused.local_gaussian(self.tests[t], self.reps[b], sig, self.zmax, out=out[key])
                if after is not None:
                    after(key, out[key])
        return out

    def predict(self, k_test, alpha):
        """y[s, t] = sum_j K_test[s, t, j] alpha[(s,) j] over ALL training molecules: partial sums over my columns and
        one all-reduce(SUM) (aqml.py:876 with ks2 sharded by columns).  alpha: (n_train,) or (nsigmas, n_train)."""
        import torch
        y = None
        for b in self.mine:
            lo, hi = self.bounds[b]
            a = alpha[..., lo:hi]
            for t in range(self.world):
                k = k_test[("test", t, b)]
                part = torch.matmul(k, a) if a.dim() == 1 else torch.einsum("stn,sn->st", k, a)
                if y is None:
                    y = torch.zeros(tuple(part.shape[:-1]) + (self.n2,), dtype=part.dtype, device=part.device)
                tlo, thi = self.tbounds[t]
                y[..., tlo:thi] += part
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
        return y

    def close(self):
        for r in self.reps:
            if r is not None:
                r.close()
        for r in self.tests:
            if r is not None:
                r.close()
        self.reps, self.tests = [None] * self.nblk, [None] * self.world

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
