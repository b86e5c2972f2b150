"""world_size-2 gloo tests of the multi-GPU host logic (row-block sharding, prediction all-reduce)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aqml_b200 import parallel


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 100001):
        for world in (1, 2, 3, 8):
            r = [parallel.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_shard_by_cost_balances():
    rng = np.random.default_rng(0)
    costs = rng.integers(1, 400, size=5000)
    parts = parallel.shard_by_cost(costs, 8)
    assert parts[0][0] == 0 and parts[-1][1] == 5000
    tot = [costs[l:h].sum() for l, h in parts]
    assert max(tot) < 1.05 * costs.sum() / 8 + 400


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1234)
        K = torch.rand((3, 11, 37), dtype=torch.float64, generator=g)       # (S, n_test, n_train) full
        alpha = torch.rand((37,), dtype=torch.float64, generator=g)
        alpha_s = torch.rand((3, 37), dtype=torch.float64, generator=g)
        lo, hi = parallel.shard_range(37, rank, world)
        y = parallel.predict_allreduce(K[:, :, lo:hi].contiguous(), alpha[lo:hi])
        assert torch.allclose(y, K @ alpha, rtol=1e-13, atol=0)
        y2 = parallel.predict_allreduce(K[:, :, lo:hi].contiguous(), alpha_s[:, lo:hi].contiguous())
        assert torch.allclose(y2, torch.einsum("stn,sn->st", K, alpha_s), rtol=1e-13, atol=0)
        rlo, rhi = parallel.shard_range(11, rank, world)
        full = parallel.gather_rows(K[:, rlo:rhi].contiguous())
        assert torch.equal(full, K)
        # representation row blocks: equal and ragged contributions
        X = torch.rand((23, 5), dtype=torch.float64, generator=g)
        for counts in ([12, 11], [20, 3]):
            lo2 = sum(counts[:rank])
            got = parallel.allgather_rows(X[lo2:lo2 + counts[rank]].contiguous(), counts)
            assert torch.equal(got, X)
        # distributed block-cyclic Cholesky solve of (K + lambda I) alpha = y, ragged last block, 1 and 3 right-hand sides
        n = 61
        M = torch.rand((n, n), dtype=torch.float64, generator=g)
        A = M @ M.T / n + torch.eye(n, dtype=torch.float64)
        Y = torch.rand((n, 3), dtype=torch.float64, generator=g)
        ref = torch.linalg.solve(A + 1e-4 * torch.eye(n, dtype=torch.float64), Y)
        for nb in (7, 16, 64):
            lay = parallel.RowBlockCyclic(n, nb, rank, world)
            assert sorted(np.concatenate([parallel.RowBlockCyclic(n, nb, r, world).rows() for r in range(world)])) == list(range(n))
            alpha = parallel.krr_solve_cyclic(A[lay.rows()].clone(), lay, Y, 1e-4)
            assert torch.allclose(alpha, ref, rtol=1e-11, atol=1e-13)
            a1 = parallel.krr_solve_cyclic(A[lay.rows()].clone(), lay, Y[:, 1].contiguous(), 1e-4)
            assert torch.allclose(a1, ref[:, 1], rtol=1e-11, atol=1e-13)
        lay = parallel.RowBlockCyclic(n, 16, rank, world)
        bad = A.clone()
        bad[40, 40] = -5.0
        try:
            parallel.cholesky_cyclic_(bad[lay.rows()].clone(), lay)
            raised = False
        except np.linalg.LinAlgError:
            raised = True
        assert raised  # on every rank, so nobody is left waiting in a collective
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gloo_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(2))


# ----------------------------------------------------------------------------------------------------------
# aqml_b200/sharded.py: the KRR feed cut over N ranks (host logic; the kernels are stand-ins on the CPU)
# ----------------------------------------------------------------------------------------------------------
def test_folded_blocks_balance_and_cover_the_triangle():
    from aqml_b200 import sharded
    for world in (1, 2, 3, 4, 8):
        nblk = 2 * world
        tiles = [sharded.my_tiles(r, world) for r in range(world)]
        # every rank: exactly 2N + 1 tiles, two of them diagonal
        assert all(len(t) == 2 * world + 1 and sum(b == c for b, c in t) == 2 for t in tiles)
        allt = sorted(t for ts in tiles for t in ts)
        assert allt == [(b, c) for b in range(nblk) for c in range(b + 1)]          # the lower block triangle, once
        assert sorted(sharded.fold_owner(b, world) for b in range(nblk)) == sorted(list(range(world)) * 2)
    bounds = sharded.block_bounds(20000, 16)
    assert bounds[0][0] == 0 and bounds[-1][1] == 20000 and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))


class _FakeRep(object):
    """Stands in for fused.PackedRep on the CPU: storage = one uint8 tensor, 'computed' only by its owner."""

    def __init__(self, nbytes, fill):
        self.buf = torch.full((nbytes,), fill, dtype=torch.uint8)
        self.nmol = nbytes

    def slab_tensor(self, which=0):
        return self.buf

    def close(self):
        pass


def _sharded_worker(rank, world, port, tmp):
    from aqml_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # share_reps: block b (ragged sizes) is computed by its owner only, every rank ends with every block
        owners = [sharded.fold_owner(b, world) for b in range(2 * world)]
        reps = [_FakeRep(1000 + 37 * b, (b + 1) if owners[b] == rank else 0) for b in range(2 * world)]
        parallel.share_reps(reps, owners)
        assert all(bool((r.buf == b + 1).all()) for b, r in enumerate(reps))
        # predict: K_test sharded by my training blocks + all-reduce == the full product
        nas1, nas2 = np.full(41, 3), np.full(7, 3)
        sk = sharded.ShardedKernels((nas1, np.ones(123, dtype=int)), (nas2, np.ones(21, dtype=int)), None, 9, rank, world)
        g = torch.Generator().manual_seed(7)
        K = torch.rand((2, 7, 41), dtype=torch.float64, generator=g)
        alpha = torch.rand((2, 41), dtype=torch.float64, generator=g)
        # K_test arrives as tiles (test block t of N, my training block b)
        assert [hi - lo for lo, hi in sk.tbounds] == [4, 3][:world] or world != 2
        ks = {("test", t, b): K[:, sk.tbounds[t][0]:sk.tbounds[t][1], sk.bounds[b][0]:sk.bounds[b][1]].contiguous()
              for b in sk.mine for t in range(world)}
        y = sk.predict(ks, alpha)
        assert torch.allclose(y, torch.einsum("stn,sn->st", K, alpha), rtol=1e-13, atol=0)
        y1 = sk.predict({k: v[0] for k, v in ks.items()}, alpha[0])
        assert torch.allclose(y1, K[0] @ alpha[0], rtol=1e-13, atol=0)
        # bookkeeping: the ranks' unique elements add up to the lower triangle + K_test
        n = torch.tensor([float(sk.unique_elements(2))], dtype=torch.float64)
        dist.all_reduce(n)
        assert int(n.item()) == 2 * (41 * 42 // 2 + 7 * 41)
        open(os.path.join(tmp, f"sharded_ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_feed_gloo_world2(tmp_path):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f"sharded_ok{r}").exists() for r in range(2))
