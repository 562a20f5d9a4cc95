"""Host-side multi-GPU logic on CPU: sharding schemes + the packed-token all-gather over gloo (world_size 2)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bitdance_b200 import parallel as par


def test_shard_schemes_cover_everything_once():
    for n, world, per in [(10, 4, 2), (50000, 8, 384), (7, 2, 8), (0, 2, 4)]:
        seen = []
        for r in range(world):
            seen += list(par.contiguous_shard(n, world, r))
        assert sorted(seen) == list(range(n))
        seen = []
        for r in range(world):
            for rg in par.strided_batches(n, per, world, r):
                seen += list(rg)
        assert sorted(seen) == list(range(n))
    # the reference's layout: iteration k, rank r -> [world*n*k + r*n, +n)   (sample_ddp_parallel.py:143-150)
    assert [list(x) for x in par.strided_batches(16, 2, 4, 1)] == [[2, 3], [10, 11]]
    assert list(par.contiguous_shard(10, 4, 3)) == [9]
    assert par.rank_seed(3, 8, 5) == 29


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = (torch.arange(2 * 6 * 1, dtype=torch.int32).view(2, 6, 1) + 1000 * rank)
        out = par.gather_token_grids(local)
        # a weight set as the engines keep it: float payload + a table of process-local device pointers (int64)
        w = [torch.full((3,), float(rank)), torch.full((2,), 7000 + rank, dtype=torch.int64), None,
             torch.full((4,), float(rank)).to(torch.bfloat16)]
        shipped = par.broadcast_tensors(w, src=0)
        assert shipped == 3 * 4 + 4 * 2
        q.put((rank, out.tolist(), w[0].tolist(), w[1].tolist(), w[3].float().tolist()))  # plain lists: no fd-passing
    finally:
        dist.destroy_process_group()


def test_gather_token_grids_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    base = torch.arange(12, dtype=torch.int32).view(2, 6, 1)
    want = torch.cat([base, base + 1000], dim=0)
    for rank, out, w, table, wb in res:
        assert out == want.tolist()   # rank-major, identical on every rank
        assert w == [0.0, 0.0, 0.0] and wb == [0.0] * 4  # broadcast from rank 0
        assert table == [7000 + rank] * 2               # the pointer table stays the receiver's own
