"""N > 1 host logic on CPU: world_size-2 gloo processes scatter a batch, 'solve' their shard and
gather trajectories; no collective sits between scatter and gather."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rda_planner_b200.sharding import shard_bounds, scatter_batch, gather_batch


def test_shard_bounds_cover_everything():
    for total in (1, 7, 8, 2048, 2049):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    full = None
    if rank == 0:
        g = torch.Generator().manual_seed(0)
        full = {'nom_s': torch.randn(total, 3, 5, generator=g), 'kind': torch.arange(total, dtype=torch.int32).reshape(total, 1)}
    shard, n = scatter_batch(full, total, torch.device('cpu'))
    lo, hi = shard_bounds(total, world, rank)
    assert n == hi - lo and (shard['kind'][:, 0] == torch.arange(lo, hi, dtype=torch.int32)).all()
    local = shard['nom_s'] * 2.0 + shard['kind'][:, :, None].float()     # stand-in for the solve
    out = gather_batch(local, total)
    if rank == 0:
        ref = full['nom_s'] * 2.0 + full['kind'][:, :, None].float()
        q.put(bool(torch.equal(out, ref)))
    dist.destroy_process_group()


@pytest.mark.parametrize('total', [5, 8])
def test_scatter_solve_gather_gloo(total):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
