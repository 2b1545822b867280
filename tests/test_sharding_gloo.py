"""World-size-2 test (gloo, CPU) of the multi-GPU host logic: contiguous frame shards that
tile the sequence, seam pairs, and the max-over-ranks timing reduction bench.py uses."""
import os
import socket
import sys

import pytest

from conftest import ROOT


def _worker(rank, world, port, nframes, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from multimot_track_b200.sharding import shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(nframes, rank, world)
    bounds = [None] * world
    dist.all_gather_object(bounds, (lo, hi))
    t = torch.tensor([1.0 + rank], dtype=torch.float64)           # pretend device time of this rank
    g = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(g, t)                                          # bench.py: per-rank times, the job takes the slowest rank's
    assert [float(x.item()) for x in g] == [1.0 + r for r in range(world)]
    t = torch.tensor([max(float(x.item()) for x in g)], dtype=torch.float64)
    n = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    dist.barrier()
    if rank == 0:
        q.put((bounds, float(t.item()), int(n.item())))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("nframes", [101, 64, 1])
def test_two_rank_sharding(nframes):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nframes, q)) for r in range(2)]
    for p in procs:
        p.start()
    bounds, tmax, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert bounds[0][0] == 0 and bounds[-1][1] == nframes and bounds[0][1] == bounds[1][0]       # contiguous, tiles the sequence
    assert total == nframes and tmax == 2.0                                                        # slowest rank defines the time


def test_shard_helpers():
    from multimot_track_b200.sharding import job_throughput, seam_pairs, shard_bounds
    for F in (0, 1, 7, 100000):
        for G in (1, 2, 4, 8):
            b = [shard_bounds(F, r, G) for r in range(G)]
            assert b[0][0] == 0 and b[-1][1] == F and all(b[i][1] == b[i + 1][0] for i in range(G - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    assert seam_pairs(100, 4) == [(24, 25), (49, 50), (74, 75)]
    assert seam_pairs(100, 1) == []
    assert job_throughput([32, 32], [0.5, 1.0]) == 64.0
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)
