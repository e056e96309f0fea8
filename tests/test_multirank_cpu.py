"""CPU test of the N>1 host logic (world_size 2, gloo): replicas-only sharding -- each rank owns an independent sequence /
set of BA windows, no data-path collective; the only communication is the barrier and the max-over-ranks reduction of the
timing, exactly what bench.py does over NCCL on the GPU box."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard(n_items, world, rank):
    """work-queue hand-off used by the multi-GPU drivers: contiguous, balanced, no overlap"""
    per, rem = divmod(n_items, world)
    start = rank * per + min(rank, rem)
    return range(start, start + per + (1 if rank < rem else 0))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    # every rank renders ITS OWN sequence (seed 10 + rank), like run_ours()
    ring = bench.frame_ring(2, 4, seed=10 + rank)
    mine = list(shard(10, world, rank))
    # per-rank "time" -> MAX over ranks; aggregate value = world * units / max time (bench.py contract)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sums = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(sums, torch.tensor([float(ring.astype(np.int64).sum())], dtype=torch.float64))
    q.put((rank, float(t.item()), mine, [float(s.item()) for s in sums]))
    dist.destroy_process_group()


def test_two_rank_replicas_gloo():
    world, port = 2, 29611
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [2.0, 2.0]                       # max over ranks reached both
    assert sorted(res[0][2] + res[1][2]) == list(range(10))        # shards cover the queue once
    assert res[0][3] == res[1][3] and res[0][3][0] != res[0][3][1]  # independent sequences per rank, gathered consistently


def test_shard_is_balanced():
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 4, 8):
            parts = [list(shard(n, w, r)) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
