"""world_size-2 gloo test of the N>1 host logic: pairs are dealt one per rank (no data-path collective),
the bench takes the MAX of the per-rank times, and rank 0 gathers the per-pair results."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scenedepthestimation_b200 import match, synthetic as syn

    ids = match.shard(range(1, 19), rank, world)
    # stand-in for the per-pair result: a checksum of the seeded synthetic pair this rank would process
    sums = torch.zeros(19, dtype=torch.float64)
    for i in ids:
        il, ir, _ = syn.textured_pair(16, 24, 8, 1000 + i)
        sums[i] = float(il.sum()) + float(ir.sum())
    t = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # bench.py: time = max over ranks
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)       # match.py: final gather of per-pair outputs
    if rank == 0:
        q.put((float(t.item()), sums.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_pair_per_rank_sharding_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, sums = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert abs(tmax - 0.2) < 1e-12
    from scenedepthestimation_b200 import synthetic as syn

    for i in range(1, 19):
        il, ir, _ = syn.textured_pair(16, 24, 8, 1000 + i)
        assert sums[i] == float(il.sum()) + float(ir.sum())
