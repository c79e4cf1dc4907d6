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


def _band_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scenedepthestimation_b200 import sharded, synthetic as syn

    H, W = 11, 24                                       # 11 rows over 2 ranks: bands of 6 and 5 rows
    il, ir, _ = syn.textured_pair(H, W, 8, 77)
    bands = sharded.band_rows(H, world)
    r0, n = bands[rank]
    max_rows = max(m for _, m in bands)
    send = torch.zeros((2, max_rows, W), dtype=torch.uint8)
    send[0, :n] = torch.from_numpy(il[r0:r0 + n])
    send[1, :n] = torch.from_numpy(ir[r0:r0 + n])
    recv = torch.empty((world, 2, max_rows, W), dtype=torch.uint8)
    left, right = torch.empty((H, W), dtype=torch.uint8), torch.empty((H, W), dtype=torch.uint8)
    sharded.gather_bands(send, recv, bands, left, right)                 # ShardedMatcher._gather's body
    whole = bool(np.array_equal(left.numpy(), il) and np.array_equal(right.numpy(), ir))
    # the "every rank will launch" word: one rank with bad inputs makes every rank abandon the pair
    go = torch.tensor([0 if rank == 1 else 1], dtype=torch.int32)
    dist.all_reduce(go, op=dist.ReduceOp.MIN)
    q.put((rank, bands, whole, int(go.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_row_band_gather_and_go_flag_gloo():
    """Host logic of the one-pair-over-N-GPUs path (sharded.py): uneven bands, padded all_gather, the go word's MIN."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_band_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bands, whole, go in got:
        assert bands == [(0, 6), (6, 5)] and whole and go == 0
