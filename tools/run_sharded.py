#!/usr/bin/env python
"""torchrun entry: one pair split by rows over the ranks (ShardedMatcher) vs the single-GPU path; parity + timing.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_sharded.py [cfg] [fused] [fault]"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, sharded, synthetic as syn

def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, D = syn.CONFIGS[cfg]
    il, ir, _ = syn.textured_pair(H, W, D, 77)
    weights = syn.glorot_weights()
    opts = sys.argv[2:]
    mode = "fused" if "fused" in opts else "exact"
    m = sharded.ShardedMatcher(H, W, D, weights, mode=mode)
    r0, n = m.row0, m.rows
    bl, br = torch.from_numpy(il[r0:r0 + n]).cuda(), torch.from_numpy(ir[r0:r0 + n]).cuda()
    if "fault" in opts:
        # one rank hands in a band of the wrong shape: EVERY rank must raise for this pair (nobody waits in a kernel), and the
        # next, well-formed pair must work again
        raised = False
        try:
            m.match(bl[:-1] if rank == world - 1 else bl, br)
        except (ValueError, RuntimeError) as exc:
            raised = True
            print(f"[rank {rank}] abandoned pair raised: {type(exc).__name__}: {str(exc)[:90]}", flush=True)
        flag = torch.tensor([1 if raised else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            print("fault injection: some rank did not raise", flush=True)
            sys.exit(2)
    dl, dr = m.match(bl, br)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        ref_l, ref_r = eng.match_pair(torch.from_numpy(il).cuda(), torch.from_numpy(ir).cuda(), m.packed, D, 5, mode=mode)
        ok = bool(torch.equal(ref_l, dl) and torch.equal(ref_r, dr))
        print(f"[sharded x{world}] {cfg} {W}x{H} D={D} {mode}: bit-identical to single GPU: {ok}", flush=True)
        t = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            eng.match_pair(torch.from_numpy(il).cuda(), torch.from_numpy(ir).cuda(), m.packed, D, 5, mode=mode)
            torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
        print(f"[single] {min(t) * 1e3:.2f} ms", flush=True)
        eng._ws.clear()
        torch.cuda.empty_cache()
    ts = []
    for _ in range(4):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        m.match(bl, br)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    tt = torch.tensor([min(ts[1:])], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"[sharded x{world}] {float(tt) * 1e3:.2f} ms per pair (max over ranks)", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)

main()
