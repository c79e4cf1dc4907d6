#!/usr/bin/env python
"""torchrun entry (development): one pair split over the ranks by rows, `time_sharded.py [cfg] [exact|fused]`: per-stage times,
ms per pair, and the result compared with the single-GPU path of the same mode."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, sharded, synthetic as syn
cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
mode = sys.argv[2] if len(sys.argv) > 2 else "exact"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, D = syn.CONFIGS[cfg]
il, ir, _ = syn.textured_pair(H, W, D, 1004)
weights = syn.glorot_weights()
m = sharded.ShardedMatcher(H, W, D, weights, mode=mode)
bl, br = torch.from_numpy(il[m.row0:m.row0 + m.rows]).cuda(), torch.from_numpy(ir[m.row0:m.row0 + m.rows]).cuda()
for _ in range(2):
    ref = m.match(bl, br)
ref = [t.clone() for t in ref]
if rank == 0:   # the whole pair on this GPU alone, same mode
    packed = eng.pack_weights(weights, 5)
    one = eng.match_pair(torch.from_numpy(il).cuda(), torch.from_numpy(ir).cuda(), packed, D, mode=mode)
    torch.cuda.synchronize()
    print(f"[x{world}] {cfg} {mode}: equal to the single-GPU result: left {bool(torch.equal(one[0], ref[0]))} right {bool(torch.equal(one[1], ref[1]))}", flush=True)
    del one
    eng._ws.clear()
    torch.cuda.empty_cache()
# per-stage breakdown of one pair (CUDA events around the stages of ShardedMatcher.match, re-enacted here)
def staged():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
    r0, n, nl = m.row0, m.rows, m.nl
    ev[0].record()
    m.u8_send[0, :n].copy_(bl); m.u8_send[1, :n].copy_(br)
    m._gather(m.u8_send, m.u8_recv, m.il, m.ir)
    ev[1].record()
    feats = []
    for img in (m.il, m.ir):
        padded = eng.standardize_pad(img, nl)
        feats.append(eng.conv_tower(padded[r0:r0 + n + 2 * nl], m.packed, nl))
    ev[2].record()
    CL, CR = m._cost_volume(feats[0], feats[1])
    ev[3].record()
    m.go.fill_(1)
    dist.all_reduce(m.go, op=dist.ReduceOp.MIN)
    ev[4].record()
    m.epoch += 1
    shard = sharded._shard(m.rank, m.world, m.H, r0, n, m.xchg.data_ptr(), m.prev, m.next, m.epoch, m.go.data_ptr(), m.timeout_ms)
    out_ = (m.S[0], m.S[1], m.f_send[0, :n], m.f_send[1, :n])
    if m.fused:
        sharded.sgm_fused_band(CL, CR, m.il, m.ir, m.D, shard, m.sgm_ws, keep_volumes=False, out=out_)
    else:
        sharded.sgm_band(CL, CR, m.il, m.ir, m.D, shard, keep_volumes=False, out=out_, ws=m.sgm_ws)
    ev[5].record()
    m._gather(m.f_send, m.f_recv, m.dl, m.dr)
    ev[6].record()
    fl, _ = eng.lr_flags(m.dl, m.dr, right=False)
    eng.median5(eng.lrc_fill(m.dl, fl), m.dl)
    ev[7].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(7)]
for _ in range(2):
    st = staged()
tt = torch.tensor(st, device="cuda"); tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
if rank == 0:
    names = ["gather u8", "standardise + conv", "cost volume", "go all-reduce", "SGM", "gather WTA", "L-R / fill / median"]
    print(f"[x{world}] stages (max over ranks, ms): " + ", ".join(f"{n_} {v:.2f}" for n_, v in zip(names, tmax.tolist())), flush=True)
    print(f"[x{world}] stages (rank 0, ms): " + ", ".join(f"{v:.2f}" for v in st), flush=True)
if m.fused:   # the four sweeps one by one (every launch gets its own epoch: the flags of the previous one must not count)
    feats = []
    for img in (m.il, m.ir):
        padded = eng.standardize_pad(img, m.nl)
        feats.append(eng.conv_tower(padded[m.row0:m.row0 + m.rows + 2 * m.nl], m.packed, m.nl))
    CL, CR = m._cost_volume(feats[0], feats[1])
    res = []
    for rep in range(3):
        row = []
        for sw in range(4):
            m.go.fill_(1)
            dist.all_reduce(m.go, op=dist.ReduceOp.MIN)
            m.epoch += 1
            shard = sharded._shard(m.rank, m.world, m.H, m.row0, m.rows, m.xchg.data_ptr(), m.prev, m.next, m.epoch, m.go.data_ptr(), m.timeout_ms)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sharded.sgm_fused_band(CL, CR, m.il, m.ir, m.D, shard, m.sgm_ws, sweep_mask=1 << sw, keep_volumes=False,
                                   out=(m.S[0], m.S[1], m.f_send[0, :m.rows], m.f_send[1, :m.rows]))
            e1.record(); torch.cuda.synchronize()
            row.append(e0.elapsed_time(e1))
        res = row
    tt = torch.tensor(res, device="cuda"); tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tmin = tt.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"[x{world}] fused sweeps 0..3, ms (max over ranks): " + ", ".join(f"{v:.2f}" for v in tmax.tolist())
              + "   (min over ranks): " + ", ".join(f"{v:.2f}" for v in tmin.tolist()), flush=True)
    del CL, CR, feats
for _ in range(3):
    m.match(bl, br, check=False)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = m.match(bl, br, check=False)
e1.record(); torch.cuda.synchronize(); m.status()
t = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
same = bool(torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1]))
if rank == 0:
    print(f"[x{world}] {cfg} {mode}: {float(t):.2f} ms per pair, deterministic {same}", flush=True)
dist.barrier(); dist.destroy_process_group()
