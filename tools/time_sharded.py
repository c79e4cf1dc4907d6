#!/usr/bin/env python
"""torchrun entry (development): the row-sharded exact pair at c4 under several MCCNN_SGM_STAGGER_NS settings."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, sharded, synthetic as syn
cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, D = syn.CONFIGS[cfg]
il, ir, _ = syn.textured_pair(H, W, D, 1004)
m = sharded.ShardedMatcher(H, W, D, syn.glorot_weights())
bl, br = torch.from_numpy(il[m.row0:m.row0 + m.rows]).cuda(), torch.from_numpy(ir[m.row0:m.row0 + m.rows]).cuda()
for _ in range(2):
    ref = m.match(bl, br)
ref = [t.clone() for t in ref]
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
    sharded.sgm_band(CL, CR, m.il, m.ir, m.D, shard, keep_volumes=False, out=(m.S[0], m.S[1], m.f_send[0, :n], m.f_send[1, :n]), ws=m.sgm_ws)
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
    print(f"[x{world}] {cfg} stagger={os.environ.get('MCCNN_SGM_STAGGER_NS', 'default')}: {float(t):.2f} ms per pair, deterministic {same}", flush=True)
dist.barrier(); dist.destroy_process_group()
