#!/usr/bin/env python
"""Per-sweep timing of MCCNN_SGM_FUSED at a config (development): python tools/time_fused_sweeps.py c4"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, synthetic as syn

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
W, H, D = syn.CONFIGS[cfg]
il, ir, _ = syn.textured_pair(H, W, D, 77)
g = torch.Generator(device="cuda"); g.manual_seed(1)
f = lambda: torch.nn.functional.normalize(torch.randn((H, W, 64), device="cuda", generator=g), dim=-1).contiguous()
fl, fr = f(), f()
il, ir = torch.from_numpy(il).cuda(), torch.from_numpy(ir).cuda()
CL, CR = eng.cost_volume_fast(fl, fr, D)
E = 2.0 * H * W * D
for _ in range(2):   # allocations, module load
    eng.sgm(CL, CR, il, ir, D, keep_volumes=False, mode="fused")
torch.cuda.synchronize()
for dbg in (0,):
    for mask, name, bpe in ((1, "sweep0 down+downright", 8), (2, "sweep1 left+downleft", 12), (4, "sweep2 right+upright", 12), (8, "sweep3 upleft+wta", 8), (15, "all", 40)):
        os.environ["MCCNN_FUSED_SWEEPS"] = str(mask)
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = eng.sgm(CL, CR, il, ir, D, keep_volumes=False, mode="fused"); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{cfg} debug={dbg} {name}: {ms:.2f} ms  -> {bpe * E / ms / 1e6:.0f} GB/s on {bpe} B/eval/side", flush=True)
        del r
