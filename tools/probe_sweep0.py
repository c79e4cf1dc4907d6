#!/usr/bin/env python
"""Development probe: time sequences of fused-SGM calls with different sweep masks (looking for erratic launches)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, synthetic as syn

cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"
seq = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "15,15,1,1,1,1,15,1,2,1,8,1,1").split(",")]
W, H, D = syn.CONFIGS[cfg]
il, ir, _ = syn.textured_pair(H, W, D, 77)
g = torch.Generator(device="cuda"); g.manual_seed(1)
f = lambda: torch.nn.functional.normalize(torch.randn((H, W, 64), device="cuda", generator=g), dim=-1).contiguous()
fl, fr = f(), f()
il, ir = torch.from_numpy(il).cuda(), torch.from_numpy(ir).cuda()
CL, CR = eng.cost_volume_fast(fl, fr, D)
out = []
for mask in seq:
    os.environ["MCCNN_FUSED_SWEEPS"] = str(mask)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = eng.sgm(CL, CR, il, ir, D, keep_volumes=False, mode="fused"); e1.record(); torch.cuda.synchronize()
    out.append(f"{mask}:{e0.elapsed_time(e1):.2f}")
print(cfg, " ".join(out), flush=True)
