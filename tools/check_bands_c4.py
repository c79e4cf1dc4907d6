#!/usr/bin/env python
"""Development check at full size: fused SGM unsharded vs emulated bands (one GPU), and fused vs exact final WTA maps."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, sharded, synthetic as syn

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
worlds = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "2,4").split(",")]
W, H, D = syn.CONFIGS[cfg]
il, ir, _ = syn.textured_pair(H, W, D, 1004)
il, ir = torch.from_numpy(il).cuda(), torch.from_numpy(ir).cuda()
packed = eng.pack_weights(syn.glorot_weights(), 5)
fl = eng.conv_tower(eng.standardize_pad(il, 5), packed, 5)
fr = eng.conv_tower(eng.standardize_pad(ir, 5), packed, 5)
CL, CR = eng.cost_volume_fast(fl, fr, D)
del fl, fr
_, _, dl, dr = eng.sgm(CL, CR, il, ir, D, keep_volumes=False, mode="fused")
_, _, dl2, dr2 = eng.sgm(CL, CR, il, ir, D, keep_volumes=False, mode="fused")
print(f"{cfg}: unsharded fused repeatable: {bool(torch.equal(dl, dl2) and torch.equal(dr, dr2))}", flush=True)
_, _, el, er = eng.sgm(CL, CR, il, ir, D, keep_volumes=False, mode="exact")
print(f"{cfg}: fused vs exact WTA pixels differing: left {int((dl != el).sum())} right {int((dr != er).sum())} of {H * W}", flush=True)
del el, er
torch.cuda.empty_cache()
for world in worlds:
    _, _, bl, br = sharded.emulate_fused_bands(CL, CR, il, ir, D, world, keep_volumes=False)
    torch.cuda.synchronize()
    nl, nr = int((bl != dl).sum()), int((br != dr).sum())
    print(f"{cfg}: {world} emulated bands vs unsharded: differing pixels left {nl} right {nr}", flush=True)
    if nl:
        ys, xs = torch.nonzero(bl != dl, as_tuple=True)
        print("   rows", ys.min().item(), ys.max().item(), "cols", xs.min().item(), xs.max().item(), "first", [(int(y), int(x)) for y, x in zip(ys[:6], xs[:6])], flush=True)
    del bl, br
    torch.cuda.empty_cache()
