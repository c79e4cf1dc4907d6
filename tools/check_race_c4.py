#!/usr/bin/env python
"""Development: run the fused sweeps cumulatively twice each and report where the S volumes differ between identical runs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, synthetic as syn

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W, H, D = syn.CONFIGS[cfg]
il, ir, _ = syn.textured_pair(H, W, D, 1004)
il, ir = torch.from_numpy(il).cuda(), torch.from_numpy(ir).cuda()
packed = eng.pack_weights(syn.glorot_weights(), 5)
fl = eng.conv_tower(eng.standardize_pad(il, 5), packed, 5)
fr = eng.conv_tower(eng.standardize_pad(ir, 5), packed, 5)
CL, CR = eng.cost_volume_fast(fl, fr, D)
del fl, fr
for mask in [int(v) for v in os.environ.get("MASKS", "1,3,7").split(",")]:
    os.environ["MCCNN_FUSED_SWEEPS"] = str(mask)
    ref = None
    for rep in range(reps):
        SL, SR, _, _ = eng.sgm(CL, CR, il, ir, D, keep_volumes=True, mode="fused")
        torch.cuda.synchronize()
        if ref is None:
            ref = (SL, SR)
            continue
        for name, a, b in (("L", ref[0], SL), ("R", ref[1], SR)):
            bad = (a[..., :D] != b[..., :D]).any(dim=-1)
            n = int(bad.sum())
            msg = f"{cfg} sweeps mask {mask} rep {rep} side {name}: pixels whose S row differs: {n}"
            if n:
                ys, xs = torch.nonzero(bad, as_tuple=True)
                pts = sorted(zip(ys.tolist(), xs.tolist()))
                msg += f"  rows {min(ys).item()}..{max(ys).item()} cols {min(xs).item()}..{max(xs).item()} first {pts[:5]} last {pts[-5:]}"
            print(msg, flush=True)
        del SL, SR
    del ref
    torch.cuda.empty_cache()
