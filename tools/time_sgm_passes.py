"""Per-pass device timing of the SGM scan kernels (developer tool): python tools/time_sgm_passes.py [cfg]"""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, synthetic as syn, sharded, _lib

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
W, H, D = syn.CONFIGS[cfg]
Dp = eng.disp_pitch(D)
g = torch.Generator(device="cuda"); g.manual_seed(0)
CL = torch.rand((H, W, Dp), device="cuda", generator=g) * 2 - 1
CR = torch.rand((H, W, Dp), device="cuda", generator=g) * 2 - 1
il = torch.randint(0, 256, (H, W), dtype=torch.uint8, device="cuda", generator=g)
ir = torch.randint(0, 256, (H, W), dtype=torch.uint8, device="cuda", generator=g)
out = (torch.empty_like(CL), torch.empty_like(CR), torch.empty((H, W), device="cuda"), torch.empty((H, W), device="cuda"))
sh = sharded._shard(0, 1, H, 0, H, None, None, None, 1)
names = ["down+up (8B)", "right (12B)", "left (12B)", "down-right (12B)", "up-right (12B)", "down-left (12B)", "up-left+WTA (8B)"]
bytes_per = [8, 12, 12, 12, 12, 12, 8]
for keep in (False,):
    tot = 0.0
    for p in range(7):
        ts = []
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sharded.sgm_band(CL, CR, il, ir, D, sh, pass_mask=1 << p, keep_volumes=keep, out=out)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = min(ts); tot += t
        gb = bytes_per[p] * 2 * H * W * D / 1e9
        print(f"{cfg} pass {p} {names[p]:20s} {t:8.3f} ms  {gb / t:7.1f} GB/s-equiv ({gb / t / 6.5478:5.1f}% of measured peak)")
    print(f"{cfg} total {tot:.2f} ms  env LAST_DEEP={os.environ.get('MCCNN_SGM_LAST_DEEP')}")
