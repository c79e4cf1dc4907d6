"""Device timing of the MC-CNN-accurate decision head alone (developer tool): python tools/time_fc_head.py [cfg ...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, synthetic as syn

for cfg in (sys.argv[1:] or ["c3"]):
    W, H, D = syn.CONFIGS[cfg]
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    f = [torch.randn((H, W, 64), device="cuda", generator=g) for _ in range(2)]
    f = [x / x.norm(dim=-1, keepdim=True) for x in f]
    head = eng.FcHeadWeights(syn.glorot_fc_weights(gain=2.5))
    x = torch.arange(W, device="cuda")[:, None]; d = torch.arange(D, device="cuda")[None, :]
    evals = H * int((x >= d).sum().item())           # evaluations with a match inside the other image
    tiles = H * sum(1 for xb in range((W + 127) // 128) for dd in range(D) if xb * 128 + 127 >= dd)
    flop = 2.0 * 2 * 384 * 384 + 2 * 384              # per evaluation: fc2 + fc3 + fc4 (fc1 is per pixel)
    ts = []
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); CL, CR = eng.cost_volume_accurate(f[0], f[1], head, D); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); del CL, CR
    t = min(ts)
    print(f"{cfg} accurate head: {t:8.2f} ms  {evals * flop / (t * 1e-3) / 1e12:7.1f} TFLOP/s useful ({tiles * 128 * flop / (t * 1e-3) / 1e12:7.1f} issued)  "
          f"{t * 1e-3 * 1.965e9 / (tiles / 148):8.0f} clk per 128-row tile  {evals / (t * 1e-3) / 1e9:6.2f} Gevals/s")
