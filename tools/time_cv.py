"""Device timing of the cost-volume kernels alone (developer tool): python tools/time_cv.py [cfg ...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, synthetic as syn

for cfg in (sys.argv[1:] or ["c4"]):
    W, H, D = syn.CONFIGS[cfg]
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    f = [torch.randn((H, W, 64), device="cuda", generator=g) for _ in range(2)]
    f = [x / x.norm(dim=-1, keepdim=True) for x in f]
    E = H * W * D
    for name, fn in (("band GEMM (CUDA cores)", eng.cost_volume), ("tensor-core slices   ", eng.cost_volume_tc)):
        ts = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); CL, CR = fn(f[0], f[1], D); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1)); del CL, CR
        t = min(ts[1:])
        print(f"{cfg} cost volume, {name}: {t:8.3f} ms  {E * 64 / (t * 1e-3) / 148 / 1.965e9:5.1f} useful products/clk/SM  "
              f"{(2 * E * 4 + 2 * H * W * 256) / t / 1e6:6.0f} GB/s algorithmic")
