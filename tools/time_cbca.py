"""Device timing of one cross-based aggregation iteration (developer tool): python tools/time_cbca.py [cfg ...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, synthetic as syn, _lib

for cfg in (sys.argv[1:] or ["c4"]):
    W, H, D = syn.CONFIGS[cfg]
    Dp = eng.disp_pitch(D)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    CL = torch.rand((H, W, Dp), device="cuda", generator=g) * 2 - 1
    out, tmp = torch.empty_like(CL), torch.empty_like(CL)
    il, ir, _ = syn.textured_pair(H, W, D, 3)
    il, ir = torch.from_numpy(il).cuda(), torch.from_numpy(ir).cuda()
    al, ar = eng.cross_arms(il), eng.cross_arms(ir)
    lib = _lib.load()
    print(cfg, "mean arm length", al.float().mean().item())
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.mccnn_cbca(CL.data_ptr(), out.data_ptr(), tmp.data_ptr(), al.data_ptr(), ar.data_ptr(), H, W, D, -1, 14,
                                  torch.cuda.current_stream().cuda_stream))
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
    print(f"{cfg} one CBCA pass of one volume: {t:.2f} ms = {16 * H * W * D / t / 1e6:.0f} GB/s of 16 B/eval")
