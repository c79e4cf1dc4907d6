#!/usr/bin/env python
"""Development probe for MCCNN_SGM_FUSED: `small` runs the kernels-vs-oracle tests of tests/test_gpu_fused.py, a config name
times both modes at that BASELINE config. Usage: python tools/try_fused.py [small|c1|c2|c3|c4|c5 ...]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, synthetic as syn


def check():
    """The comparison with the CPU oracle lives in tests/ (the oracle is test infrastructure): run it from there."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return subprocess.call([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_fused.py"), "-q", "-x", "-k",
                            "equals_fused_oracle and not full_size"], cwd=root) == 0


def timing(cfg):
    W, H, D = syn.CONFIGS[cfg]
    il, ir, _ = syn.textured_pair(H, W, D, 77)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    f = lambda: torch.nn.functional.normalize(torch.randn((H, W, 64), device="cuda", generator=g), dim=-1).contiguous()
    fl, fr = f(), f()
    il, ir = torch.from_numpy(il).cuda(), torch.from_numpy(ir).cuda()
    lib = eng._lib.load()
    for name, fn in (("cost_volume exact", eng.cost_volume), ("cost_volume fast", eng.cost_volume_fast)):
        for _ in range(2):
            out = fn(fl, fr, D); torch.cuda.synchronize()
            del out
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(fl, fr, D); e1.record(); torch.cuda.synchronize()
        print(f"{cfg} {name}: {e0.elapsed_time(e1):.2f} ms", flush=True)
    CL, CR = out
    ce = eng.cost_volume(fl, fr, D)[0]
    print(f"{cfg} fast vs exact cost volume: max |diff| {float((CL[..., :D] - ce[..., :D]).abs().max()):.3g}", flush=True)
    del ce
    for mode in ("exact", "fused"):
        for keep in (False,):
            for _ in range(2):
                r = eng.sgm(CL, CR, il, ir, D, keep_volumes=keep, mode=mode); torch.cuda.synchronize()
                del r
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = eng.sgm(CL, CR, il, ir, D, keep_volumes=keep, mode=mode); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print(f"{cfg} sgm {mode}: {ms:.2f} ms", flush=True)
            if mode == "exact":
                ref = (r[2].clone(), r[3].clone())
            else:
                dl, dr = r[2], r[3]
                print(f"{cfg} fused vs exact WTA maps: left differs at {int((dl != ref[0]).sum())} of {H * W}, right {int((dr != ref[1]).sum())}", flush=True)
            del r
            torch.cuda.empty_cache()


if __name__ == "__main__":
    what = sys.argv[1:] or ["small"]
    ok = True
    if "small" in what:
        ok = check()
    for cfg in what:
        if cfg in syn.CONFIGS:
            timing(cfg)
    print("ALL OK" if ok else "FAILURES", flush=True)
    sys.exit(0 if ok else 1)
