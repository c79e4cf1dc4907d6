#!/usr/bin/env python
"""Development probe for MCCNN_SGM_FUSED: kernels vs oracle.stereo.sgm_all_paths_fused on a list of shapes, then timing at a
BASELINE config. Usage: python tools/try_fused.py [small|c1|c2|c3|c4|c5 ...]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, synthetic as syn


def check(H, W, D, seed=0, kind="tex"):
    from oracle import stereo as st
    if kind == "tex":
        il, ir, _ = syn.textured_pair(H, W, D, seed)
        fl, fr, _ = syn.correlated_features(H, W, D, 64, seed)
    else:
        il, ir = syn.noise_pair(H, W, seed)
        fl, fr = syn.unit_features(H, W, 64, seed)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    CL, CR = eng.cost_volume(dev(fl), dev(fr), D)
    SL, SR, dl, dr = eng.sgm(CL, CR, dev(il), dev(ir), D, keep_volumes=True, mode="fused")
    torch.cuda.synchronize()
    cl, cr = CL[..., :D].cpu().numpy(), CR[..., :D].cpu().numpy()
    esl, esr = st.sgm_all_paths_fused(cl, cr, st.sgm_penalties(il), st.sgm_penalties(ir))
    gl, gr = SL[..., :D].cpu().numpy(), SR[..., :D].cpu().numpy()
    okv = bool(np.array_equal(gl, esl) and np.array_equal(gr, esr))
    okd = bool(np.array_equal(dl.cpu().numpy(), st.wta(esl)) and np.array_equal(dr.cpu().numpy(), st.wta(esr)))
    _, _, dl2, dr2 = eng.sgm(CL, CR, dev(il), dev(ir), D, keep_volumes=False, mode="fused")
    okn = bool(torch.equal(dl, dl2) and torch.equal(dr, dr2))
    print(f"{H}x{W} D={D} {kind}: S == fused oracle: {okv}; WTA: {okd}; no-store variant same maps: {okn}"
          + ("" if okv else f"  maxdiff {np.nanmax(np.abs(gl - esl)):.3g} / {np.nanmax(np.abs(gr - esr)):.3g}, differing {int((gl != esl).sum())}"), flush=True)
    return okv and okd and okn


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
        for (H, W, D, kind) in [(6, 10, 8, "noise"), (20, 48, 32, "tex"), (40, 24, 128, "noise"), (9, 300, 128, "tex"), (33, 65, 1, "tex"),
                                (17, 19, 3, "noise"), (50, 130, 80, "tex"), (64, 40, 228, "noise"), (30, 70, 400, "tex"), (12, 20, 1000, "noise"),
                                (3, 3, 5, "noise"), (100, 9, 33, "noise"), (5, 700, 20, "tex")]:
            ok = check(H, W, D, seed=H + W, kind=kind) and ok
    for cfg in what:
        if cfg in syn.CONFIGS:
            timing(cfg)
    print("ALL OK" if ok else "FAILURES", flush=True)
    sys.exit(0 if ok else 1)
