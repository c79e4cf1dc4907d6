#!/usr/bin/env python
"""Census of MCCNN_SGM_FUSED against the reference-exact mode, per BASELINE config, on the GPU.

For each config: the same inputs (seeded textured pair; features whose best match follows the pair's disparity field, so
that the disparity maps and the bad-pixel rate mean something) go through both modes; reported are
  * cost volume: max |fast - exact| and max relative difference (relative to max(|exact|, 1): costs lie in [-1, 1]);
  * aggregated volumes S: the same two numbers for both sides, in row chunks;
  * raw WTA maps: pixels whose disparity differs, and for those the gap, in the EXACT volume, between the cost of the disparity
    the fused mode picked and the exact minimum, in units of fp32 ulps of the minimum and relative: a flip at a gap of a few
    ulps is a tie that either rounding order may break either way;
  * final (L-R checked, filled, median-filtered) left map: pixels that differ, max |difference|;
  * bad-2.0 rate of both final maps against the synthetic ground truth (error_calculate.py's rule, |d - gt| > 1 on valid GT).
Usage: python tools/fused_census.py [c1 c2 c5 c3 c4] [--out profiles/r02_fused_census.json]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, synthetic as syn


def ulp32(x):
    x = x.abs().clamp(min=1e-30)
    return torch.pow(2.0, torch.floor(torch.log2(x)) - 23)


def census(cfg):
    W, H, D = syn.CONFIGS[cfg]
    seed = 4000 + int(cfg[1])
    il, ir, gt = syn.textured_pair(H, W, D, seed)
    if cfg == "c4":   # numpy feature synthesis takes minutes at this size: noise-free correlated features on the device
        g = torch.Generator(device="cuda"); g.manual_seed(seed)
        base = torch.randn((H, W, 64), device="cuda", generator=g)
        base = torch.nn.functional.avg_pool2d(base.permute(2, 0, 1)[None], 3, 1, 1)[0].permute(1, 2, 0)
        xs = (torch.arange(W, device="cuda")[None, :] - torch.from_numpy(gt).cuda().long()).clamp(0, W - 1)
        fl = torch.gather(base, 1, xs[..., None].expand(H, W, 64)) + 0.35 * base.std() * torch.randn((H, W, 64), device="cuda", generator=g)
        fr = base + 0.05 * torch.randn((H, W, 64), device="cuda", generator=g)
        fl = torch.nn.functional.normalize(fl, dim=-1).contiguous()
        fr = torch.nn.functional.normalize(fr, dim=-1).contiguous()
    else:
        fl, fr, _ = syn.correlated_features(H, W, D, 64, seed)
        fl, fr = torch.from_numpy(fl).cuda(), torch.from_numpy(fr).cuda()
    il_d, ir_d = torch.from_numpy(il).cuda(), torch.from_numpy(ir).cuda()
    out = {"config": cfg, "W": W, "H": H, "D": D, "pixels": H * W}

    CL, CR = eng.cost_volume(fl, fr, D)
    FL, FR = eng.cost_volume_fast(fl, fr, D)
    chunk = max(1, (1 << 27) // (W * D))

    def vol_diff(A, B):
        mx, mr = 0.0, 0.0
        for y0 in range(0, H, chunk):
            a, b = A[y0:y0 + chunk, :, :D], B[y0:y0 + chunk, :, :D]
            d = (a - b).abs()
            mx = max(mx, float(d.max()))
            mr = max(mr, float((d / a.abs().clamp(min=1.0)).max()))
        return mx, mr

    out["cost_volume_max_abs"], out["cost_volume_max_rel"] = vol_diff(CL, FL)
    SL, SR, dl, dr = eng.sgm(CL, CR, il_d, ir_d, D, keep_volumes=True)
    del CL, CR
    TL, TR, fdl, fdr = eng.sgm(FL, FR, il_d, ir_d, D, keep_volumes=True, mode="fused")
    del FL, FR
    torch.cuda.empty_cache()
    for name, S, T, d, fd in (("left", SL, TL, dl, fdl), ("right", SR, TR, dr, fdr)):
        mx, mr = vol_diff(S, T)
        differ = d != fd
        n = int(differ.sum())
        rec = {"S_max_abs": mx, "S_max_rel": mr, "wta_pixels_differ": n, "wta_fraction_differ": n / (H * W)}
        if n:
            ys, xs = differ.nonzero(as_tuple=True)
            best = S[ys, xs, :D].min(dim=-1).values
            at = S[ys, xs, fd[ys, xs].long()]
            gap = (at - best)
            gap_ulp = gap / ulp32(best)
            rec.update({"gap_ulps_max": float(gap_ulp.max()), "gap_ulps_median": float(gap_ulp.median()),
                        "gap_rel_max": float((gap / best.abs().clamp(min=1.0)).max()),
                        "within_4_ulps": int((gap_ulp <= 4).sum()), "within_16_ulps": int((gap_ulp <= 16).sum()),
                        "within_64_ulps": int((gap_ulp <= 64).sum()),
                        "disparity_jump_max": float((d[ys, xs] - fd[ys, xs]).abs().max())})
        out[name] = rec
    del SL, SR, TL, TR
    torch.cuda.empty_cache()
    fin_e, _ = eng.disparity_pipeline(il_d, ir_d, fl, fr, D)
    fin_e = fin_e.clone()
    fin_f, _ = eng.disparity_pipeline(il_d, ir_d, fl, fr, D, mode="fused")
    dfin = fin_e != fin_f
    out["final_pixels_differ"] = int(dfin.sum())
    out["final_fraction_differ"] = float(dfin.float().mean())
    out["final_max_abs_diff"] = float((fin_e - fin_f).abs().max())
    g = torch.from_numpy(gt).cuda()
    valid = torch.isfinite(g) & (g != 0)

    def bad(m):
        return float((valid & ((m.floor() - g).abs() > 1)).sum()) / (H * W)   # astype(int) then |d - gt| > 1 (error_calculate.py:68-83)

    out["bad_rate_exact"], out["bad_rate_fused"] = bad(fin_e), bad(fin_f)
    out["bad_rate_delta"] = out["bad_rate_fused"] - out["bad_rate_exact"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c1", "c2", "c5", "c3"])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = []
    for cfg in a.configs:
        r = census(cfg)
        res.append(r)
        print(json.dumps(r), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
