#!/usr/bin/env python
"""Benchmark of the MC-CNN stereo hot path (BASELINE.json metric: full-resolution Middlebury-2014-shaped
pairs/sec and Gdisp-evals/s on B200, next to the reference's CPU path on the box's own host cores).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference, host cores

A step = one pass of the whole hot path (standardise, conv tower x2, cost volume, 8-path SGM, WTA, L-R
check + fill, median) over one synthetic stereo pair of config c4 (2880x1988, 800 disparities,
random-init weights). With N > 1 every rank processes its own pair (pairs are independent units: weak
scaling, no data-path collective). Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

def kernels_per_step(D: int, mode: str = "exact") -> int:
    """standardise (2x2) + conv (2x5) + cost volume + SGM passes + L-R flags, fill (2), median. The cost volume is one band-GEMM
    launch, or for D >= 512 the tensor-core variant: 2 slice kernels, fill, main kernel, fix-up (pipeline.cu). Fused mode: one
    cost-volume launch and 4 SGM sweeps."""
    if mode == "fused":
        return 2 * 2 + 2 * 5 + 1 + 4 + 4
    return 2 * 2 + 2 * 5 + (5 if D >= 512 else 1) + 7 + 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None, help="c1..c5 (SURVEY.md App. B); the metric is quoted on c4 (default; c3 with --arch accurate)")
    ap.add_argument("--arch", default="fast", choices=["fast", "accurate"], help="matching cost: MC-CNN-fast dot product (the reference's "
                    "net, the headline) or the MC-CNN-accurate fully-connected decision head (BASELINE config 3; own oracle)")
    ap.add_argument("--mode", default="exact", choices=["exact", "fused"], help="exact (default, the headline): the reference's arithmetic "
                    "bit for bit. fused: the opt-in throughput mode (fp32 SGM state, 8 paths in 4 sweeps, fp32-accumulated cost volume; "
                    "held to north_star's 1e-4 tolerance, census in profiles/)")
    ap.add_argument("--batch", type=int, default=0, help="pairs per step per GPU (default 1; 32 for c5 = 256 pairs over 8 ranks), "
                    "kept in flight on --depth CUDA streams like match.py's streamed loop")
    ap.add_argument("--depth", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.config is None:
        a.config = "c3" if a.arch == "accurate" else "c4"
    return a


# ----------------------------------------------------------------------------------------- helpers
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_pair(cfg: str, seed: int):
    from scenedepthestimation_b200 import synthetic as syn

    W, H, D = syn.CONFIGS[cfg]
    il, ir, gt = syn.textured_pair(H, W, D, seed)
    return il, ir, gt, (W, H, D)


# ----------------------------------------------------------------------------------------- CPU legs
def cpu_hot_path_band(il, ir, weights, D, rows, threads, head=None):
    """The oracle (CPU restatement of the reference) on a horizontal band of `rows` rows of the pair:
    conv tower (torch CPU fp32, standing in for TensorFlow) + cost volume + SGM + WTA + L-R + median."""
    import torch

    from oracle import conv_tower as ct
    from oracle import stereo as st
    from scenedepthestimation_b200 import synthetic as syn

    torch.set_num_threads(threads)
    H = il.shape[0]
    r0 = max(0, (H - rows) // 2)
    bl, br = np.ascontiguousarray(il[r0:r0 + rows]), np.ascontiguousarray(ir[r0:r0 + rows])
    t0 = time.perf_counter()
    fl, fr = ct.compute_feature(syn.standardise(bl), syn.standardise(br), 11, 11, 64, weights)
    if head is None:
        st.disparity_pipeline(bl, br, fl, fr, D)
    else:  # MC-CNN-accurate: the decision head (numpy fp32) as the matching cost, then the same post-processing chain
        from oracle import fc_head as fh

        cl, cr = fh.head_cost_volume(fl, fr, head, D)
        sl, sr = st.sgm_all_paths(cl, cr, st.sgm_penalties(bl), st.sgm_penalties(br))
        wl, wr = st.wta(sl), st.wta(sr)
        st.median5(st.lrc_fill(wl, st.lr_flags(wl, wr)[0]), wl)
    return time.perf_counter() - t0


def cpu_sample(il, ir, weights, D, target_s, threads, head=None):
    """Pick a band height that costs about target_s seconds, run it, return (pairs/s equivalent, description)."""
    import numba

    numba.set_num_threads(threads)
    H, W = il.shape
    cpu_hot_path_band(il[:, :64], ir[:, :64], weights, min(D, 16), 4, threads, head)  # JIT warm-up, not timed
    probe_rows = 4
    t = cpu_hot_path_band(il, ir, weights, D, probe_rows, threads, head)
    rows = int(max(probe_rows, min(H, probe_rows * target_s / max(t, 1e-3))))
    t = cpu_hot_path_band(il, ir, weights, D, rows, threads, head)
    pairs_per_s = (rows / H) / t
    return pairs_per_s, t, f"{rows} of {H} rows x {W} px x {D} disparities of the same pair (whole hot path), {t:.1f} s"


# ----------------------------------------------------------------------------------------- main
_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line: libraries write banners to fd 1 (NCCL prints its version there whatever
    NCCL_DEBUG_FILE says), so fd 1 is pointed at stderr for the whole run and the line goes to a duplicate of the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: str):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def main():
    a = parse()
    claim_stdout()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    threads = len(os.sched_getaffinity(0))
    from scenedepthestimation_b200 import synthetic as syn

    W, H, D = syn.CONFIGS[a.config]
    evals = H * W * D
    batch = a.batch if a.batch > 0 else (32 if a.config == "c5" else 1)
    shape_name = {"c1": "Middlebury-2006 third-size", "c2": "Middlebury-2005/2006 half-size", "c3": "Middlebury-2014 half-size",
                  "c4": "Middlebury-2014-shaped full-res", "c5": "KITTI-shaped"}[a.config]
    net = ("MC-CNN-fast (5x 3x3 conv, 64 maps, random-init)" if a.arch == "fast" else
           "MC-CNN-accurate (5x 3x3 conv, 64 maps + fully-connected decision head 128-384-384-384-1, random-init; own oracle)")
    config = {"workload": f"{a.config}: synthetic {shape_name} pair {W}x{H}, {D} disparities, {net} "
                          "+ 8-path SGM + WTA + L-R check/fill + 5x5 median",
              "pairs_per_step_per_gpu": batch, "parallelism": f"pair-per-rank x{world}" + (f", {a.depth} pairs in flight per GPU" if batch > 1 else ""),
              "l2": f"per-step working set (4 fp32 volumes per pair in flight, {4 * evals * 4 / 1e9:.1f} GB each set) exceeds the 126 MB L2; no flush needed",
              "arithmetic": ("fused mode (opt-in): fp32 SGM path state, 8 contributions added to S in 4 sweeps, fp32-accumulated cost volume; "
                             "within north_star's 1e-4 of the reference-exact mode, disparities equal except at near-ties (profiles/r02_fused_census.json)")
                            if a.mode == "fused" else
                            "reference-exact (fp64 SGM state and cost accumulator, fp32 S rounded per path in reference order)" if a.arch == "fast" else
                            "decision head: fp16 operands, fp32 accumulation on tcgen05 (own oracle, |d cost| <= 2e-3); SGM and post-processing reference-exact"}

    if a.impl == "reference":
        if rank != 0:
            return
        il, ir, _, _ = make_pair(a.config, 1000 + 4)
        weights = syn.glorot_weights()
        head_w = syn.glorot_fc_weights() if a.arch == "accurate" else None
        import numba

        numba.set_num_threads(threads)
        cpu_hot_path_band(il[:, :64], ir[:, :64], weights, min(D, 16), 4, threads, head_w)
        t = cpu_hot_path_band(il, ir, weights, D, 4, threads, head_w)
        per_step_s = max(2.0, min(20.0, 120.0 / max(1, a.steps + a.warmup)))
        rows = int(max(4, min(H, 4 * per_step_s / max(t, 1e-3))))
        for _ in range(a.warmup):
            cpu_hot_path_band(il, ir, weights, D, rows, threads, head_w)
        ts = [cpu_hot_path_band(il, ir, weights, D, rows, threads, head_w) for _ in range(a.steps)]
        tot = float(np.sum(ts))
        v = (rows / H) * a.steps / tot
        sample = f"each step = {rows} of {H} rows x {W} px x {D} disparities (whole hot path), scaled to pairs"
        emit(json.dumps({
            "impl": "reference", "metric": "pairs_per_sec", "value": v, "unit": "pairs/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "gdisp_evals_per_sec": v * evals / 1e9,
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of the reference (njit-parallel oracle; torch CPU conv stands in for TensorFlow); "
                    "the reference has no CPU implementation of SGM/L-R/median and cannot be imported without TensorFlow"}))
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local % torch.cuda.device_count())
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from scenedepthestimation_b200 import engine as eng

    il, ir, _, _ = make_pair(a.config, 1000 + 4 + rank)
    weights = syn.glorot_weights()
    packed = eng.pack_weights(weights, 5)
    head_w = syn.glorot_fc_weights() if a.arch == "accurate" else None
    head = eng.FcHeadWeights(head_w) if head_w is not None else None
    if head is not None and batch > 1:
        raise SystemExit("--arch accurate runs one pair per step")
    d_il, d_ir = torch.from_numpy(il).cuda(), torch.from_numpy(ir).cuda()
    ws = torch.empty((eng.match_accurate_workspace_bytes if head is not None else eng.match_workspace_bytes)(H, W, D, 5),
                     dtype=torch.uint8, device="cuda")
    out = (torch.empty((H, W), device="cuda"), torch.empty((H, W), device="cuda"))
    slots = []
    if batch > 1:  # pair-batch per rank: `depth` pairs in flight, each on its own stream with its own workspace
        for _ in range(a.depth):
            slots.append((torch.cuda.Stream(), torch.empty_like(ws), (torch.empty_like(out[0]), torch.empty_like(out[1]))))

    def step():
        if batch == 1:
            eng.match_pair(d_il, d_ir, packed, D, 5, out=out, workspace=ws, head=head, mode=a.mode)
            return
        cur = torch.cuda.current_stream()
        for st, _, _ in slots:
            st.wait_stream(cur)
        for i in range(batch):
            st, w_, o_ = slots[i % len(slots)]
            with torch.cuda.stream(st):
                eng.match_pair(d_il, d_ir, packed, D, 5, out=o_, workspace=w_, mode=a.mode)
        for st, _, _ in slots:
            cur.wait_stream(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, a.warmup)):
        step()
    sampler = ClockSampler(torch.cuda.current_device())
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    value = world * batch * a.steps / (total_ms / 1e3)

    # ---- end to end through the public host-buffer API: pinned host images in, host disparity out, every step
    h_il, h_ir = torch.from_numpy(il).pin_memory(), torch.from_numpy(ir).pin_memory()
    h_dl, h_dr = torch.empty((H, W), dtype=torch.float32).pin_memory(), torch.empty((H, W), dtype=torch.float32).pin_memory()

    streamed = None
    if batch > 1:
        from scenedepthestimation_b200 import match as match_mod

        del slots[:]
        streamed = match_mod.StreamedMatcher(H, W, weights, ndisp=D, scale=1, depth=a.depth, mode=a.mode)

    def e2e_step():
        if streamed is not None:  # match.py's loop: host u8 pairs in, host u8 disparity maps out, `depth` pairs in flight
            for i in range(batch):
                streamed.submit(il, ir, i)
            streamed.drain()
            return
        dl_, dr_ = eng.match_pair(h_il.cuda(non_blocking=True), h_ir.cuda(non_blocking=True), packed, D, 5, out=out, workspace=ws, head=head, mode=a.mode)
        h_dl.copy_(dl_, non_blocking=True)
        h_dr.copy_(dr_, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller holds the result on the host

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * batch * a.steps / float(e2e_s.item())

    # ---- roofline of the dominant kernel (the SGM scanline kernel, 7 launches per pair), timed alone with
    # CUDA events on the launching stream; algorithmic bytes per pair = 16 * H*W*D (both sides,
    # read C once + write S once)
    from scenedepthestimation_b200 import _lib
    import ctypes as C

    lib = _lib.load()
    stage = np.zeros(7, np.float32)
    reps = max(3, a.steps)
    for _ in range(reps):
        eng.match_pair(d_il, d_ir, packed, D, 5, stage_ms=stage, out=out, workspace=ws, head=head, mode=a.mode)
    stage /= reps
    n_sgm = 4 if a.mode == "fused" else 7
    sgm_launch_ms = float(stage[3]) / n_sgm
    peak, peak_src = load_peaks()
    alg_bytes_launch = 16.0 * evals / n_sgm
    achieved = alg_bytes_launch / (sgm_launch_ms * 1e-3) / 1e9
    roofline = {"kernel": "sgm_chain_kernel x3 + sgm_fused_last_kernel (4 sweeps per pair: down+down-right+up, left+down-left, right+up-right, up-left+WTA)"
                if a.mode == "fused" else "sgm_scan_kernel (7 launches per pair: down+up fused, right, left, 4 diagonals, last + WTA)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None,
                "algorithmic_bytes_per_launch": alg_bytes_launch, "avg_launch_ms": sgm_launch_ms,
                "stage_ms": {"features": float(stage[0]), "cost_volume": float(stage[1]), "sgm": float(stage[3]),
                             "lr_check_fill": float(stage[5]), "median": float(stage[6])}}
    if head is not None:
        # the decision head dominates: tensor-pipe bound; useful FLOP = evaluations with a match inside the other image x
        # (fc2 + fc3 + fc4), fc1 is per pixel and not counted; peak = measured cuBLAS bf16 (MEASURED_PEAKS.json)
        valid_evals = H * sum(min(D, x + 1) for x in range(W))
        flop = valid_evals * (2.0 * 2 * 384 * 384 + 2 * 384)
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        tpeak = 1633.5
        if os.path.exists(pk):
            with open(pk) as f:
                tpeak = float(json.load(f).get("bf16_tflops", tpeak))
        ach = flop / (float(stage[1]) * 1e-3) / 1e12
        roofline = {"kernel": "fc_head_kernel (tcgen05, fp16 operands / fp32 accumulation; 2 fc1 kernels + fill + 1 launch per pair)",
                    "bound": "tensor", "achieved": ach, "peak": tpeak, "peak_source": "measured cuBLAS bf16 burst (MEASURED_PEAKS.json)",
                    "unit": "TFLOP/s", "frac": ach / tpeak, "traffic": None, "algorithmic_flop_per_launch": flop,
                    "avg_launch_ms": float(stage[1]),
                    "stage_ms": {"features": float(stage[0]), "cost_volume": float(stage[1]), "sgm": float(stage[3]),
                                 "lr_check_fill": float(stage[5]), "median": float(stage[6])}}
    tr = os.path.join(ROOT, "profiles", "sgm_traffic.json")
    if os.path.exists(tr) and a.config == "c4" and head is None and a.mode == "exact":  # the ncu capture is of the c4 launch
        with open(tr) as f:
            t = json.load(f)
        roofline["traffic"] = t.get("dram_bytes_per_launch")
        roofline["traffic_source"] = t.get("source")

    # ---- the other stages against THEIR bound, from the same stage timings
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    tpk = 1633.5
    if os.path.exists(pk):
        with open(pk) as f:
            tpk = float(json.load(f).get("bf16_tflops", tpk))
    sm_clk_hz, n_sm = 1.965e9, torch.cuda.get_device_properties(0).multi_processor_count
    conv_flop = 2.0 * H * W * 2 * (9 * 64 + 4 * 9 * 64 * 64)  # both images, useful fp32-equivalent FLOP
    # bytes the SGM launches really move per evaluation and side: the order-exact schedule 8 + 5 x 12 + 8, the fused sweeps 8 + 12 + 12 + 8
    real_sgm_bytes = (40.0 if a.mode == "fused" else 76.0) * 2 * evals
    stage_rooflines = {
        "conv_tower": {"bound": "tensor", "useful_tflops": conv_flop / (float(stage[0]) * 1e-3) / 1e12,
                       "issued_tflops": 3 * conv_flop / (float(stage[0]) * 1e-3) / 1e12, "peak_tflops": tpk,
                       "frac_issued": 3 * conv_flop / (float(stage[0]) * 1e-3) / 1e12 / tpk,
                       "note": "fp32-class accuracy from fp16 MMAs: hi*hi + hi*lo + lo*hi, 3 MMAs per k-step"},
        "sgm_real_traffic": {"bound": "hbm", "bytes_per_pair": real_sgm_bytes, "achieved_gbs": real_sgm_bytes / (float(stage[3]) * 1e-3) / 1e9,
                             "peak_gbs": peak, "frac": real_sgm_bytes / (float(stage[3]) * 1e-3) / 1e9 / peak,
                             "note": "what the 4 sweeps move: 40 B per evaluation and side" if a.mode == "fused" else
                                     "what the 7 launches really move (ncu: no re-reads); the reference's per-path fp32 rounding of S fixes the pass structure"},
    }
    if head is None and a.mode == "fused":
        stage_rooflines["cost_volume"] = {"bound": "hbm (fp32-accumulated band GEMM)", "algorithmic_gbs": (2 * evals * 4 + 2 * H * W * 256) / (float(stage[1]) * 1e-3) / 1e9,
                                          "peak_gbs": peak, "frac": (2 * evals * 4 + 2 * H * W * 256) / (float(stage[1]) * 1e-3) / 1e9 / peak}
    elif head is None:
        prod = 64.0 * evals  # products of the exact cost volume (one value serves both volumes)
        stage_rooflines["cost_volume"] = {"bound": "fp32 pipe (exact fp64-accumulate contract)", "products_per_clk_per_sm":
                                          prod / (float(stage[1]) * 1e-3) / sm_clk_hz / n_sm, "pipe_limit_products_per_clk_per_sm": 128.0 / 3,
                                          "algorithmic_gbs": (2 * evals * 4 + 2 * H * W * 256) / (float(stage[1]) * 1e-3) / 1e9, "peak_gbs": peak,
                                          "note": "3 FP32 lane-slots per product (FMUL, FFMA, FADD of the rounding residuals); tensor-core variant for D >= 512"}
    roofline["stages"] = stage_rooflines

    # ---- N > 1: the same pair ALSO split by rows over all ranks (strong scaling of one pair): NVLink
    # peer-memory hand-off of the SGM path state inside the scan kernels, all_gather of the image / WTA bands
    # ---- the opt-in fused mode on the same pair, reported BESIDE the reference-exact headline (never instead of it): device-
    # resident pairs/s, stage times, and how the result compares with the exact mode's on this very pair
    fused = None
    if a.mode == "exact" and batch == 1 and head is None:
        try:
            ref_l, ref_r = [t.clone() for t in eng.match_pair(d_il, d_ir, packed, D, 5, out=out, workspace=ws)]
            for _ in range(3):
                eng.match_pair(d_il, d_ir, packed, D, 5, out=out, workspace=ws, mode="fused")
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(a.steps):
                eng.match_pair(d_il, d_ir, packed, D, 5, out=out, workspace=ws, mode="fused")
            f1.record()
            barrier()
            fms = torch.tensor([f0.elapsed_time(f1)], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(fms, op=dist.ReduceOp.MAX)
            fstage = np.zeros(7, np.float32)
            for _ in range(reps):
                fl_, fr_ = eng.match_pair(d_il, d_ir, packed, D, 5, stage_ms=fstage, out=out, workspace=ws, mode="fused")
            fstage /= reps
            fused = {"mode": "fused (MCCNN_SGM_FUSED): fp32 SGM state, 8 paths in 4 sweeps, tcgen05 cost volume on an fp16 hi/lo split; "
                             "north_star's 1e-4 contract, not the reference's bits",
                     "pairs_per_sec": world * a.steps / (float(fms.item()) / 1e3), "ms_per_pair": float(fms.item()) / a.steps,
                     "stage_ms": {"features": float(fstage[0]), "cost_volume": float(fstage[1]), "sgm": float(fstage[3]),
                                  "lr_check_fill": float(fstage[5]), "median": float(fstage[6])},
                     "sgm_hbm_frac_on_real_bytes": 40.0 * 2 * evals / (float(fstage[3]) * 1e-3) / 1e9 / peak,
                     "sgm_hbm_frac_on_algorithmic_bytes": 16.0 * evals / (float(fstage[3]) * 1e-3) / 1e9 / peak,
                     "cost_volume_hbm_frac": (2 * evals * 4 + 2 * H * W * 256) / (float(fstage[1]) * 1e-3) / 1e9 / peak,
                     "vs_exact_on_this_pair": {"left_map_pixels_differ": int((fl_ != ref_l).sum()), "right_wta_pixels_differ": int((fr_ != ref_r).sum()),
                                               "pixels": H * W, "left_max_abs_diff": float((fl_ - ref_l).abs().max())},
                     "census": "profiles/r02p_fused_census.json"}
        except Exception as exc:  # the headline line must be printed whatever happens here
            fused = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    def sharded_pair(mode):
        """One pair over all ranks in `mode`: ms per pair (device time, max over ranks) and the result compared, in the run, with
        the single-GPU path of the same mode on rank 0."""
        from scenedepthestimation_b200 import sharded as sh

        eng._ws.clear()
        torch.cuda.empty_cache()
        il0, ir0, _, _ = make_pair(a.config, 1000 + 4)  # every rank cuts its band out of rank 0's pair
        m = sh.ShardedMatcher(H, W, D, weights, mode=mode)
        bl = torch.from_numpy(np.ascontiguousarray(il0[m.row0:m.row0 + m.rows])).cuda()
        br = torch.from_numpy(np.ascontiguousarray(ir0[m.row0:m.row0 + m.rows])).cuda()
        for _ in range(2):
            m.match(bl, br)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(a.steps):
            m.match(bl, br, check=False)   # no host round trip per pair; the status of the last pair is read below
        s1.record()
        barrier()
        m.status()
        sms = torch.tensor([s0.elapsed_time(s1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        # checked, not asserted: rank 0 runs the SAME pair through the single-GPU path and compares both output maps bit
        # for bit with what the sharded run left on every rank (the maps are all-gathered, so each rank holds whole maps)
        sdl, sdr = m.match(bl, br)
        identical, max_diff = torch.ones(1, device="cuda", dtype=torch.int32), torch.zeros(1, device="cuda", dtype=torch.float64)
        if rank == 0:
            ws1 = torch.empty(eng.match_workspace_bytes(H, W, D, 5), dtype=torch.uint8, device="cuda")
            rdl, rdr = eng.match_pair(torch.from_numpy(il0).cuda(), torch.from_numpy(ir0).cuda(), packed, D, 5, workspace=ws1, mode=mode)
            same = torch.equal(sdl.view(torch.int32), rdl.view(torch.int32)) and torch.equal(sdr.view(torch.int32), rdr.view(torch.int32))
            identical.fill_(1 if same else 0)
            max_diff.fill_(max(float((sdl - rdl).abs().max()), float((sdr - rdr).abs().max())))
            del ws1, rdl, rdr
        # every rank holds the same gathered maps: a checksum over ranks must agree with rank 0's
        chk = torch.stack([sdl.double().sum(), sdr.double().sum()])
        chk_max, chk_min = chk.clone(), chk.clone()
        dist.all_reduce(chk_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(chk_min, op=dist.ReduceOp.MIN)
        dist.broadcast(identical, 0)
        dist.broadcast(max_diff, 0)
        del m
        torch.cuda.empty_cache()
        return {"ms_per_pair": float(sms.item()) / a.steps, "pairs_per_sec": a.steps / (float(sms.item()) / 1e3),
                "scaling": "strong",
                "bit_identical": bool(identical.item()) and bool(torch.equal(chk_max, chk_min)),
                "max_abs_diff": float(max_diff.item()),
                "checked_against": f"engine.match_pair (single-GPU path, {mode} mode) on the same pair, rank 0; both output maps compared as raw "
                                   "bits; per-rank checksums of the gathered maps agree"}

    sharded = None
    if world > 1 and batch == 1 and head is None and a.mode == "exact":
        del ws
        try:
            sharded = sharded_pair("exact")
            sharded["partition"] = (f"{world} row bands of one pair; conv/cost volume/horizontal SGM band-local, "
                                    "vertical+diagonal SGM path state handed over NVLink peer memory inside the scan kernel")
        except Exception as exc:  # the one-pair-per-rank line above must be printed whatever happens here
            sharded = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        try:   # the opt-in fused mode on the same bands, beside the exact number (compared with the single-GPU FUSED result)
            fs = sharded_pair("fused")
            fs["partition"] = (f"{world} row bands; the two row sweeps of the fused SGM advance in lock step on all ranks (one fp32 state per "
                               "step streamed to the neighbour over NVLink), the column and diagonal sweeps hand over per scanline")
            sharded["fused_mode"] = fs
        except Exception as exc:
            sharded["fused_mode"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- N > 1: BASELINE config 5 as well -- 256 KITTI-shaped pairs (1242x375, 228 disparities) dealt evenly to the ranks, several
    # pairs in flight per GPU on their own streams (device-resident inputs; exact arithmetic), so that the scaling record carries it
    c5 = None
    if world > 1 and a.config == "c4" and batch == 1 and head is None and a.mode == "exact":
        try:
            W5, H5, D5 = syn.CONFIGS["c5"]
            per_rank = 256 // world
            i5l, i5r, _, _ = make_pair("c5", 1000 + 5 + rank)
            d5l, d5r = torch.from_numpy(i5l).cuda(), torch.from_numpy(i5r).cuda()
            nws5 = eng.match_workspace_bytes(H5, W5, D5, 5)
            slots5 = [(torch.cuda.Stream(), torch.empty(nws5, dtype=torch.uint8, device="cuda"),
                       (torch.empty((H5, W5), device="cuda"), torch.empty((H5, W5), device="cuda"))) for _ in range(a.depth)]

            def step5():
                cur = torch.cuda.current_stream()
                for st, _, _ in slots5:
                    st.wait_stream(cur)
                for i in range(per_rank):
                    st, w_, o_ = slots5[i % len(slots5)]
                    with torch.cuda.stream(st):
                        eng.match_pair(d5l, d5r, packed, D5, 5, out=o_, workspace=w_)
                for st, _, _ in slots5:
                    cur.wait_stream(st)

            step5()
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            step5()
            c1.record()
            barrier()
            cms = torch.tensor([c0.elapsed_time(c1)], device="cuda", dtype=torch.float64)
            dist.all_reduce(cms, op=dist.ReduceOp.MAX)
            c5 = {"workload": f"c5: {per_rank * world} KITTI-shaped pairs {W5}x{H5}, {D5} disparities, {per_rank} per rank, {a.depth} in flight per GPU",
                  "pairs_per_sec": per_rank * world / (float(cms.item()) / 1e3), "ms_total": float(cms.item()), "scaling": "weak"}
            del slots5
        except Exception as exc:
            c5 = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, t, desc = cpu_sample(il, ir, weights, D, a.cpu_seconds, threads, head_w)
        cpu = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": desc}
        if head is None:
            # config 1 (463x370, 80 disparities: the reference's own CPU-runnable case) always, as a WHOLE
            # pair (nothing extrapolated), with this repo's time for the same pair beside it
            try:
                W1, H1, D1 = syn.CONFIGS["c1"]
                i1l, i1r, _, _ = make_pair("c1", 1001)
                t1 = cpu_hot_path_band(i1l, i1r, weights, D1, H1, threads)
                t1 = min(t1, cpu_hot_path_band(i1l, i1r, weights, D1, H1, threads))
                g1l, g1r = torch.from_numpy(i1l).cuda(), torch.from_numpy(i1r).cuda()
                ws1 = torch.empty(eng.match_workspace_bytes(H1, W1, D1, 5), dtype=torch.uint8, device="cuda")
                for _ in range(3):
                    eng.match_pair(g1l, g1r, packed, D1, 5, workspace=ws1)
                torch.cuda.synchronize()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(10):
                    eng.match_pair(g1l, g1r, packed, D1, 5, workspace=ws1)
                g1.record()
                torch.cuda.synchronize()
                cpu["config1_whole_pair"] = {"workload": f"c1: {W1}x{H1}, {D1} disparities, whole pair, whole hot path", "cpu_seconds": t1,
                                             "cpu_pairs_per_sec": 1.0 / t1, "gpu_ms_per_pair": g0.elapsed_time(g1) / 10}
            except Exception as exc:
                cpu["config1_whole_pair"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}

    if rank == 0:
        emit(json.dumps({
            "metric": "pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(3, a.warmup), "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": ("f32" if a.mode == "fused" else "f64") if head is None else "f16 x f16 -> f32 (head), f64 (SGM state)",
            "mode": a.mode, "data": "synthetic", "config": config,
            "gdisp_evals_per_sec": value * evals / 1e9,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": 2 * H * W * batch,
                    "d2h_bytes_per_step": (2 * H * W * 4 if batch == 1 else H * W) * batch,
                    "api": "engine.match_pair (mccnn_match_pair) with pinned host u8 images in, host fp32 maps out" if batch == 1 else
                           "match.StreamedMatcher (match.py's loop): host u8 pairs in, host u8 disparity maps out"},
            "gpu_launches": (kernels_per_step(D, a.mode) + (3 if head is not None else 0) - (4 if head is not None and D >= 512 and a.mode == "exact" else 0)) * a.steps * batch, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "fused_mode": fused, "single_pair_sharded": sharded, "pair_batch_c5": c5}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
