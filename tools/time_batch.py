"""Throughput of a batch of equal-shape pairs: sequential vs streamed (developer tool): python tools/time_batch.py c5 32"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import match, synthetic as syn
cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"; n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
W, H, D = syn.CONFIGS[cfg]
rng = np.random.default_rng(0)
pairs = [(rng.integers(0, 256, (H, W), dtype=np.uint8), rng.integers(0, 256, (H, W), dtype=np.uint8)) for _ in range(4)]
w = syn.glorot_weights()
batch = [pairs[i % 4] for i in range(n)]
match.match_batch(batch[:2], w, ndisp=D)
t0 = time.perf_counter(); match.match_batch(batch, w, ndisp=D); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"{cfg} sequential (pageable copies, sync per pair): {n / (t1 - t0):.1f} pairs/s")
for depth in (1, 2, 3, 4):
    list(match.match_stream(iter(batch[:4]), w, ndisp=D, depth=depth))
    t0 = time.perf_counter(); out = list(match.match_stream(iter(batch), w, ndisp=D, depth=depth)); t1 = time.perf_counter()
    print(f"{cfg} streamed depth={depth}: {n / (t1 - t0):.1f} pairs/s")
