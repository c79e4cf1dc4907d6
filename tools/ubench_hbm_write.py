#!/usr/bin/env python
"""HBM write-only / read-only / copy bandwidth on this GPU (torch kernels; CUDA events; best of 5): the denominators a write-only
kernel (the cost volume) should be compared with."""
import torch
n = 9 * (1 << 30)            # 9 Gi floats = 36 GiB
a = torch.empty(n // 2, dtype=torch.float32, device="cuda")
b = torch.empty(n // 2, dtype=torch.float32, device="cuda")
def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)
for name, fn, nbytes in (("write-only  (fill_)", lambda: a.fill_(1.0), a.numel() * 4),
                         ("write-only  (cudaMemset via zero_)", lambda: a.zero_(), a.numel() * 4),
                         ("read-only   (sum)", lambda: a.sum(), a.numel() * 4),
                         ("copy        (b.copy_(a), read + write bytes)", lambda: b.copy_(a), 2 * a.numel() * 4)):
    fn(); torch.cuda.synchronize()
    ms = best(fn)
    print(f"{name}: {nbytes / 1e9:.1f} GB in {ms:.2f} ms = {nbytes / ms / 1e6:.0f} GB/s", flush=True)
