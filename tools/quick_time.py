"""Quick per-stage device timing of the hot path on synthetic inputs (developer tool)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, synthetic as syn, _lib

def main():
    cfgs = sys.argv[1:] or ["c2"]
    for c in cfgs:
        W, H, D = syn.CONFIGS[c]
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        il = torch.randint(0, 256, (H, W), dtype=torch.uint8, device="cuda", generator=g)
        ir = torch.randint(0, 256, (H, W), dtype=torch.uint8, device="cuda", generator=g)
        packed = eng.pack_weights(syn.glorot_weights(), 5)
        ws = torch.empty(eng.match_workspace_bytes(H, W, D, 5), dtype=torch.uint8, device="cuda")
        out = (torch.empty((H, W), device="cuda"), torch.empty((H, W), device="cuda"))
        for it in range(3):
            st = np.zeros(7, np.float32)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            eng.match_pair(il, ir, packed, D, 5, stage_ms=st, out=out, workspace=ws)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            evals = H * W * D
            print(f"{c} {W}x{H} D={D} it{it}: total {1e3*(t1-t0):.2f} ms  conv {st[0]:.2f} cv {st[1]:.2f} sgm {st[3]:.2f} lrc {st[5]:.3f} med {st[6]:.3f}"
                  f" | sgm {evals*16/st[3]/1e6:.1f} G eval-paths/s, eff BW(76B/eval/side) {evals*2*76/st[3]/1e6:.0f} GB/s", flush=True)
main()
