"""One pass of the hot path over one synthetic pair (for ncu captures): python tools/one_pair.py [cfg] [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, synthetic as syn

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
W, H, D = syn.CONFIGS[cfg]
il, ir, _ = syn.textured_pair(H, W, D, 1004)
il, ir = torch.from_numpy(il).cuda(), torch.from_numpy(ir).cuda()
packed = eng.pack_weights(syn.glorot_weights(), 5)
ws = torch.empty(eng.match_workspace_bytes(H, W, D, 5), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    eng.match_pair(il, ir, packed, D, 5, workspace=ws)
torch.cuda.synchronize()
print("ok", cfg)
