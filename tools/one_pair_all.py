"""One pass of every kernel of the path (incl. the optional stages) over one synthetic pair, for an ncu metrics pass:
python tools/one_pair_all.py [cfg]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedepthestimation_b200 import engine as eng, synthetic as syn, process_functional as pf

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
W, H, D = syn.CONFIGS[cfg]
il, ir, _ = syn.textured_pair(H, W, D, 1004)
il, ir = torch.from_numpy(il).cuda(), torch.from_numpy(ir).cuda()
packed = eng.pack_weights(syn.glorot_weights(), 5)
ws = torch.empty(eng.match_workspace_bytes(H, W, D, 5), dtype=torch.uint8, device="cuda")
eng.match_pair(il, ir, packed, D, 5, workspace=ws)                                   # warm-up, reference-default path
torch.cuda.synchronize()
prm = pf.sgm_params(cbca_iters=1, subpixel=1, bilateral=1)
dl, dr = eng.match_pair(il, ir, packed, D, 5, params=prm, workspace=ws)             # every stage once
eng.encode_u8(dl, 1)
torch.cuda.synchronize()
print("ok", cfg)
