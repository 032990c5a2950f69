"""Small target for `ncu --set full`: the bf16 pipeline on B slices, a few steps (no timing reported from here)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_u_b200 as ysp
from yolo_u_b200.synth import synth_state_dicts

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
det_sd, seg_sd = synth_state_dicts(0)
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
P = ysp.Predictor(det_sd, seg_sd, mode=mode)
g = torch.Generator().manual_seed(1)
x = torch.rand(B, 4, 240, 240, generator=g).cuda()
tg = (torch.rand(B, 1, 240, 240, generator=g) > 0.5).float().cuda()
for _ in range(steps):
    o = P.predict_raw(x, tg)
torch.cuda.synchronize()
print("ok", int(o["counts"].sum()))
