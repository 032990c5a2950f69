"""Per-layer device times of the pipeline (CUDA events inside libysp) -> markdown table.  Run under gpurun."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_u_b200 as ysp
from yolo_u_b200.synth import calibrate, synth_state_dicts

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
det_sd, seg_sd = synth_state_dicts(0)
det_sd, seg_sd = calibrate(det_sd, seg_sd)
P = ysp.Predictor(det_sd, seg_sd, mode=mode)
g = torch.Generator().manual_seed(1)
x = torch.rand(B, 4, 240, 240, generator=g).cuda()
tg = (torch.rand(B, 1, 240, 240, generator=g) > 0.5).float().cuda()
for _ in range(3):
    P.predict_raw(x, tg)
torch.cuda.synchronize()
P.engine.profile(2)
n = 3
for _ in range(n):
    P.predict_raw(x, tg)
P.engine.profile(0)
rep = P.engine.profile_report()
tot = sum(r["ms"] for r in rep)
print(f"mode {mode} B {B}: sum of per-step device times {tot / n:.3f} ms/step ({len(rep)} steps, {sum(r['launches'] for r in rep) // n} launches)\n")
print("| step | kind | ms/step | share | GB/s (algorithmic) | TFLOP/s |")
print("|---|---|---:|---:|---:|---:|")
for r in sorted(rep, key=lambda r: -r["ms"]):
    ms = r["ms"] / n
    print(f"| {r['name']} | {r['kind']} | {ms:.4f} | {100 * r['ms'] / tot:.1f}% | {r['bytes'] / r['ms'] / 1e6:.0f} | {r['flops'] / r['ms'] / 1e9:.1f} |")
