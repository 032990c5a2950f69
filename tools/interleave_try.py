"""A/B: one engine running full batches back to back vs k engines (own handles, own workspaces) that take alternate FULL
batches on k streams -- does the tail of batch i (decoder stage 4, mask/Dice, NMS) and the latency-bound small-map layers
hide under the other batch's kernels?  Run under gpurun:  python tools/interleave_try.py [mode] [B]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_u_b200 as ysp
from yolo_u_b200.synth import calibrate, synth_state_dicts

mode = sys.argv[1] if len(sys.argv) > 1 else "tc32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
det_sd, seg_sd = calibrate(*synth_state_dicts(0))
g = torch.Generator().manual_seed(1)
xs = [torch.rand(B, 4, 240, 240, generator=g).cuda() for _ in range(4)]
tg = (torch.rand(B, 1, 240, 240, generator=g) > 0.5).float().cuda()
STEPS = 24

for k in (1, 2, 3):
    Ps = [ysp.Predictor(det_sd, seg_sd, mode=mode) for _ in range(k)]
    ss = [torch.cuda.Stream() for _ in range(k)]
    outs = [None] * k

    def run(n):
        cur = torch.cuda.current_stream()
        for s in ss:
            s.wait_stream(cur)
        for i in range(n):
            j = i % k
            with torch.cuda.stream(ss[j]):
                outs[j] = Ps[j].predict_raw(xs[i % 4], tg)
        for s in ss:
            cur.wait_stream(s)

    run(2 * k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(STEPS)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    print(f"{mode} B={B}: {k} engine(s), alternate full batches: {ms:.3f} ms per batch = {B / ms * 1e3:.0f} slices/s", flush=True)
    del Ps
    torch.cuda.empty_cache()
