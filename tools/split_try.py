"""A/B: one engine on B slices vs k engines on B/k slices each, on k streams (do the latency-bound small-map layers of one part
hide under the full-GPU decoder kernels of another?).  Run under gpurun:  python tools/split_try.py [mode] [B]"""
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
xs = [torch.rand(B, 4, 240, 240, generator=g).cuda() for _ in range(3)]
tg = (torch.rand(B, 1, 240, 240, generator=g) > 0.5).float().cuda()


def timeit(fn, steps=20):
    for i in range(3):
        fn(xs[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(xs[i % 3])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


P = ysp.Predictor(det_sd, seg_sd, mode=mode)
ref = {k: v.clone() for k, v in P.predict_raw(xs[0], tg).items()}
print(f"{mode} B={B}: 1 engine  {timeit(lambda x: P.predict_raw(x, tg)):.3f} ms/step")
for k in (2, 3, 4):
    Ps = [ysp.Predictor(det_sd, seg_sd, mode=mode) for _ in range(k)]
    ss = [torch.cuda.Stream() for _ in range(k)]
    cuts = [B * i // k for i in range(k + 1)]

    def run(x):
        cur = torch.cuda.current_stream()
        outs = []
        for i in range(k):
            ss[i].wait_stream(cur)
            with torch.cuda.stream(ss[i]):
                outs.append(Ps[i].predict_raw(x[cuts[i]:cuts[i + 1]], tg[cuts[i]:cuts[i + 1]]))
        for i in range(k):
            cur.wait_stream(ss[i])
        return outs

    outs = run(xs[0])
    torch.cuda.synchronize()
    ml = torch.cat([o["mask_logits"] for o in outs])
    print(f"{mode} B={B}: {k} engines {timeit(run):.3f} ms/step   max |diff| of mask logits vs 1 engine {float((ml - ref['mask_logits']).abs().max()):.2e}")
