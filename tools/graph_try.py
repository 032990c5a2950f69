import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_u_b200 as ysp
from yolo_u_b200.synth import synth_state_dicts
B = 256
det_sd, seg_sd = synth_state_dicts(0)
P = ysp.Predictor(det_sd, seg_sd, mode="bf16")
g = torch.Generator().manual_seed(1)
xs = [torch.rand(B, 4, 240, 240, generator=g).cuda() for _ in range(3)]
tg = (torch.rand(B, 1, 240, 240, generator=g) > 0.5).float().cuda()
xin = xs[0].clone()
for _ in range(3):
    P.predict_raw(xin, tg)
torch.cuda.synchronize()
def timeit(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager ms/step", timeit(lambda i: P.predict_raw(xs[i % 3], tg)))
ref = {k: v.clone() for k, v in P.predict_raw(xin, tg).items()}
try:
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        P.predict_raw(xin, tg)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(gr):
        out = P.predict_raw(xin, tg)
    def run(i):
        xin.copy_(xs[i % 3]); gr.replay()
    print("graph ms/step (incl. input copy)", timeit(run))
    print("graph ms/step (no copy)", timeit(lambda i: gr.replay()))
    xin.copy_(xs[0]); gr.replay(); torch.cuda.synchronize()
    print("same result:", all(torch.equal(out[k], ref[k]) for k in ("counts", "det_count", "mask_logits")))
except Exception as e:
    print("graph capture failed:", repr(e)[:400])
