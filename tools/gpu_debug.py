"""Per-layer comparison of libysp intermediates against oracle forward hooks (run under gpurun)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yolo_u_b200 as ysp
from oracle.model import build_models, synth_inputs, pad_to_multiple

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
pred, seg = build_models(0)
x, lg, tg = synth_inputs(2, 240)
acts = {}
det = pred.model.model
hooks = [det.model[i].register_forward_hook(lambda m, i_, o, k=i: acts.__setitem__(f"det:model.{k}", o)) for i in range(21)]
for i in range(5):
    hooks.append(seg.decoder[i].register_forward_hook(lambda m, i_, o, k=i: acts.__setitem__(f"seg:decoder.{k}", o)))
with torch.no_grad():
    y_ref, raws_ref = pred.model(pad_to_multiple(x))
for h in hooks[:21]:
    h.remove()
with torch.no_grad():
    enc = {}
    hk = [seg.encoder[i].register_forward_hook(lambda m, i_, o, k=i: enc.__setitem__(f"seg:encoder.{k}", o)) for i in range(5)]
    out_ref = seg(x, lg)
acts.update(enc)

eng = ysp.Engine("cuda:0", mode)
eng.keep_intermediates(True)
eng.load_state_dict("det", det.state_dict())
eng.load_state_dict("seg", seg.state_dict())
eng.finalize(True, True)
y, raws = eng.detector_forward(x.cuda())
torch.cuda.synchronize()
for k in sorted((k for k in acts if k.startswith("det:")), key=lambda s: int(s.split(".")[-1])):
    try:
        t = eng.debug_tensor(k).cpu()
    except Exception as e:
        continue
    r = acts[k]
    print(f"{k:18s} shape {tuple(t.shape)} ref {tuple(r.shape)} max-abs {float((t - r).abs().max()):.3e} ref-absmax {float(r.abs().max()):.3e}")
for i, (r, rr) in enumerate(zip(raws, raws_ref)):
    print(f"raw{i} max-abs {float((r.cpu() - rr).abs().max()):.3e}")
print(f"y box max-abs {float((y[:, :4].cpu() - y_ref[:, :4]).abs().max()):.3e}  cls {float((y[:, 4:].cpu() - y_ref[:, 4:]).abs().max()):.3e}")
out = eng.segpp_forward(x.cuda(), lg.cuda())
torch.cuda.synchronize()
for k in sorted(k for k in acts if k.startswith("seg:")):
    try:
        t = eng.debug_tensor(k).cpu()
    except Exception as e:
        print(k, "n/a", e)
        continue
    r = acts[k]
    print(f"{k:18s} shape {tuple(t.shape)} max-abs {float((t - r).abs().max()):.3e} ref-absmax {float(r.abs().max()):.3e}")
print(f"seg out max-abs {float((out.cpu() - out_ref).abs().max()):.3e}  ref std {float(out_ref.std()):.3f}")
