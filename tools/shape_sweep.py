import sys; sys.path.insert(0,'/root/repo')
import torch, yolo_u_b200 as ysp
from oracle.model import build_models
pred, seg = build_models(0)
P32 = ysp.Predictor.from_modules(pred, seg, device="cuda:0", mode="fp32")
P16 = ysp.Predictor.from_modules(pred, seg, device="cuda:0", mode="bf16")
g = torch.Generator().manual_seed(3)
for (B,H,W) in [(3,160,160),(2,96,64),(5,200,248),(1,240,240),(7,64,64),(2,136,72),(33,240,240),(2,480,480),(2,72,200)]:
    x = torch.rand(B,4,H,W,generator=g).cuda(); tg=(torch.rand(B,1,H,W,generator=g)>0.5).float().cuda()
    a = P32.predict_raw(x,tg); a={k:v.clone() for k,v in a.items()}
    b = P16.predict_raw(x,tg); torch.cuda.synchronize()
    err=(a["mask_logits"]-b["mask_logits"]).abs().max().item()
    flips=((a["mask_logits"]>0)!=(b["mask_logits"]>0)).float().mean().item()
    std=a["mask_logits"].std().item()
    ok = err <= 0.2*max(std,1.0)+0.1 and flips < 0.03
    print((B,H,W), f"err {err:.3f} std {std:.2f} flips {flips:.4f} dets {a['det_count'].sum().item()} {b['det_count'].sum().item()}", "OK" if ok else "FAIL")
