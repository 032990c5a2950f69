"""Generate tests/golden/nms_golden.pt by running the REFERENCE's own /root/reference/nms.py (unmodified, through
oracle/ref_shim.py, torchvision.ops.nms branch -- nms.py:151-154) on seeded synthetic predictions.

Run here (the only place /root/reference exists):   python tests/golden/make_nms_golden.py
The fixture stores the seeds/recipes + the reference outputs, NOT the inputs: tests regenerate the inputs with
`make_case` below (torch CPU generators are deterministic across machines), which keeps the fixture small.
Cases mirror SURVEY 8(d) cfg 5 sweeps: distinct scores / rounded (ties) / all-equal / all-below-threshold /
clustered boxes / multi-class offsets / agnostic / max_det truncation / nms.py:181-183 docstring example.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

# name -> dict(seed, B, nc, A, kind, conf, iou, max_det, agnostic)
CASES = {
    "distinct_8400":  dict(seed=1, B=3, nc=1, A=8400, kind="uniform", conf=0.001, iou=0.7, max_det=300, agnostic=False),
    "ties_8400":      dict(seed=2, B=3, nc=1, A=8400, kind="ties", conf=0.001, iou=0.7, max_det=300, agnostic=False),
    "allequal_2100":  dict(seed=3, B=2, nc=1, A=2100, kind="allequal", conf=0.001, iou=0.7, max_det=300, agnostic=False),
    "below_thr":      dict(seed=4, B=2, nc=1, A=1344, kind="below", conf=0.25, iou=0.45, max_det=300, agnostic=False),
    "clustered_1344": dict(seed=5, B=4, nc=1, A=1344, kind="clustered", conf=0.25, iou=0.45, max_det=300, agnostic=False),
    "clustered_nodet_cap": dict(seed=6, B=2, nc=1, A=3000, kind="clustered", conf=0.05, iou=0.45, max_det=3000, agnostic=False),
    "multiclass_3":   dict(seed=7, B=2, nc=3, A=2100, kind="clustered", conf=0.1, iou=0.5, max_det=300, agnostic=False),
    "multiclass_agn": dict(seed=8, B=2, nc=3, A=2100, kind="clustered", conf=0.1, iou=0.5, max_det=300, agnostic=True),
    "maxdet_7":       dict(seed=9, B=2, nc=1, A=525, kind="uniform", conf=0.001, iou=0.7, max_det=7, agnostic=False),
    "ragged_mixed":   dict(seed=10, B=5, nc=1, A=525, kind="ragged", conf=0.25, iou=0.45, max_det=300, agnostic=False),
    "iou_zero":       dict(seed=11, B=2, nc=1, A=525, kind="clustered", conf=0.25, iou=0.0, max_det=300, agnostic=False),
    "iou_one":        dict(seed=12, B=2, nc=1, A=525, kind="clustered", conf=0.25, iou=1.0, max_det=300, agnostic=False),
}


def make_case(seed, B, nc, A, kind, **_):
    """Synthetic `prediction` [B, 4+nc, A] fp32 (xywh px, class scores in [0,1])."""
    g = torch.Generator().manual_seed(seed)
    cxcy = torch.rand(B, 2, A, generator=g) * 640
    wh = torch.rand(B, 2, A, generator=g) * 192 + 2
    cls = torch.rand(B, nc, A, generator=g)
    if kind == "ties":
        cls = (cls * 100).round() / 100
    elif kind == "allequal":
        cls = torch.full_like(cls, 0.5)
    elif kind == "below":
        cls = cls * 0.2
    elif kind in ("clustered", "ragged"):
        k = 12
        centers = torch.rand(B, 2, k, generator=g) * 560 + 40
        which = torch.randint(0, k, (B, A), generator=g)
        cxcy = torch.gather(centers, 2, which[:, None, :].expand(B, 2, A)) + torch.randn(B, 2, A, generator=g) * 6
        base = torch.rand(B, 2, k, generator=g) * 100 + 30
        wh = torch.gather(base, 2, which[:, None, :].expand(B, 2, A)) * (1 + 0.15 * torch.randn(B, 2, A, generator=g)).clamp(0.5, 1.5)
        cls = cls ** 3
        if kind == "ragged":            # image 0: nothing passes; image 1: exactly one; others mixed
            cls[0] *= 0.1
            cls[1] *= 0.1
            cls[1, 0, 17] = 0.9
    return torch.cat([cxcy, wh, cls], 1).contiguous()


def main():
    from oracle.ref_shim import load_reference_nms
    ref = load_reference_nms()
    out = {}
    for name, c in CASES.items():
        pred = make_case(**c)
        dets, keep = ref.non_max_suppression(pred.clone(), c["conf"], c["iou"], agnostic=c["agnostic"],
                                             max_det=c["max_det"], return_idxs=True)
        out[name] = dict(cfg=c, dets=[d.clone() for d in dets],
                         keep=[k.clone().to(torch.int64).view(-1) for k in keep])
        print(name, [int(k.numel()) for k in keep])
    # nms.py:181-183 docstring example, both reference back-ends
    import torchvision
    boxes = torch.tensor([[0, 0, 10, 10], [5, 5, 15, 15]], dtype=torch.float32)
    scores = torch.tensor([0.9, 0.8])
    out["docstring"] = dict(boxes=boxes, scores=scores, thr=0.5,
                            keep_torchnms=ref.TorchNMS.nms(boxes, scores, 0.5),
                            keep_tv=torchvision.ops.nms(boxes, scores, 0.5))
    print("docstring", out["docstring"]["keep_torchnms"].tolist(), out["docstring"]["keep_tv"].tolist())
    # TorchNMS.nms on distinct scores (where the reference's two back-ends agree): core-level golden
    g = torch.Generator().manual_seed(21)
    b = torch.rand(1500, 2, generator=g) * 300
    wh = torch.rand(1500, 2, generator=g) * 80 + 4
    boxes = torch.cat([b, b + wh], 1)
    scores = torch.rand(1500, generator=g)
    out["core_1500"] = dict(seed=21, thr=0.5, keep_torchnms=ref.TorchNMS.nms(boxes, scores, 0.5),
                            keep_tv=torchvision.ops.nms(boxes, scores, 0.5))
    print("core_1500", out["core_1500"]["keep_tv"].numel(),
          torch.equal(out["core_1500"]["keep_tv"], out["core_1500"]["keep_torchnms"]))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nms_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
