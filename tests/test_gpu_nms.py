"""GPU parity: ysp_nms / ysp_nms_core through the C ABI vs (a) golden outputs of the reference's own nms.py and
(b) the plain-C oracle on seeded inputs.  Bar: keep indices bit-exact (torch.equal), boxes bit-exact."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_nms_golden import CASES, make_case  # noqa: E402


@pytest.fixture(scope="module")
def ysp():
    import yolo_u_b200
    assert torch.cuda.is_available()
    return yolo_u_b200


@pytest.mark.parametrize("name", sorted(CASES))
def test_nms_matches_reference_golden(ysp, golden_nms, name):
    g = golden_nms[name]
    c = g["cfg"]
    pred = make_case(**c).cuda()
    orig = pred.clone()
    dets, keep = ysp.non_max_suppression(pred, c["conf"], c["iou"], agnostic=c["agnostic"], max_det=c["max_det"],
                                         return_idxs=True)
    for b in range(c["B"]):
        assert torch.equal(keep[b].view(-1).long().cpu(), g["keep"][b]), f"{name}: keep idx, image {b}"
        assert torch.equal(dets[b].cpu().reshape(-1, 6), g["dets"][b].reshape(-1, 6)), f"{name}: boxes, image {b}"
    # nms.py:84-86 side effect
    from oracle.nms import xywh2xyxy
    want = xywh2xyxy(orig[:, :4].transpose(1, 2).cpu()).transpose(1, 2)
    assert torch.equal(pred[:, :4].cpu(), want)
    assert torch.equal(pred[:, 4:].cpu(), orig[:, 4:].cpu())


def test_nms_docstring_and_core_golden(ysp, golden_nms):
    g = golden_nms["docstring"]
    assert ysp.TorchNMS.nms(g["boxes"].cuda(), g["scores"].cuda(), g["thr"]).tolist() == [0, 1]
    g = golden_nms["core_1500"]
    gen = torch.Generator().manual_seed(g["seed"])
    b = torch.rand(1500, 2, generator=gen) * 300
    wh = torch.rand(1500, 2, generator=gen) * 80 + 4
    boxes = torch.cat([b, b + wh], 1)
    scores = torch.rand(1500, generator=gen)
    keep = ysp.TorchNMS.nms(boxes.cuda(), scores.cuda(), g["thr"])
    assert keep.dtype == torch.int64 and torch.equal(keep.cpu(), g["keep_tv"])
    idxs = torch.randint(0, 3, (1500,), generator=gen)
    import torchvision
    want = torchvision.ops.batched_nms(boxes, scores, idxs, 0.5)
    got = ysp.TorchNMS.batched_nms(boxes.cuda(), scores.cuda(), idxs.cuda(), 0.5)
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("kind,A,B,conf,iou", [("uniform", 8400, 24, 0.001, 0.7), ("ties", 8400, 16, 0.001, 0.7),
                                               ("allequal", 8400, 4, 0.001, 0.7), ("clustered", 8400, 16, 0.001, 0.7),
                                               ("clustered", 1344, 64, 0.25, 0.45), ("below", 1344, 8, 0.25, 0.45),
                                               ("uniform", 20000, 2, 0.001, 0.7)])
def test_nms_vs_c_oracle_sweep(ysp, kind, A, B, conf, iou):
    """BASELINE cfg 5 sweeps (a subset of the 1024 images per sweep so the CPU oracle finishes in seconds)."""
    from oracle import cnms
    pred = make_case(seed=100 + A % 97 + B, B=B, nc=1, A=A, kind=kind)
    want_d, want_k = cnms.nms_batched(pred, conf, iou, 300, nthreads=8)
    dets, keep = ysp.non_max_suppression(pred.cuda(), conf, iou, return_idxs=True)
    for b in range(B):
        assert torch.equal(keep[b].view(-1).long().cpu(), want_k[b]), f"image {b}"
        assert torch.equal(dets[b].cpu().reshape(-1, 6), want_d[b].reshape(-1, 6))


@pytest.mark.slow
@pytest.mark.timeout(900)
@pytest.mark.parametrize("kind", ["uniform", "ties", "allequal", "below", "clustered"])
def test_nms_cfg5_full_sweep(ysp, kind):
    """BASELINE cfg 5 in full: ALL 1024 images x 8400 anchors per sweep (distinct scores / ties / all-equal / all below
    threshold / clustered boxes), conf 0.001, IoU 0.7, max_det 300 -- every image's keep-index tensor and boxes bit-exact
    against the plain-C oracle (all host cores)."""
    from oracle import cnms
    pred = make_case(seed=5000 + len(kind), B=1024, nc=1, A=8400, kind=kind)
    conf = 0.25 if kind == "below" else 0.001          # "below": every score is < 0.2, nothing may pass the confidence filter
    want_d, want_k = cnms.nms_batched(pred, conf, 0.7, 300, nthreads=os.cpu_count() or 8)
    dets, keep = ysp.non_max_suppression(pred.cuda(), conf, 0.7, return_idxs=True)
    bad = [b for b in range(1024) if not torch.equal(keep[b].view(-1).long().cpu(), want_k[b])
           or not torch.equal(dets[b].cpu().reshape(-1, 6), want_d[b].reshape(-1, 6))]
    assert not bad, f"{kind}: {len(bad)} of 1024 images differ, first {bad[:5]}"
    if kind == "below":
        assert all(k.numel() == 0 for k in keep)
    else:
        assert sum(k.numel() for k in keep) > 1024


def test_nms_options(ysp):
    from oracle import nms as onms
    pred = make_case(seed=77, B=3, nc=3, A=2100, kind="clustered")
    for kw in (dict(classes=[0, 2]), dict(max_nms=150), dict(agnostic=True), dict(max_det=5), dict(nc=2)):
        want_d, want_k = onms.non_max_suppression(pred.clone(), 0.1, 0.5, return_idxs=True, **kw)
        got_d, got_k = ysp.non_max_suppression(pred.clone().cuda(), 0.1, 0.5, return_idxs=True, **kw)
        for b in range(3):
            assert torch.equal(got_k[b].view(-1).long().cpu(), want_k[b].view(-1).long()), (kw, b)
            assert torch.equal(got_d[b].cpu().reshape(want_d[b].shape), want_d[b]), (kw, b)
    # list/tuple input, no idxs, empty batch element types
    out = ysp.non_max_suppression((pred.clone().cuda(), None), 0.999999, 0.5)
    assert isinstance(out, list) and all(o.shape == (0, 6) for o in out)
    out, k = ysp.non_max_suppression(pred.clone().cuda(), 0.999999, 0.5, return_idxs=True)
    assert all(x.shape == (0, 1) for x in k)
    # end2end shortcut nms.py:66-70
    e2e = torch.rand(2, 300, 6).cuda()
    o = ysp.non_max_suppression(e2e, 0.5)
    assert all(torch.equal(a, p[p[:, 4] > 0.5][:300]) for a, p in zip(o, e2e))


def test_nms_full_cfg5_properties(ysp):
    """BASELINE cfg 5 at full size (1024 x 8400): size-independent properties + spot parity on 8 images."""
    from oracle import cnms
    pred = make_case(seed=5, B=1024, nc=1, A=8400, kind="uniform")
    dets, keep = ysp.non_max_suppression(pred.clone().cuda(), 0.001, 0.7, return_idxs=True)
    assert len(dets) == 1024
    for b in range(0, 1024, 37):
        d, k = dets[b], keep[b].view(-1)
        assert d.shape[0] == k.shape[0] <= 300
        assert (d[:-1, 4] >= d[1:, 4]).all()                         # score-descending
        assert k.unique().numel() == k.numel()                       # no duplicates
        assert torch.equal(d[:, 4].cpu(), pred[b, 4, k.cpu()])       # indices point at their scores
        # idempotence: re-running NMS on the survivors keeps all of them
        sub = pred[b:b + 1, :, k.cpu()].clone().cuda()
        _, k2 = ysp.non_max_suppression(sub, 0.001, 0.7, return_idxs=True)
        assert k2[0].numel() == k.numel()
    idx = [0, 1, 511, 1023]
    _, want_k = cnms.nms_batched(pred[idx], 0.001, 0.7, 300, nthreads=4)
    for j, b in enumerate(idx):
        assert torch.equal(keep[b].view(-1).cpu(), want_k[j])
