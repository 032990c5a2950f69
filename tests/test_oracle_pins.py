"""CPU: pin the oracle against everything checkable the reference holds for this path:
structural pins printed in the reference sources, and golden outputs of the reference's own nms.py."""
import os
import sys

import pytest
import torch

from oracle import cnms
from oracle import nms as onms
from oracle.model import DetectionModel, build_models, dice_from_counts, mask_counts, synth_inputs

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_nms_golden import CASES, make_case  # noqa: E402


def test_backbone_param_counts():
    # /root/reference/YOLOSegPlusPlus.py:127-135 prints these per-layer counts for the 4-ch YOLOv12n backbone
    want = [608, 4672, 6640, 36992, 26080, 147712, 180864, 295424, 689408]
    det = DetectionModel()
    got = [sum(p.numel() for p in det.model[i].parameters()) for i in range(9)]
    assert got == want


def test_total_params_and_head_size():
    det = DetectionModel()
    n_det = sum(p.numel() for p in det.parameters())
    assert n_det == 2568387 + 16 or n_det == 2568387      # DFL arange conv (16) is a frozen constant
    _, seg = build_models(0)
    head = sum(p.numel() for n, p in seg.named_parameters() if not n.startswith("encoder."))
    assert head == 63764


def test_anchor_count_and_grids_at_160(models):
    pred, _ = models
    y, raws = pred.model(torch.zeros(1, 4, 160, 160))
    assert y.shape == (1, 5, 525)                          # visualize_logits.py:39
    assert [tuple(r.shape[1:]) for r in raws] == [(65, 20, 20), (65, 10, 10), (65, 5, 5)]   # generate_objectmaps.py:92


def test_segpp_shapes_at_240(models):
    _, seg = models
    x, lg, _ = synth_inputs(1, 240)
    with torch.no_grad():
        out = seg(x, lg)
    assert out.shape == (1, 1, 240, 240)
    assert out.min() < 0 < out.max()                       # calibrated: both signs (F14)


def test_nms_docstring_example(golden_nms):
    g = golden_nms["docstring"]                            # nms.py:181-183
    assert g["keep_tv"].tolist() == [0, 1] and g["keep_torchnms"].tolist() == [0, 1]
    assert onms.nms_core(g["boxes"], g["scores"], g["thr"]).tolist() == [0, 1]


def test_nms_core_golden(golden_nms):
    g = golden_nms["core_1500"]
    gen = torch.Generator().manual_seed(g["seed"])
    b = torch.rand(1500, 2, generator=gen) * 300
    wh = torch.rand(1500, 2, generator=gen) * 80 + 4
    boxes = torch.cat([b, b + wh], 1)
    scores = torch.rand(1500, generator=gen)
    assert torch.equal(onms.nms_core(boxes, scores, g["thr"]), g["keep_tv"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_nms_restatements_match_reference_golden(golden_nms, name):
    g = golden_nms[name]
    c = g["cfg"]
    pred = make_case(**c)
    # torch restatement
    dets, keep = onms.non_max_suppression(pred.clone(), c["conf"], c["iou"], agnostic=c["agnostic"], max_det=c["max_det"],
                                          return_idxs=True)
    # plain-C restatement
    cd, ck = cnms.nms_batched(pred, c["conf"], c["iou"], c["max_det"], agnostic=c["agnostic"], nthreads=2)
    for b in range(c["B"]):
        assert torch.equal(keep[b].view(-1).long(), g["keep"][b]), f"torch restatement, image {b}"
        assert torch.equal(ck[b], g["keep"][b]), f"C restatement, image {b}"
        assert torch.equal(cd[b], g["dets"][b].reshape(-1, 6)), f"C restatement boxes, image {b}"
        assert torch.equal(dets[b].reshape(-1, 6), g["dets"][b].reshape(-1, 6))


def test_mask_counts_c_vs_torch():
    g = torch.Generator().manual_seed(3)
    lg = torch.randn(3, 1, 64, 64, generator=g)
    lg[0, 0, 0, :4] = torch.tensor([0.0, 5e-8, -5e-8, 1e-6])      # sigmoid == 0.5 band
    tg = (torch.rand(3, 1, 64, 64, generator=g) > 0.5).float()
    tg[2] = 0
    a = mask_counts(lg, tg)
    b = cnms.mask_counts(lg, tg)
    assert torch.equal(a.int(), b)
    d = dice_from_counts(a)
    assert d.shape == (3,) and 0 <= d.min() and d.max() <= 1
    empty = torch.tensor([[0, 0, 0], [0, 5, 0], [3, 4, 6]])
    assert dice_from_counts(empty).tolist() == pytest.approx([1.0, 0.0, 0.6])
