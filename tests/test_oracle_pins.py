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


# ---- pins against the reference's OWN seg-head file (tests/golden/make_segpp_golden.py ran /root/reference/YOLOSegPlusPlus.py,
# ---- _YOLOSegPlusPlus.py, dataset.py:89-97 and evaluate_model.py:157-178 unmodified through oracle/ref_shim.py) ---------------
@pytest.fixture(scope="module")
def golden_segpp():
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "segpp_golden.pt"), weights_only=False)


def test_segpp_oracle_matches_reference_file(models, golden_segpp):
    from make_segpp_golden import seg_inputs
    _, seg = models
    g = golden_segpp
    assert sorted(seg.state_dict().keys()) == g["state_dict_keys"]            # a reference best.pth loads strictly
    assert g["head_params"] == 63764
    stages = {}
    hooks = [seg.decoder[i].register_forward_hook(lambda m, i_, o, k=i: stages.__setitem__(k, o)) for i in range(5)]
    with torch.no_grad():
        for (b, h, w), want in zip(g["shapes"], g["logits"]):
            x, lg = seg_inputs(b, h, w)
            got = seg(x, lg)
            assert got.shape == want.shape
            # same modules, same op order: equal up to the CPU's conv algorithm choice (oneDNN picks per ISA)
            assert (got - want).abs().max().item() <= 2e-5, (b, h, w)
        x, lg = seg_inputs(*g["shapes"][0])
        seg(x, lg)
    for hk in hooks:
        hk.remove()
    for k, v in stages.items():                                                # topology: concat order, skip pops, per stage
        smp = v.flatten()[:: max(v.numel() // 4096, 1)][:4096]
        assert (smp - g["stage_samples"][k]).abs().max().item() <= 2e-5, f"decoder stage {k}"
        mean, std, amax = g["stage_stats"][k]
        assert abs(float(v.double().mean()) - mean) <= 1e-5 and abs(float(v.abs().max()) - amax) <= 1e-4


def test_ablation_oracle_matches_reference_file(golden_segpp):
    from make_segpp_golden import seg_inputs
    from oracle.model import build_models, synth_init_
    from oracle.modules import C3Ghost
    _, seg = build_models(0)
    seg.decoder[0][0] = C3Ghost(128, 96, n=1)
    synth_init_(seg.decoder[0][0], golden_segpp["abl_seed"], lin_gain=2.0)
    seg.eval()
    x, lg = seg_inputs(2, 96, 96)
    with torch.no_grad():
        skips = []
        h = x
        for idx, m in enumerate(seg.encoder):
            h = m(h)
            if idx in (2, 4):
                skips.append(h)
        for idx, m in enumerate(seg.decoder):
            if idx == 0:
                h = skips.pop()
            elif idx == 2:
                h = torch.cat([h, skips.pop()], 1)
            h = m(h)
        got = seg.output(h)
    assert (got - golden_segpp["abl_logits"]).abs().max().item() <= 2e-5


def test_objectmap_and_metric_lines_match_reference(golden_segpp):
    from make_segpp_golden import metric_case, objectmaps
    from oracle.model import objectmap_transform, tp_fp_fn
    g = golden_segpp
    for mp, want in zip(objectmaps(), g["objectmap_out"]):                     # dataset.py:89-93,97
        assert torch.allclose(objectmap_transform(mp.squeeze(0)), want, atol=1e-7, rtol=0)
    pred, mask = metric_case()
    c = mask_counts(pred, mask)                                                # evaluate_model.py:157-158,166-168
    cc = cnms.mask_counts(pred, mask)
    assert torch.equal(c.int(), cc)
    for i, (tp, fp, fn) in enumerate(g["tp_fp_fn"]):
        assert (int(c[i, 0]), int(c[i, 1] - c[i, 0]), int(c[i, 2] - c[i, 0])) == (int(tp), int(fp), int(fn))
    assert torch.equal((torch.sigmoid(pred[:1]) > 0.5).float(), g["pred_binary0"])
    TP, FP, FN = tp_fp_fn(c)
    assert TP / (TP + FP + 1e-6) == pytest.approx(g["precision_recall"][0], rel=1e-12)    # :177-178
    assert TP / (TP + FN + 1e-6) == pytest.approx(g["precision_recall"][1], rel=1e-12)
    from yolo_u_b200.metrics import SegMetrics                                 # host logic of the product: same formulas
    m = SegMetrics()
    m.update(c)
    r = m.compute()
    assert r["precision"] == pytest.approx(g["precision_recall"][0], rel=1e-12)
    assert r["recall"] == pytest.approx(g["precision_recall"][1], rel=1e-12)
