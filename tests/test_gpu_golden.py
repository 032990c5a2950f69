"""GPU: the CUDA path against golden outputs of the reference's OWN files (tests/golden/segpp_golden.pt, produced by
tests/golden/make_segpp_golden.py running /root/reference/YOLOSegPlusPlus.py, _YOLOSegPlusPlus.py, dataset.py:89-97 and
evaluate_model.py:157-178 unmodified).  Tolerances: north_star's 1e-3 max-abs on mask logits; integer counters exact."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_segpp_golden import metric_case, objectmaps, seg_inputs  # noqa: E402


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "segpp_golden.pt"), weights_only=False)


@pytest.mark.parametrize("mode", ["fp32", "tc32"])
def test_segpp_forward_vs_reference_file(models, golden, mode):
    import yolo_u_b200 as ysp
    pred, seg = models
    m = ysp.YOLOSegPlusPlus(pred, mode=mode)
    m.load_state_dict(seg.state_dict(), strict=True)
    for (b, h, w), want in zip(golden["shapes"], golden["logits"]):
        x, lg = seg_inputs(b, h, w)
        got = m(x.cuda(), lg.cuda()).cpu()
        err = (got - want).abs().max().item()
        print(f"{mode} {b}x{h}x{w}: max-abs vs reference-file golden {err:.3e}")
        assert err <= 1e-3, (mode, b, h, w, err)


def test_ablation_vs_reference_file(models, golden):
    import yolo_u_b200 as ysp
    from oracle.model import build_models, synth_init_
    from oracle.modules import C3Ghost
    pred, seg = build_models(0)
    seg.decoder[0][0] = C3Ghost(128, 96, n=1)
    synth_init_(seg.decoder[0][0], golden["abl_seed"], lin_gain=2.0)
    x, lg = seg_inputs(2, 96, 96)
    for mode in ("fp32", "tc32"):
        m = ysp.YOLOSegPlusPlus(pred, mode=mode, use_logits=False)
        m.load_state_dict(seg.state_dict(), strict=True)
        assert (m(x.cuda()).cpu() - golden["abl_logits"]).abs().max().item() <= 1e-3, mode


def test_objectmap_transform_vs_reference_lines(golden):
    import yolo_u_b200 as ysp
    for mp, want in zip(objectmaps(), golden["objectmap_out"]):                # dataset.py:89-93,97 executed verbatim
        got = ysp.objectmap_transform(mp.cuda()).cpu().view(want.shape)
        assert (got - want).abs().max().item() <= 1e-6


def test_mask_counters_vs_reference_lines(golden):
    import yolo_u_b200 as ysp
    pred, mask = metric_case()
    c, m = ysp.mask_counts(pred.cuda(), mask.cuda(), want_mask=True)            # evaluate_model.py:157-158,166-168 verbatim
    c = c.cpu()
    for i, (tp, fp, fn) in enumerate(golden["tp_fp_fn"]):
        assert (int(c[i, 0]), int(c[i, 1] - c[i, 0]), int(c[i, 2] - c[i, 0])) == (int(tp), int(fp), int(fn))
    assert torch.equal(m[:1].cpu().float().view_as(golden["pred_binary0"]), golden["pred_binary0"])
    met = ysp.SegMetrics()
    met.update(c)
    r = met.compute()
    assert r["precision"] == pytest.approx(golden["precision_recall"][0], rel=1e-12)      # :177-178
    assert r["recall"] == pytest.approx(golden["precision_recall"][1], rel=1e-12)
