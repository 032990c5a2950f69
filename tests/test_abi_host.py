"""CPU: the C-ABI library loads and exports every symbol include/ysp.h declares; host-side logic (state_dict
compatibility, sharding, metric aggregation, error behaviour without a GPU).  No compute calls here."""
import ctypes
import os
import re
import shutil
import subprocess

import pytest
import torch

import yolo_u_b200 as ysp
from yolo_u_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ysp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ysp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = ysp.lib()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"libysp.so does not export {n}"
    assert set(names) == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with include/ysp.h"
    assert L.ysp_version() >= 100


def test_library_sass_is_blackwell_native():
    """The built library carries the sm_100a instructions the design rests on (B200_PROFILING.md's SASS mnemonics):
    tcgen05 MMAs (UTCHMMA) fed by TMA (UTMALDG) with TMEM loads (LDTM) in the epilogues, cp.async staging (LDGSTS) for the
    halo / decoder tiles, and packed fp32 FMAs (FFMA2) in the CUDA-core kernels."""
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", os.path.join(ROOT, "yolo-u_b200", "libysp.so")], capture_output=True, text=True,
                          check=True).stdout
    assert "sm_100a" in sass
    for mnemonic, least in (("UTCHMMA", 100), ("UTMALDG", 4), ("LDTM", 8), ("UTCBAR", 4), ("LDGSTS", 8), ("FFMA2", 100)):
        assert len(re.findall(r"\b" + mnemonic + r"\b", sass)) >= least, mnemonic
    # per kernel: every conv / decoder / attention kernel of the two tensor-core modes issues tcgen05 MMAs (UTCHMMA) and reads
    # its accumulators from TMEM (LDTM); the parity-mode kernels convert with the saturating F2FP
    per_fn = {}
    for blk in sass.split("Function : ")[1:]:
        per_fn[blk.split("\n", 1)[0].strip()] = blk
    for kern in ("conv_tc32_kernel", "conv_halo32_kernel", "dlc32_kernel", "attention_tc_kernel", "conv_tc_kernel", "conv_halo_kernel",
                 "dlc_tc_kernel"):
        bodies = [b for n, b in per_fn.items() if kern in n]
        assert bodies, f"no SASS for {kern}"
        for b in bodies:
            assert "UTCHMMA" in b and "LDTM" in b, kern
    assert all("F2FP.SATFINITE" in b for n, b in per_fn.items() if "conv_tc32_kernel" in n or "dlc32_kernel" in n)
    assert any("UTMALDG" in b for n, b in per_fn.items() if "dlc32_kernel" in n)            # dlc32's P tiles arrive by TMA


def test_library_has_no_work_skipping_switch():
    """The dlc_tc timing probes (which drop MMAs / arithmetic) are compiled only with -DYSP_PROBES; the shipped library
    must not even contain the name of the environment variable that used to enable them."""
    blob = open(os.path.join(ROOT, "yolo-u_b200", "libysp.so"), "rb").read()
    assert b"YSP_DLC_PROBE" not in blob
    assert b"YSP_PROBES" not in open(os.path.join(ROOT, "yolo-u_b200", "build.py"), "rb").read()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_no_cpu_fallback():
    h = ctypes.c_void_p()
    rc = ysp.lib().ysp_create(ctypes.byref(h), 0, 0)
    assert rc == -2 and b"no CPU fallback" in ysp.lib().ysp_last_error()
    with pytest.raises(ysp.YspError):
        ysp.non_max_suppression(torch.zeros(1, 5, 10))
    with pytest.raises(ysp.YspError):
        ysp.TorchNMS.nms(torch.zeros(2, 4), torch.zeros(2), 0.5)
    with pytest.raises(ysp.YspError):
        ysp.mask_counts(torch.zeros(1, 1, 8, 8), None)
    with pytest.raises(ysp.YspError):
        ysp.Engine("cpu")


def test_nms_argument_checks_match_reference():
    # nms.py:59-60 are asserts and fire before anything touches the device
    with pytest.raises(AssertionError, match="Invalid Confidence threshold"):
        ysp.non_max_suppression(torch.zeros(1, 5, 10), conf_thres=1.5)
    with pytest.raises(AssertionError, match="Invalid IoU"):
        ysp.non_max_suppression(torch.zeros(1, 5, 10), iou_thres=-0.1)
    assert ysp.TorchNMS.nms(torch.zeros(0, 4), torch.zeros(0), 0.5).shape == (0,)


def test_state_dict_is_reference_compatible(models):
    pred, seg = models
    m = ysp.YOLOSegPlusPlus(pred)
    ref_sd, sd = seg.state_dict(), m.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys())
    assert all(ref_sd[k].shape == sd[k].shape for k in sd)
    m.load_state_dict(ref_sd)                                  # evaluate_model.py:243
    assert torch.equal(m.state_dict()["output.weight"], ref_sd["output.weight"])
    assert all(not p.requires_grad for p in m.encoder.parameters())       # YOLOSegPlusPlus.py:151-153
    head = sum(p.numel() for n, p in m.named_parameters() if not n.startswith("encoder."))
    assert head == 63764


def test_sharding_is_a_partition():
    for nv, ws in [(64, 1), (64, 2), (64, 4), (64, 8), (7, 3), (2, 4)]:
        got = [ysp.shard_volumes(nv, ws, r) for r in range(ws)]
        assert got[0][0] == 0 and got[-1][1] == nv
        assert all(got[i][1] == got[i + 1][0] for i in range(ws - 1))
        sizes = [b - a for a, b in got]
        assert max(sizes) - min(sizes) <= 1
    assert ysp.shard_slices(64, 155, 8, 3) == (3 * 8 * 155, 4 * 8 * 155)
    assert ysp.batches(0, 10, 4) == [(0, 4), (4, 8), (8, 10)]
    with pytest.raises(ValueError):
        ysp.shard_volumes(4, 2, 2)


def test_dice_and_metrics_match_oracle():
    from oracle.model import dice_from_counts as odice, tp_fp_fn
    g = torch.Generator().manual_seed(0)
    t = torch.randint(0, 50, (16,), generator=g)
    p = torch.randint(0, 50, (16,), generator=g)
    inter = torch.minimum(t, p) // 2
    counts = torch.stack([inter, p, t], 1)
    counts[0] = 0
    counts[1] = torch.tensor([0, 7, 0])
    assert torch.allclose(ysp.dice_from_counts(counts), odice(counts))
    m = ysp.SegMetrics()
    m.update(counts[:8].int())
    m.update(counts[8:].int())
    r = m.compute()
    tp, fp, fn = tp_fp_fn(counts)
    assert (r["TP"], r["FP"], r["FN"]) == (tp, fp, fn)
    assert r["dice"] == pytest.approx(float(odice(counts).double().mean()), abs=1e-7)
    assert r["slices"] == 16


def test_synthetic_checkpoint_matches_reference_layout(models):
    """yolo_u_b200.synth (product-side random checkpoints for the bench) has exactly the reference's parameter names /
    shapes, so the same dict loads into the oracle (strict) and into libysp."""
    from yolo_u_b200.synth import detector_spec, seg_spec, synth_state_dicts
    pred, seg = models
    det_ref, seg_ref = pred.model.model.state_dict(), seg.state_dict()
    assert [k for k, _, _ in detector_spec()] == list(det_ref.keys())
    assert {k: tuple(s) for k, s, _ in detector_spec()} == {k: tuple(v.shape) for k, v in det_ref.items()}
    assert {k: tuple(s) for k, s, _ in seg_spec()} == {k: tuple(v.shape) for k, v in seg_ref.items()}
    d, s = synth_state_dicts(3)
    assert s["encoder.0.conv.weight"] is d["model.0.conv.weight"]          # shared encoder, like YOLOSegPlusPlus.py:150
    from oracle.model import DetectionModel
    DetectionModel().fuse().load_state_dict(d)                             # strict
