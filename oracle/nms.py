"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of /root/reference/nms.py for the path the
pipeline uses: non-rotated, single-label, no a-priori labels.

Rules fixed in SURVEY 8(c):
  (i)   NMS order = stable descending sort by score, ties -> lower original index first (what
        `torchvision.ops.nms`, the back-end taken at nms.py:151-154, does on CPU);
  (ii)  box j is suppressed iff a kept earlier box i has inter/(area_i+area_j-inter) > thr with every fp32
        operation rounded separately (no FMA contraction);
  (iii) no wall-clock limit (nms.py:162-164 is non-deterministic);
  (iv)  `conf > thr` is strict (nms.py:76,121); class offset cls*max_wh added in fp32 before IoU (nms.py:143,149).

PINNED: tests/golden/nms_*.pt hold outputs of the reference file itself, run through the import shim
in oracle/ref_shim.py by tests/golden/make_nms_golden.py; tests/test_oracle_pins.py checks this restatement (and
the C restatement oracle/nms_oracle.c) against them.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch


def xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    """SURVEY App. A.4 (upstream ultralytics.utils.ops.xywh2xyxy)."""
    y = torch.empty_like(x, dtype=torch.float32)
    xy, wh = x[..., :2], x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def nms_core(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """Greedy NMS with torchvision-CPU semantics, written with numpy fp32 scalars (separately rounded ops).
    Returns int64 kept indices in score order.  O(N*K); use for N up to a few thousand."""
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64)
    b = boxes.detach().cpu().numpy().astype(np.float32)
    s = scores.detach().cpu().numpy().astype(np.float32)
    order = np.argsort(-s, kind="stable")
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    thr = np.float32(iou_threshold)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    zero = np.float32(0)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(zero, xx2 - xx1)
        h = np.maximum(zero, yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > thr]] = True
    return torch.from_numpy(np.asarray(keep, dtype=np.int64))


def non_max_suppression(prediction, conf_thres: float = 0.25, iou_thres: float = 0.45, classes=None,
                        agnostic: bool = False, multi_label: bool = False, labels=(), max_det: int = 300,
                        nc: int = 0, max_time_img: float = 0.05, max_nms: int = 30000, max_wh: int = 7680,
                        rotated: bool = False, end2end: bool = False, return_idxs: bool = False,
                        core=None):
    """Restates nms.py:13-166.  `core(boxes, scores, thr)` defaults to torchvision.ops.nms when importable (the
    branch the reference pipeline takes, evaluate_model.py:24) else `nms_core` above."""
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    assert not rotated and not multi_label and not labels and not end2end, "outside the hot path (SURVEY 2)"
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    if core is None:
        try:
            import torchvision
            core = torchvision.ops.nms
        except Exception:  # pragma: no cover
            core = nms_core
    dev = prediction.device
    bs, ch, na = prediction.shape
    nc = nc or (ch - 4)
    extra = ch - nc - 4
    mi = 4 + nc
    cand = prediction[:, 4:mi].amax(1) > conf_thres
    pred = prediction.transpose(-1, -2)
    pred[..., :4] = xywh2xyxy(pred[..., :4])          # in place on the caller's tensor, like nms.py:84-86
    cls_filter = None if classes is None else torch.tensor(classes, device=dev)
    out = [torch.zeros((0, 6 + extra), device=dev)] * bs
    keepi = [torch.zeros((0, 1), device=dev)] * bs
    all_idx = torch.arange(na, device=dev)
    for b in range(bs):
        sel = cand[b]
        x, xk = pred[b][sel], all_idx[sel]
        if not x.shape[0]:
            continue
        conf, j = x[:, 4:mi].max(1, keepdim=True)
        ok = conf.view(-1) > conf_thres
        x = torch.cat((x[:, :4], conf, j.float(), x[:, mi:]), 1)[ok]
        xk = xk[ok]
        if cls_filter is not None:
            ok = (x[:, 5:6] == cls_filter).any(1)
            x, xk = x[ok], xk[ok]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:
            top = torch.sort(x[:, 4], descending=True, stable=True).indices[:max_nms]
            x, xk = x[top], xk[top]
        off = x[:, 5:6] * (0 if agnostic else max_wh)
        k = core(x[:, :4] + off, x[:, 4], iou_thres)[:max_det]
        out[b] = x[k]
        keepi[b] = xk[k].view(-1)
    return (out, keepi) if return_idxs else out
