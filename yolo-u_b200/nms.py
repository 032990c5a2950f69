"""Drop-in for the reference's nms.py (same names, argument meaning and error behaviour) backed by ysp_nms /
ysp_nms_core (csrc/nms.cu).  Mirrors /root/reference/nms.py:13-166 (non_max_suppression) and :239-337 (TorchNMS).

Differences that are deliberate and documented:
  * tie order is the stable one of torchvision.ops.nms (the back-end the reference pipeline takes, nms.py:151-154);
  * no wall-clock limit (nms.py:162-164 silently truncates work; `max_time_img` is accepted and ignored);
  * multi_label / labels / rotated are outside the hot path -> NotImplementedError.
The in-place xywh->xyxy overwrite of the caller's tensor (nms.py:84-86) IS kept.
"""
from __future__ import annotations

import torch

from ._lib import check, lib, require_cuda


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def non_max_suppression(prediction, conf_thres: float = 0.25, iou_thres: float = 0.45, classes=None,
                        agnostic: bool = False, multi_label: bool = False, labels=(), max_det: int = 300, nc: int = 0,
                        max_time_img: float = 0.05, max_nms: int = 30000, max_wh: int = 7680, rotated: bool = False,
                        end2end: bool = False, return_idxs: bool = False):
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    require_cuda(prediction, "non_max_suppression")
    dev = prediction.device
    cls_t = None
    if classes is not None:
        cls_t = torch.tensor(classes, device=dev)
    if prediction.shape[-1] == 6 or end2end:      # nms.py:66-70: end-to-end heads need no suppression, only slicing
        output = [p[p[:, 4] > conf_thres][:max_det] for p in prediction]
        if cls_t is not None:
            output = [p[(p[:, 5:6] == cls_t).any(1)] for p in output]
        return output
    if rotated or multi_label or (labels is not None and len(labels)):
        raise NotImplementedError("rotated / multi_label / labels are outside the B200 hot path (SURVEY 2)")
    if prediction.dim() != 3:
        raise ValueError(f"prediction must be [B, 4+nc, A], got {tuple(prediction.shape)}")
    bs, ch, na = prediction.shape
    nc = nc or (ch - 4)
    extra = ch - nc - 4
    pred = prediction if (prediction.dtype == torch.float32 and prediction.is_contiguous()) else prediction.float().contiguous()
    max_det = int(max_det)
    out_boxes = torch.empty(bs, max(max_det, 1), 6 + extra, dtype=torch.float32, device=dev)
    out_idx = torch.empty(bs, max(max_det, 1), dtype=torch.int64, device=dev)
    out_cnt = torch.zeros(bs, dtype=torch.int32, device=dev)
    L = lib()
    ws = torch.empty(L.ysp_nms_workspace_bytes(bs, ch, na, max_det) + 256, dtype=torch.uint8, device=dev)
    cls_i32 = cls_t.to(torch.int32).contiguous().view(-1) if cls_t is not None else None
    with torch.cuda.device(dev):
        check(L.ysp_nms(pred.data_ptr(), bs, ch, na, nc, conf_thres, iou_thres, max_det, int(max_nms), float(max_wh),
                        int(bool(agnostic)), cls_i32.data_ptr() if cls_i32 is not None else None,
                        cls_i32.numel() if cls_i32 is not None else 0, out_boxes.data_ptr(), out_idx.data_ptr(),
                        out_cnt.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)))
        # nms.py:84-86 side effect on the caller's tensor
        if pred is prediction:
            check(L.ysp_xywh2xyxy_inplace(pred.data_ptr(), bs, ch, na, _stream(dev)))
        else:
            check(L.ysp_xywh2xyxy_inplace(pred.data_ptr(), bs, ch, na, _stream(dev)))
            prediction[:, :4] = pred[:, :4].to(prediction.dtype)
    counts = out_cnt.tolist()                     # the one D2H of the call (ragged lists need host sizes)
    empty = torch.zeros((0, 6 + extra), device=dev)
    empty_k = torch.zeros((0, 1), device=dev)     # the reference's quirky empty keep tensor (float, [0,1])
    output = [out_boxes[b, :n] if n else empty for b, n in enumerate(counts)]
    if not return_idxs:
        return output
    keepi = [out_idx[b, :n] if n else empty_k for b, n in enumerate(counts)]
    return output, keepi


class TorchNMS:
    """nms.py:169-337.  `nms` and `batched_nms` run ysp_nms_core; `fast_nms` (an approximate variant the pipeline
    never calls) is out of scope.

    Semantics pinned to `torchvision.ops.nms` (the back end `non_max_suppression` takes whenever torchvision is imported,
    nms.py:151-154, which is the case in evaluate_model.py): stable descending sort (ties -> lower index first) and a box is
    suppressed iff IoU > threshold.  For a degenerate pair (both boxes of zero area: union 0, IoU = 0/0 = NaN) the
    comparison is false and the box is KEPT, as torchvision does; the reference's own pure-torch `TorchNMS.nms` keeps
    `iou <= thr` (nms.py:288-294) and would therefore drop it, and it breaks score ties with an unstable argsort
    (SURVEY F7).  Zero-area duplicate boxes are the only inputs on which the two differ for distinct scores."""

    @staticmethod
    def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
        if boxes.numel() == 0:
            return torch.empty((0,), dtype=torch.int64, device=boxes.device)
        require_cuda(boxes, "TorchNMS.nms")
        dev = boxes.device
        b = boxes.float().contiguous()
        s = scores.float().contiguous()
        n = b.shape[0]
        keep = torch.empty(n, dtype=torch.int64, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        L = lib()
        ws = torch.empty(L.ysp_nms_workspace_bytes(1, 5, n, n) + 256, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(L.ysp_nms_core(b.data_ptr(), s.data_ptr(), n, float(iou_threshold), keep.data_ptr(), cnt.data_ptr(),
                                 ws.data_ptr(), ws.numel(), _stream(dev)))
        return keep[: int(cnt.item())]

    @staticmethod
    def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float,
                    use_fast_nms: bool = False) -> torch.Tensor:
        if boxes.numel() == 0:
            return torch.empty((0,), dtype=torch.int64, device=boxes.device)
        if use_fast_nms:
            raise NotImplementedError("fast_nms is outside the B200 hot path (SURVEY 2)")
        max_coordinate = boxes.max()
        offsets = idxs.to(boxes) * (max_coordinate + 1)
        return TorchNMS.nms(boxes + offsets[:, None], scores, iou_threshold)

    @staticmethod
    def fast_nms(*a, **k):
        raise NotImplementedError("fast_nms is outside the B200 hot path (SURVEY 2)")
