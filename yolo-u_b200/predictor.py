"""The reference's predictor.py is an empty file and YOLOSegPlusPlus.inference() is a stub; the only end-to-end
inference pipeline is the loop body of evaluate_model.py:134-174.  `Predictor.predict` is that loop body:

    yolo_out = YOLO_predictor.model(img)                 :141   detector (zero-padded to %32, decision D1)
    logits   = sigmoid(cls_branch[0][:, -1:])            :142-144  (cropped to H/8 x W/8)
    dets     = non_max_suppression(detect_branch)        :147
    pred     = model(img, logits)                        :156
    mask     = sigmoid(pred) > 0.5 ; TP/FP/FN ; Dice     :157-174

executed as ONE C-ABI call (ysp_pipeline) on hand-written sm_100a kernels."""
from __future__ import annotations

from typing import Mapping, Optional

import torch

from .engine import Engine
from .metrics import dice_from_counts


class Predictor:
    def __init__(self, det_state_dict: Mapping[str, torch.Tensor], seg_state_dict: Mapping[str, torch.Tensor],
                 device="cuda:0", mode: str = "fp32"):
        self.engine = Engine(device, mode)
        self.engine.load_state_dict("det", det_state_dict)
        self.engine.load_state_dict("seg", seg_state_dict)
        self.engine.finalize(det=True, seg=True)
        self._out = {}

    @classmethod
    def from_modules(cls, predictor, segpp, device="cuda:0", mode: str = "fp32"):
        return cls(predictor.model.model.state_dict(), segpp.state_dict(), device, mode)

    @torch.no_grad()
    def predict_raw(self, img: torch.Tensor, target: Optional[torch.Tensor] = None, conf_thres: float = 0.25,
                    iou_thres: float = 0.45, max_det: int = 300, want_mask: bool = False):
        """Device-side results, padded, no synchronisation (for benchmarking / graph capture)."""
        return self.engine.pipeline(img, target, conf_thres, iou_thres, max_det, out=self._out, want_mask=want_mask)

    @torch.no_grad()
    def predict(self, img: torch.Tensor, target: Optional[torch.Tensor] = None, conf_thres: float = 0.25,
                iou_thres: float = 0.45, max_det: int = 300):
        """Returns (mask_logits [B,1,H,W], dets list[[n_i,6]], keep_idx list[int64 [n_i]], counts int32 [B,3])."""
        o = self.predict_raw(img, target, conf_thres, iou_thres, max_det)
        n = o["det_count"].tolist()
        dets = [o["det_boxes"][b, :k] for b, k in enumerate(n)]
        keep = [o["det_idx"][b, :k] for b, k in enumerate(n)]
        return o["mask_logits"], dets, keep, o["counts"]


def predict(engine_or_predictor, img, **kw):
    return engine_or_predictor.predict(img, **kw)
