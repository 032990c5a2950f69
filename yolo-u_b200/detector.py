"""Drop-in for the detector callable `predictor.model(img)` (ultralytics AutoBackend -> DetectionModel, used at
evaluate_model.py:141, generate_objectmaps.py:88, visualize_logits.py:66): YOLOv12n, 4-channel, nc=1, returning
`[y [B,5,A], [P3, P4, P5]]` with raw maps [B,65,h,w].  Runs ysp_detector_forward; inputs whose H/W are not multiples
of 32 are zero-padded bottom/right inside the first conv (decision D1)."""
from __future__ import annotations

from typing import Mapping

import torch

from .engine import Engine


class B200Detector:
    def __init__(self, det_state_dict: Mapping[str, torch.Tensor], device="cuda:0", mode: str = "tc32", engine: Engine = None):
        self.engine = engine if engine is not None else Engine(device, mode)
        self.engine.load_state_dict("det", det_state_dict)
        self.engine.finalize(det=True, seg=False)

    @classmethod
    def from_predictor(cls, predictor, device="cuda:0", mode: str = "tc32"):
        """predictor.model = AutoBackend; predictor.model.model = DetectionModel (state_dict keys 'model.N...')."""
        return cls(predictor.model.model.state_dict(), device, mode)

    @torch.no_grad()
    def __call__(self, img: torch.Tensor):
        y, raws = self.engine.detector_forward(img, want_raw=True)
        return [y, raws]

    forward = __call__
