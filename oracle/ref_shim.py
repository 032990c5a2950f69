"""ORACLE (test infrastructure) -- import the reference's own nms.py in THIS container.

/root/reference/nms.py imports three names from the absent `ultralytics` package (nms.py:8-10).  This shim
installs a stub package providing them (restated per SURVEY App. A.4; `batch_probiou` is never reached on the
non-rotated path) and loads the reference file unmodified with importlib.  `import torchvision` first so the
file takes its `torchvision.ops.nms` branch (nms.py:151-154).  Only usable where /root/reference exists, i.e.
for generating tests/golden/ fixtures -- never at test or bench time on the GPU box.
"""
from __future__ import annotations

import importlib.util
import logging
import sys
import types

import torch


def load_reference_nms(path: str = "/root/reference/nms.py", use_torchvision: bool = True):
    if use_torchvision:
        import torchvision  # noqa: F401  (flips nms.py:151 onto the torchvision branch)
    from .nms import xywh2xyxy

    def box_iou(box1, box2, eps=1e-7):
        (a1, a2), (b1, b2) = box1.float().unsqueeze(1).chunk(2, 2), box2.float().unsqueeze(0).chunk(2, 2)
        inter = (torch.min(a2, b2) - torch.max(a1, b1)).clamp_(0).prod(2)
        return inter / ((a2 - a1).prod(2) + (b2 - b1).prod(2) - inter + eps)

    def batch_probiou(*a, **k):
        raise NotImplementedError("rotated path is outside the hot path")

    pkg = types.ModuleType("ultralytics")
    utils = types.ModuleType("ultralytics.utils")
    utils.LOGGER = logging.getLogger("ultralytics-shim")
    metrics = types.ModuleType("ultralytics.utils.metrics")
    metrics.box_iou, metrics.batch_probiou = box_iou, batch_probiou
    ops = types.ModuleType("ultralytics.utils.ops")
    ops.xywh2xyxy = xywh2xyxy
    pkg.utils, utils.metrics, utils.ops = utils, metrics, ops
    saved = {k: sys.modules.get(k) for k in ("ultralytics", "ultralytics.utils", "ultralytics.utils.metrics",
                                             "ultralytics.utils.ops")}
    sys.modules.update({"ultralytics": pkg, "ultralytics.utils": utils, "ultralytics.utils.metrics": metrics,
                        "ultralytics.utils.ops": ops})
    try:
        spec = importlib.util.spec_from_file_location("_reference_nms", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod
