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


def load_reference_segpp(path: str = "/root/reference/YOLOSegPlusPlus.py"):
    """Import the reference's own YOLOSegPlusPlus.py (or the ablation _YOLOSegPlusPlus.py) UNMODIFIED.

    The file needs two absent imports (YOLOSegPlusPlus.py:2,5): `ultralytics.nn.modules` -- stubbed with the restated blocks
    of oracle/modules.py (so what this pins is the reference's OWN code: DoubleLightConv :33-58, ECA :60-88, the decoder
    topology :150-178 and forward's concat / skip-pop order :242-272; the upstream blocks stay a restatement) -- and
    `custom_yolo_predictor.custom_detseg_predictor.CustomDetectionPredictor`, used as a type annotation only.
    The constructor must run where CUDA "is available" (on a CUDA-less host it hits the `verbose` NameError at :229,
    SURVEY F10): use `segpp_ctor_patch()` around the construction.  Only for generating tests/golden/ fixtures here.
    """
    from . import modules as om

    class _NeverBuilt(torch.nn.Module):            # imported by name at :2 but never instantiated on this path
        def __init__(self, *a, **k):
            raise NotImplementedError("not on the hot path")

    pkg = types.ModuleType("ultralytics")
    nn_ = types.ModuleType("ultralytics.nn")
    mods = types.ModuleType("ultralytics.nn.modules")
    for name in ("C3Ghost", "DWConv", "C2f", "Conv", "C3k2", "LightConv"):
        setattr(mods, name, getattr(om, name))
    for name in ("DWConvTranspose2d", "ConvTranspose", "CBAM"):
        setattr(mods, name, _NeverBuilt)
    pkg.nn, nn_.modules = nn_, mods
    cyp = types.ModuleType("custom_yolo_predictor")
    cdp = types.ModuleType("custom_yolo_predictor.custom_detseg_predictor")
    cdp.CustomDetectionPredictor = type("CustomDetectionPredictor", (), {})
    cyp.custom_detseg_predictor = cdp
    names = {"ultralytics": pkg, "ultralytics.nn": nn_, "ultralytics.nn.modules": mods, "custom_yolo_predictor": cyp,
             "custom_yolo_predictor.custom_detseg_predictor": cdp}
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update(names)
    try:
        spec = importlib.util.spec_from_file_location("_reference_segpp_" + str(abs(hash(path))), path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


class segpp_ctor_patch:
    """Context manager: make `torch.cuda.is_available()` true (and get_device_name harmless) while the reference
    YOLOSegPlusPlus constructor runs, so it takes its 'CUDA' branch (:195-196) instead of the broken hook branch (:229)."""

    def __enter__(self):
        self._a, self._n = torch.cuda.is_available, torch.cuda.get_device_name
        torch.cuda.is_available = lambda: True
        torch.cuda.get_device_name = lambda *a, **k: "stub (constructor only)"
        return self

    def __exit__(self, *exc):
        torch.cuda.is_available, torch.cuda.get_device_name = self._a, self._n
        return False


def reference_lines(path: str, first: int, last: int, must_contain: str) -> str:
    """Source lines [first, last] (1-based) of a reference file, dedented -- for executing a fragment of a method body
    verbatim.  `must_contain` guards against the line numbers drifting."""
    import textwrap
    with open(path) as f:
        lines = f.readlines()[first - 1:last]
    src = textwrap.dedent("".join(lines))
    if must_contain not in src:
        raise RuntimeError(f"{path}:{first}-{last} no longer contains {must_contain!r}")
    return src
