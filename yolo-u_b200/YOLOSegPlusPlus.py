"""Drop-in for the reference's YOLOSegPlusPlus module (/root/reference/YOLOSegPlusPlus.py:90-272).

Same constructor signature, same state_dict keys (`encoder.N...`, `decoder.N...`, `output.{weight,bias}`, `param`),
so `load_state_dict(torch.load("best.pth"))` (evaluate_model.py:234-243) works unchanged -- but `forward(x, logits)`
does not run PyTorch modules: it hands the tensors to ysp_segpp_forward (csrc/engine.cu), i.e. to the hand-written
sm_100a kernels.  The nn.Modules below are PARAMETER CONTAINERS that reproduce the reference's parameter tree; they
have no forward of their own.  No CPU fallback: calling forward on CPU tensors raises.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn

from .engine import Engine


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the arithmetic runs in libysp (YOLOSegPlusPlus.forward)")


class _Conv(_NoForward):
    """Parameter tree of ultralytics `Conv` (conv.weight, bn.{weight,bias,running_mean,running_var})."""

    def __init__(self, c1, c2, k=1, g=1):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, 1, k // 2, groups=g, bias=False)
        self.bn = nn.BatchNorm2d(c2)


class _LightConv(_NoForward):
    def __init__(self, c1, c2, k):
        super().__init__()
        self.conv1 = _Conv(c1, c2, 1)
        self.conv2 = _Conv(c2, c2, k, g=math.gcd(c2, c2))


class _GhostConv(_NoForward):
    def __init__(self, c1, c2):
        super().__init__()
        c_ = c2 // 2
        self.cv1 = _Conv(c1, c_, 1)
        self.cv2 = _Conv(c_, c_, 5, g=c_)


class _GhostBottleneck(_NoForward):
    def __init__(self, c1, c2):
        super().__init__()
        c_ = c2 // 2
        self.conv = nn.Sequential(_GhostConv(c1, c_), nn.Identity(), _GhostConv(c_, c2))
        self.shortcut = nn.Identity()


class C3Ghost(_NoForward):
    def __init__(self, c1, c2, n=1):
        super().__init__()
        c_ = int(c2 * 0.5)
        self.cv1 = _Conv(c1, c_, 1)
        self.cv2 = _Conv(c1, c_, 1)
        self.cv3 = _Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(_GhostBottleneck(c_, c_) for _ in range(n)))


class DoubleLightConv(_NoForward):
    """YOLOSegPlusPlus.py:33-58 parameter tree."""

    def __init__(self, in_channels, out_channels, k1=3, k2=3):
        super().__init__()
        self.conv = nn.Sequential(_LightConv(in_channels, out_channels, k1), _LightConv(out_channels, out_channels, k2))
        self.residual_conv = (nn.Conv2d(in_channels, out_channels, kernel_size=1) if in_channels != out_channels
                              else nn.Identity())


class ECA(_NoForward):
    """YOLOSegPlusPlus.py:60-88 parameter tree (conv.weight [1,1,3])."""

    def __init__(self, k_size: int = 3):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.conv = nn.Conv1d(1, 1, kernel_size=k_size, padding=(k_size - 1) // 2, bias=False)


class YOLOSegPlusPlus(nn.Module):
    """`YOLOSegPlusPlus(predictor)(x, logits) -> mask logits [B,1,H,W]`.

    predictor: anything with `.model.model.model[0:5]` (the fused detector's first five modules, shared by
    reference like YOLOSegPlusPlus.py:150).  `mode`: "tc32" (tensor-core parity mode, default), "bf16" (throughput) or "fp32" (CUDA-core reference).
    """

    def __init__(self, predictor, verbose: bool = False, target_modules_indices: List[int] = [2, 4, 6],
                 mode: str = "tc32", use_logits: bool = True):
        super().__init__()
        self.encoder = nn.ModuleList(module for module in predictor.model.model.model[0:5])
        for param in self.encoder.parameters():
            param.requires_grad = False
        self.encoder.eval()
        self.upsample = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False)
        self.decoder = nn.ModuleList([
            # use_logits=False = the reference's ablation module `_YOLOSegPlusPlus.py` (no bottleneck channel, :157)
            nn.Sequential(C3Ghost(128 + (1 if use_logits else 0), 96, n=1), ECA()),
            nn.Sequential(self.upsample, DoubleLightConv(96, 64)),
            nn.Sequential(C3Ghost(64 + 64, 64), ECA()),
            nn.Sequential(self.upsample, DoubleLightConv(64, 32)),
            nn.Sequential(self.upsample, DoubleLightConv(32, 16)),
        ])
        self.output = nn.Conv2d(in_channels=16, out_channels=1, kernel_size=1)
        self.sigmoid = nn.Sigmoid()
        self.param = nn.Parameter(torch.tensor([5.0]))
        self.verbose = verbose
        self.skip_connections = []
        self.mode = mode
        self.use_logits = use_logits
        self._engine: Optional[Engine] = None
        self._engine_key = None

    # -- weights -> libysp ---------------------------------------------------------------------------------------
    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def engine(self, device) -> Engine:
        key = (str(device), self.mode, self._weights_version())
        if self._engine is None or self._engine_key != key:
            eng = Engine(device, self.mode)
            eng.load_state_dict("seg", self.state_dict())
            eng.finalize(det=False, seg=True)
            self._engine, self._engine_key = eng, key
        return self._engine

    def inference(self):
        """YOLOSegPlusPlus.py:236-240 is a stub in the reference (`pass`); see predictor.predict for the pipeline."""
        return None

    @torch.no_grad()
    def forward(self, x: torch.Tensor, logits: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [B,4,H,W] fp32, logits [B,1,H/8,W/8] -> [B,1,H,W] mask logits (YOLOSegPlusPlus.py:242-272).
        With `use_logits=False` (ablation checkpoint) `logits` is ignored, as `_YOLOSegPlusPlus.forward(x)` has none."""
        if logits is None or not self.use_logits:
            if self.use_logits:
                raise TypeError("forward() missing the `logits` bottleneck [B,1,H/8,W/8]")
            logits = torch.zeros(x.shape[0], 1, x.shape[2] // 8, x.shape[3] // 8, device=x.device)
        return self.engine(x.device).segpp_forward(x, logits)
