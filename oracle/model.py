"""ORACLE (test infrastructure, NOT product code) -- the YOLOv12n detector graph, the YOLO-Seg++ head and
the de-facto inference pipeline, restated in PyTorch fp32 on CPU.

Follows:
  * detector graph: upstream `yolo12.yaml` scale n, 4-ch, nc=1 (layer table printed in
    /root/reference/YOLOSegPlusPlus.py:127-135)
  * seg head: /root/reference/YOLOSegPlusPlus.py:33-58 (DoubleLightConv), :60-88 (ECA), :150-178 (topology),
    :242-272 (forward order, skip pops, concat order)
  * pipeline: /root/reference/evaluate_model.py:141-158 (detector -> sigmoid(P3 cls) -> NMS -> seg -> mask)
  * synthetic weights/inputs: SURVEY 8(d) "Synthetic inputs"

PARITY STATUS: parity unpinned for the conv graph (see oracle/modules.py header).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .modules import (A2C2f, C3Ghost, C3k2, Concat, Conv, Detect, LightConv)

# ----------------------------------------------------------------------------------------------------------
# Detector (upstream DetectionModel for yolo12n, 4-ch, nc=1)
# ----------------------------------------------------------------------------------------------------------

# (from, module, args) -- App. A.3
_YOLO12N = [
    (-1, Conv, (4, 16, 3, 2)),                    # 0
    (-1, Conv, (16, 32, 3, 2)),                   # 1
    (-1, C3k2, (32, 64, 1, False, 0.25)),         # 2
    (-1, Conv, (64, 64, 3, 2)),                   # 3
    (-1, C3k2, (64, 128, 1, False, 0.25)),        # 4
    (-1, Conv, (128, 128, 3, 2)),                 # 5
    (-1, A2C2f, (128, 128, 2, True, 4)),          # 6
    (-1, Conv, (128, 256, 3, 2)),                 # 7
    (-1, A2C2f, (256, 256, 2, True, 1)),          # 8
    (-1, "up", ()),                               # 9
    ((-1, 6), Concat, (1,)),                      # 10
    (-1, A2C2f, (384, 128, 1, False, -1)),        # 11
    (-1, "up", ()),                               # 12
    ((-1, 4), Concat, (1,)),                      # 13
    (-1, A2C2f, (256, 64, 1, False, -1)),         # 14
    (-1, Conv, (64, 64, 3, 2)),                   # 15
    ((-1, 11), Concat, (1,)),                     # 16
    (-1, A2C2f, (192, 128, 1, False, -1)),        # 17
    (-1, Conv, (128, 128, 3, 2)),                 # 18
    ((-1, 8), Concat, (1,)),                      # 19
    (-1, C3k2, (384, 256, 1, True)),              # 20
    ((14, 17, 20), Detect, ()),                   # 21
]


class DetectionModel(nn.Module):
    """`.model` is the nn.Sequential the reference indexes as predictor.model.model.model[0:5]
    (/root/reference/YOLOSegPlusPlus.py:150)."""

    def __init__(self, nc: int = 1, ch: int = 4):
        super().__init__()
        layers, self.froms = [], []
        for i, (f, m, args) in enumerate(_YOLO12N):
            if m == "up":
                mod = nn.Upsample(None, 2, "nearest")
            elif m is Detect:
                mod = Detect(nc, (64, 128, 256))
            elif i == 0:
                mod = Conv(ch, *args[1:])
            else:
                mod = m(*args)
            layers.append(mod)
            self.froms.append(f)
        self.model = nn.Sequential(*layers)
        self.nc = nc
        # upstream `initialize_weights`: every BN of a DetectionModel gets eps=1e-3, momentum=0.03 (App. A.1)
        for mod in self.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.eps, mod.momentum = 1e-3, 0.03

    def fuse(self):
        for mod in self.modules():
            if isinstance(mod, Conv):
                mod.fuse()
        return self

    def forward(self, x):
        y = []
        for f, m in zip(self.froms, self.model):
            if f != -1:
                x = y[f] if isinstance(f, int) else [x if j == -1 else y[j] for j in f]
            x = m(x)
            y.append(x)
        return x


class AutoBackend(nn.Module):
    """Minimal stand-in for upstream AutoBackend(fuse=True): `.model` is the fused, eval() DetectionModel;
    calling it returns the Detect eval output as a list [y, [P3,P4,P5]] (evaluate_model.py:141-143)."""

    def __init__(self, det: DetectionModel):
        super().__init__()
        self.model = det.fuse().eval()
        for p in self.model.parameters():
            p.requires_grad_(False)

    def forward(self, im):
        y = self.model(im)
        return [y[0], y[1]] if isinstance(y, (tuple, list)) else y


class Predictor:
    """Stand-in for CustomDetectionPredictor after setup_model(): only `.model` (AutoBackend) is used."""

    def __init__(self, det: Optional[DetectionModel] = None, seed: int = 0):
        if det is None:
            det = DetectionModel()
            synth_init_(det, seed)
        self.model = AutoBackend(det)


# ----------------------------------------------------------------------------------------------------------
# Seg head (/root/reference/YOLOSegPlusPlus.py)
# ----------------------------------------------------------------------------------------------------------


class DoubleLightConv(nn.Module):
    """YOLOSegPlusPlus.py:33-58: two LightConv(act=SiLU) + 1x1 residual conv (bias) when channels differ."""

    def __init__(self, in_channels, out_channels, k1=3, k2=3):
        super().__init__()
        self.conv = nn.Sequential(LightConv(in_channels, out_channels, k1, act=True),
                                  LightConv(out_channels, out_channels, k2, act=True))
        self.residual_conv = (nn.Conv2d(in_channels, out_channels, kernel_size=1)
                              if in_channels != out_channels else nn.Identity())

    def forward(self, x):
        residual = self.residual_conv(x)
        out = self.conv(x)
        out = out + residual
        return out


class ECA(nn.Module):
    """YOLOSegPlusPlus.py:60-88: GAP -> conv1d(k=3, pad=1, no bias) across channels -> sigmoid -> scale."""

    def __init__(self, k_size: int = 3):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.conv = nn.Conv1d(1, 1, kernel_size=k_size, padding=(k_size - 1) // 2, bias=False)

    def forward(self, x):
        y = self.avg_pool(x)
        y = self.conv(y.squeeze(-1).transpose(-1, -2)).transpose(-1, -2).unsqueeze(-1)
        y = torch.sigmoid(y)
        return x * y.expand_as(x)


class YOLOSegPlusPlus(nn.Module):
    """YOLOSegPlusPlus.py:90-272 without the hook machinery (F10: it only exists on CUDA-less hosts and
    does not change results)."""

    def __init__(self, predictor, verbose: bool = False, target_modules_indices: Sequence[int] = (2, 4, 6)):
        super().__init__()
        self.encoder = nn.ModuleList(m for m in predictor.model.model.model[0:5])
        for p in self.encoder.parameters():
            p.requires_grad = False
        self.encoder.eval()
        self.upsample = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False)
        self.decoder = nn.ModuleList([
            nn.Sequential(C3Ghost(128 + 1, 96, n=1), ECA()),
            nn.Sequential(self.upsample, DoubleLightConv(96, 64)),
            nn.Sequential(C3Ghost(64 + 64, 64), ECA()),
            nn.Sequential(self.upsample, DoubleLightConv(64, 32)),
            nn.Sequential(self.upsample, DoubleLightConv(32, 16)),
        ])
        self.output = nn.Conv2d(16, 1, kernel_size=1)
        self.sigmoid = nn.Sigmoid()
        self.param = nn.Parameter(torch.tensor([5.0]))
        self.verbose = verbose

    def forward(self, x, logits):
        skips = []
        for idx, module in enumerate(self.encoder):
            x = module(x)
            if idx in (2, 4):
                skips.append(x)
        for idx, module in enumerate(self.decoder):
            if idx in (0, 2):
                skip = skips.pop()
                x = torch.cat([skip, logits], 1) if idx == 0 else torch.cat([x, skip], 1)
            x = module(x)
        return self.output(x)


# ----------------------------------------------------------------------------------------------------------
# Synthetic weights: default init gives near-constant logits (F14), so re-scale.
# ----------------------------------------------------------------------------------------------------------


@torch.no_grad()
def synth_init_(module: nn.Module, seed: int = 0, gain: float = 2.0, lin_gain: float = 0.3) -> nn.Module:
    """Deterministic synthetic weights with O(1) activations everywhere (default init gives near-constant
    logits, F14; a uniform gain blows up through the residual attention blocks).
    Conv followed by SiLU ~ N(0, gain/fan_in); linear convs (act=False, plain nn.Conv2d) ~ N(0, lin_gain/fan_in);
    biases ~ N(0,.1); BN gamma U(.5,1.5), beta N(0,.1), mean N(0,.1), var U(.5,1.5).  See calibrate_heads_ for the last layers."""
    g = torch.Generator().manual_seed(seed)
    lin = set()
    for name, m in module.named_modules():
        if isinstance(m, Conv) and not isinstance(m.act, nn.SiLU):
            lin.add(id(m.conv))
    for name, m in module.named_modules():
        if isinstance(m, (nn.Conv2d, nn.Conv1d)):
            if name.endswith("dfl.conv"):
                continue
            fan_in = m.weight[0].numel()
            plain = id(m) in lin or not name.endswith(".conv") and not name == "conv"
            gsel = lin_gain if plain else gain
            if isinstance(m, nn.Conv1d):
                gsel = 3.0
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * math.sqrt(gsel / fan_in))
            if m.bias is not None:
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    return module


@torch.no_grad()
def calibrate_heads_(det: Optional["DetectionModel"], seg: Optional["YOLOSegPlusPlus"], seed: int = 0):
    """Data-dependent re-scaling of the LAST linear layers only, on a fixed seeded batch, so that every seed gives
    a non-vacuous workload: Detect cls logits ~ (mean -2.5, std 2) and box-bin logits std 2 per level; mask
    logits ~ (mean 0, std 1.5) (both signs -> Dice is informative, F14)."""
    g = torch.Generator().manual_seed(seed + 999)
    x = torch.rand(2, 4, 160, 160, generator=g)
    if det is not None:
        was = det.training
        det.eval()
        feats, y = [], []
        h = x
        for f, m in zip(det.froms, det.model):
            if f != -1:
                h = y[f] if isinstance(f, int) else [h if j == -1 else y[j] for j in f]
            if isinstance(m, Detect):
                feats = h
                break
            h = m(h)
            y.append(h)
        head = det.model[-1]
        for i, ft in enumerate(feats):
            for seq, mean_t, std_t in ((head.cv3[i], -2.5, 2.0), (head.cv2[i], 0.0, 2.0)):
                o = seq(ft)
                m_, s_ = o.mean().item(), o.std().item() + 1e-6
                k = std_t / s_
                seq[-1].weight.mul_(k)
                seq[-1].bias.copy_((seq[-1].bias - m_) * k + mean_t)
        det.train(was)
    if seg is not None:
        lg = torch.sigmoid(torch.randn(2, 1, 20, 20, generator=g))
        was = seg.output.weight.clone()
        o = seg(x, lg)
        m_, s_ = o.mean().item(), o.std().item() + 1e-6
        k = 1.5 / s_
        seg.output.weight.mul_(k)
        seg.output.bias.copy_((seg.output.bias - m_) * k)


def build_models(seed: int = 0):
    """(predictor, segpp) with deterministic synthetic weights; decoder BN keeps eps=1e-5 (App. A.1)."""
    det = DetectionModel()
    synth_init_(det, seed)
    calibrate_heads_(det, None, seed)
    pred = Predictor(det)
    seg = YOLOSegPlusPlus(pred)
    synth_init_(seg.decoder, seed + 1, lin_gain=2.0)
    synth_init_(seg.output, seed + 2, lin_gain=2.0)
    seg.eval()
    calibrate_heads_(None, seg, seed)
    return pred, seg


def synth_inputs(batch: int, size: int = 240, seed: int = 0):
    g = torch.Generator().manual_seed(seed + 100)
    x = torch.rand(batch, 4, size, size, generator=g)
    logits = torch.sigmoid(torch.randn(batch, 1, size // 8, size // 8, generator=g))
    target = (torch.rand(batch, 1, size, size, generator=g) > 0.5).float()
    return x, logits, target


# ----------------------------------------------------------------------------------------------------------
# Pipeline (evaluate_model.py:134-174) with decision D1: detector on zero-padded %32 canvas
# ----------------------------------------------------------------------------------------------------------


def pad_to_multiple(x: torch.Tensor, m: int = 32) -> torch.Tensor:
    H, W = x.shape[-2:]
    ph, pw = (-H) % m, (-W) % m
    return F.pad(x, (0, pw, 0, ph)) if (ph or pw) else x


@torch.no_grad()
def pipeline(predictor, segpp, img: torch.Tensor, nms_fn, conf_thres=0.25, iou_thres=0.45):
    """Returns (mask_logits [B,1,H,W], dets list, keep list, y, bottleneck logits)."""
    H, W = img.shape[-2:]
    y, raw = predictor.model(pad_to_multiple(img))               # evaluate_model.py:141-143
    logits = torch.sigmoid(raw[0][:, -1:])[:, :, : H // 8, : W // 8]  # :144 (+ D1 crop)
    dets, keep = nms_fn(y.clone(), conf_thres, iou_thres, return_idxs=True)   # :147
    pred = segpp(img, logits)                                   # :156
    return pred, dets, keep, y, logits


# ----------------------------------------------------------------------------------------------------------
# Mask + metrics (evaluate_model.py:157-174, monai DiceMetric as configured :49-56)
# ----------------------------------------------------------------------------------------------------------


@torch.no_grad()
def mask_counts(pred_logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """int64 [B,3] = (|P∩T|, |P|, |T|) with P = sigmoid(x) > 0.5 evaluated in fp32 (evaluate_model.py:157-158)."""
    p = (torch.sigmoid(pred_logits.float()) > 0.5)
    t = target > 0.5
    B = p.shape[0]
    return torch.stack([(p & t).reshape(B, -1).sum(1), p.reshape(B, -1).sum(1), t.reshape(B, -1).sum(1)], 1)


def dice_from_counts(counts: torch.Tensor) -> torch.Tensor:
    """monai DiceMetric(ignore_empty=False) per sample: |T|>0 -> 2|P∩T|/(|P|+|T|); both empty -> 1; else 0."""
    inter, p, t = (counts[:, i].double() for i in range(3))
    d = torch.where(t > 0, 2 * inter / (p + t).clamp(min=1), torch.where(p > 0, torch.zeros_like(p), torch.ones_like(p)))
    return d.float()


def tp_fp_fn(counts: torch.Tensor):
    inter, p, t = (int(counts[:, i].sum()) for i in range(3))
    return inter, p - inter, t - inter


def dice_loss(pred_logits: torch.Tensor, target: torch.Tensor, smooth: float = 1e-5) -> torch.Tensor:
    """monai DiceLoss(sigmoid, soft_label, batch=True) as in train.py:98-104 (App. A.5)."""
    p = torch.sigmoid(pred_logits)
    sp, st = p.sum(), target.sum()
    tp = (sp + st - (p - target).abs().sum()) / 2
    return 1 - (2 * tp + smooth) / (sp + st + smooth)


# ----------------------------------------------------------------------------------------------------------
# SURVEY 8(f) "next" rows: objectmap consumer transform (dataset.py:86-97) and upstream ops.scale_boxes
# ----------------------------------------------------------------------------------------------------------


def objectmap_transform(objectmap: torch.Tensor) -> torch.Tensor:
    """dataset.py:86-97 for ONE map tensor: z-score with torch.std (unbiased), guarded for std == 0, then sigmoid."""
    mean, std = objectmap.mean(), objectmap.std()
    objectmap = (objectmap - mean) / std if std > 0 else objectmap - mean
    return torch.sigmoid(objectmap)


def scale_boxes(img1_shape, boxes, img0_shape, padding=True):
    """Upstream ultralytics.utils.ops.scale_boxes (xyxy, ratio_pad=None), as called at custom_detseg_predictor.py:177."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    boxes = boxes.clone()
    if padding:
        boxes[..., 0] -= pad[0]
        boxes[..., 1] -= pad[1]
        boxes[..., 2] -= pad[0]
        boxes[..., 3] -= pad[1]
    boxes[..., :4] /= gain
    boxes[..., 0].clamp_(0, img0_shape[1])
    boxes[..., 1].clamp_(0, img0_shape[0])
    boxes[..., 2].clamp_(0, img0_shape[1])
    boxes[..., 3].clamp_(0, img0_shape[0])
    return boxes
