"""Synthetic checkpoints for benchmarking (there is no network for datasets or trained weights): random-init
state_dicts with exactly the reference's parameter names and shapes -- the fused 4-channel nc=1 YOLOv12n detector
(ultralytics AutoBackend(fuse=True): `model.N...conv.{weight,bias}`) and the YOLO-Seg++ head
(/root/reference/YOLOSegPlusPlus.py:150-182: `encoder.*` shared with detector layers 0-4, unfused `decoder.*`,
`output.*`, `param`).  Weights use fan-in scaling so activations stay O(1); `calibrate` then rescales only the last
linear layers using THIS library's own forward so that class scores / mask logits have a realistic spread."""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

Spec = List[Tuple[str, Tuple[int, ...], str]]      # (key, shape, kind)


def _conv(s: Spec, p: str, c1: int, c2: int, k: int = 1, g: int = 1, act: bool = True, fused: bool = True):
    s.append((f"{p}.conv.weight", (c2, c1 // g, k, k), "w_act" if act else "w_lin"))
    if fused:
        s.append((f"{p}.conv.bias", (c2,), "bias"))
    else:
        s += [(f"{p}.bn.weight", (c2,), "bn_g"), (f"{p}.bn.bias", (c2,), "bn_b"), (f"{p}.bn.running_mean", (c2,), "bn_m"),
              (f"{p}.bn.running_var", (c2,), "bn_v"), (f"{p}.bn.num_batches_tracked", (), "nbt")]


def _bottleneck(s, p, c1, c2, k=(3, 3), e=0.5):
    c_ = int(c2 * e)
    _conv(s, f"{p}.cv1", c1, c_, k[0])
    _conv(s, f"{p}.cv2", c_, c2, k[1])


def _c3k(s, p, c1, c2, n=2):
    c_ = int(c2 * 0.5)
    _conv(s, f"{p}.cv1", c1, c_)
    _conv(s, f"{p}.cv2", c1, c_)
    _conv(s, f"{p}.cv3", 2 * c_, c2)
    for i in range(n):
        _bottleneck(s, f"{p}.m.{i}", c_, c_, (3, 3), 1.0)


def _c3k2(s, p, c1, c2, c3k=False, e=0.5):
    c = int(c2 * e)
    _conv(s, f"{p}.cv1", c1, 2 * c)
    _conv(s, f"{p}.cv2", 3 * c, c2)
    if c3k:
        _c3k(s, f"{p}.m.0", c, c, 2)
    else:
        _bottleneck(s, f"{p}.m.0", c, c)


def _ablock(s, p, dim):
    _conv(s, f"{p}.attn.qkv", dim, 3 * dim, act=False)
    _conv(s, f"{p}.attn.proj", dim, dim, act=False)
    _conv(s, f"{p}.attn.pe", dim, dim, 7, g=dim, act=False)
    _conv(s, f"{p}.mlp.0", dim, 2 * dim)
    _conv(s, f"{p}.mlp.1", 2 * dim, dim, act=False)


def _a2c2f(s, p, c1, c2, n, a2):
    c_ = int(c2 * 0.5)
    _conv(s, f"{p}.cv1", c1, c_)
    _conv(s, f"{p}.cv2", (1 + n) * c_, c2)
    for i in range(n):
        if a2:
            _ablock(s, f"{p}.m.{i}.0", c_)
            _ablock(s, f"{p}.m.{i}.1", c_)
        else:
            _c3k(s, f"{p}.m.{i}", c_, c_, 2)


def detector_spec() -> Spec:
    """Fused DetectionModel state_dict, module-registration order."""
    s: Spec = []
    _conv(s, "model.0", 4, 16, 3)
    _conv(s, "model.1", 16, 32, 3)
    _c3k2(s, "model.2", 32, 64, False, 0.25)
    _conv(s, "model.3", 64, 64, 3)
    _c3k2(s, "model.4", 64, 128, False, 0.25)
    _conv(s, "model.5", 128, 128, 3)
    _a2c2f(s, "model.6", 128, 128, 2, True)
    _conv(s, "model.7", 128, 256, 3)
    _a2c2f(s, "model.8", 256, 256, 2, True)
    _a2c2f(s, "model.11", 384, 128, 1, False)
    _a2c2f(s, "model.14", 256, 64, 1, False)
    _conv(s, "model.15", 64, 64, 3)
    _a2c2f(s, "model.17", 192, 128, 1, False)
    _conv(s, "model.18", 128, 128, 3)
    _c3k2(s, "model.20", 384, 256, True, 0.5)
    for i, ch in enumerate((64, 128, 256)):
        _conv(s, f"model.21.cv2.{i}.0", ch, 64, 3)
        _conv(s, f"model.21.cv2.{i}.1", 64, 64, 3)
        s += [(f"model.21.cv2.{i}.2.weight", (64, 64, 1, 1), "w_box"), (f"model.21.cv2.{i}.2.bias", (64,), "bias")]
    for i, ch in enumerate((64, 128, 256)):
        _conv(s, f"model.21.cv3.{i}.0.0", ch, ch, 3, g=ch)
        _conv(s, f"model.21.cv3.{i}.0.1", ch, 64)
        _conv(s, f"model.21.cv3.{i}.1.0", 64, 64, 3, g=64)
        _conv(s, f"model.21.cv3.{i}.1.1", 64, 64)
        s += [(f"model.21.cv3.{i}.2.weight", (1, 64, 1, 1), "w_cls"), (f"model.21.cv3.{i}.2.bias", (1,), "bias_cls")]
    s.append(("model.21.dfl.conv.weight", (1, 16, 1, 1), "dfl"))
    return s


def _ghostconv(s, p, c1, c2, act):
    c_ = c2 // 2
    _conv(s, f"{p}.cv1", c1, c_, 1, act=act, fused=False)
    _conv(s, f"{p}.cv2", c_, c_, 5, g=c_, act=act, fused=False)


def _c3ghost(s, p, c1, c2):
    c_ = int(c2 * 0.5)
    _conv(s, f"{p}.cv1", c1, c_, fused=False)
    _conv(s, f"{p}.cv2", c1, c_, fused=False)
    _conv(s, f"{p}.cv3", 2 * c_, c2, fused=False)
    _ghostconv(s, f"{p}.m.0.conv.0", c_, c_ // 2, True)
    _ghostconv(s, f"{p}.m.0.conv.2", c_ // 2, c_, False)


def _doublelight(s, p, c1, c2):
    for i, cin in enumerate((c1, c2)):
        _conv(s, f"{p}.conv.{i}.conv1", cin, c2, 1, act=False, fused=False)
        _conv(s, f"{p}.conv.{i}.conv2", c2, c2, 3, g=c2, act=True, fused=False)
    s += [(f"{p}.residual_conv.weight", (c2, c1, 1, 1), "w_lin"), (f"{p}.residual_conv.bias", (c2,), "bias")]


def seg_spec() -> Spec:
    """YOLOSegPlusPlus state_dict: param, encoder.0-4 (fused, = detector layers 0-4), decoder, output."""
    s: Spec = [("param", (1,), "param")]
    det = [(k, sh, kind) for k, sh, kind in detector_spec() if int(k.split(".")[1]) <= 4]
    s += [("encoder." + k[len("model."):], sh, kind) for k, sh, kind in det]
    _c3ghost(s, "decoder.0.0", 129, 96)
    s.append(("decoder.0.1.conv.weight", (1, 1, 3), "eca"))
    _doublelight(s, "decoder.1.1", 96, 64)
    _c3ghost(s, "decoder.2.0", 128, 64)
    s.append(("decoder.2.1.conv.weight", (1, 1, 3), "eca"))
    _doublelight(s, "decoder.3.1", 64, 32)
    _doublelight(s, "decoder.4.1", 32, 16)
    s += [("output.weight", (1, 16, 1, 1), "w_out"), ("output.bias", (1,), "bias")]
    return s


def _fill(spec: Spec, g: torch.Generator, lin_gain: float) -> "OrderedDict[str, torch.Tensor]":
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape, kind in spec:
        if kind in ("w_act", "w_lin", "w_box", "w_cls", "w_out"):
            fan_in = shape[1] * shape[2] * shape[3]
            gain = 2.0 if kind == "w_act" else lin_gain
            sd[key] = torch.randn(shape, generator=g) * math.sqrt(gain / fan_in)
        elif kind == "eca":
            sd[key] = torch.randn(shape, generator=g)
        elif kind in ("bias", "bn_b", "bn_m"):
            sd[key] = torch.randn(shape, generator=g) * 0.1
        elif kind == "bias_cls":
            sd[key] = torch.full(shape, -2.5)
        elif kind in ("bn_g", "bn_v"):
            sd[key] = torch.rand(shape, generator=g) + 0.5
        elif kind == "nbt":
            sd[key] = torch.zeros((), dtype=torch.long)
        elif kind == "dfl":
            sd[key] = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
        elif kind == "param":
            sd[key] = torch.tensor([5.0])
        else:  # pragma: no cover
            raise KeyError(kind)
    return sd


def synth_state_dicts(seed: int = 0) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """(detector_sd, seg_sd) with the seg encoder sharing the detector's layer 0-4 tensors."""
    g = torch.Generator().manual_seed(seed)
    det = _fill(detector_spec(), g, lin_gain=0.3)
    seg = _fill(seg_spec(), g, lin_gain=2.0)
    for k in list(seg):
        if k.startswith("encoder."):
            seg[k] = det["model." + k[len("encoder."):]]
    return det, seg


@torch.no_grad()
def calibrate(det_sd, seg_sd, device="cuda:0", size: int = 240, seed: int = 0):
    """Rescale ONLY the final 1x1 convs (Detect cv2/cv3 last layers, seg `output`) from statistics of this library's
    own fp32 forward on a seeded batch: class logits ~ (mean -2.5, std 2), box-bin logits std 2, mask logits ~
    (mean 0, std 1.5).  Returns new dicts.  Needs a GPU (the forward is libysp's)."""
    from .engine import Engine
    g = torch.Generator().manual_seed(seed + 999)
    x = torch.rand(4, 4, size, size, generator=g).to(device)
    eng = Engine(device, "fp32")
    eng.load_state_dict("det", det_sd)
    eng.load_state_dict("seg", seg_sd)
    eng.finalize(True, True)
    _, raws = eng.detector_forward(x)
    det_sd, seg_sd = OrderedDict(det_sd), OrderedDict(seg_sd)
    for i, r in enumerate(raws):
        for ch, pre, mean_t, std_t in ((slice(64, 65), f"model.21.cv3.{i}.2", -2.5, 2.0), (slice(0, 64), f"model.21.cv2.{i}.2", 0.0, 2.0)):
            o = r[:, ch]
            m_, s_ = o.mean().item(), o.std().item() + 1e-6
            k = std_t / s_
            det_sd[pre + ".weight"] = det_sd[pre + ".weight"] * k
            det_sd[pre + ".bias"] = (det_sd[pre + ".bias"] - m_) * k + mean_t
    lg = torch.sigmoid(torch.randn(4, 1, size // 8, size // 8, generator=g)).to(device)
    o = eng.segpp_forward(x, lg)
    m_, s_ = o.mean().item(), o.std().item() + 1e-6
    k = 1.5 / s_
    seg_sd["output.weight"] = seg_sd["output.weight"] * k
    seg_sd["output.bias"] = (seg_sd["output.bias"] - m_) * k
    del eng
    return det_sd, seg_sd
