"""SURVEY 8(f) "next" rows on either side of the hot path:

* the bottleneck producer/consumer format -- `save_objectmaps` writes what generate_objectmaps.py:91-106 writes (the raw
  P3 class-logit map `[1,1,h,w]` per image as `<name>_20.pt`), `objectmap_transform` is the training-time consumer of
  dataset.py:86-97 (`sigmoid(zscore(map))`) as one device kernel;
* `scale_boxes` -- ultralytics `ops.scale_boxes` as called at custom_detseg_predictor.py:177.
"""
from __future__ import annotations

import os
from typing import Sequence

import torch

from ._lib import check, lib, require_cuda


def objectmap_transform(maps: torch.Tensor) -> torch.Tensor:
    """maps [B,1,h,w] (or [B,h,w]) raw class-logit maps -> sigmoid((x - mean) / std) per map (torch.std semantics)."""
    require_cuda(maps, "objectmap_transform")
    x = maps if (maps.dtype == torch.float32 and maps.is_contiguous()) else maps.float().contiguous()
    B = x.shape[0]
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib().ysp_objectmap_transform(x.data_ptr(), out.data_ptr(), B, x.numel() // max(B, 1),
                                            torch.cuda.current_stream(x.device).cuda_stream))
    return out


def save_objectmaps(detector, imgs: torch.Tensor, names: Sequence[str], out_dir: str, suffix: str = "_20"):
    """generate_objectmaps.py:86-106 for a batch: detector(imgs) -> P3[:, -1:] saved per image as [1,1,h,w] .pt files."""
    y, raws = detector(imgs)
    p3 = raws[0][:, -1:].cpu()
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for i, n in enumerate(names):
        path = os.path.join(out_dir, f"{n}{suffix}.pt")
        torch.save(p3[i:i + 1].clone(), path)
        paths.append(path)
    return paths


def scale_boxes(img1_shape, boxes: torch.Tensor, img0_shape, ratio_pad=None, padding: bool = True, xywh: bool = False):
    """Rescale xyxy boxes (in place, like upstream) from the network canvas `img1_shape` (h, w) to the original image
    `img0_shape` (h, w[, c]).  `padding=True` assumes ultralytics' centred letterbox; decision D1's bottom/right zero
    padding corresponds to `padding=False` (gain 1: boxes are only clipped)."""
    if xywh:
        raise NotImplementedError("xywh rescaling is not used on this path")
    require_cuda(boxes, "scale_boxes")
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad_x = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)
        pad_y = round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    else:
        gain = ratio_pad[0][0]
        pad_x, pad_y = ratio_pad[1]
    if not padding:
        pad_x = pad_y = 0
    if boxes.numel() == 0:
        return boxes
    if boxes.dtype != torch.float32 or boxes.stride(-1) != 1:
        raise ValueError("boxes must be fp32 with a contiguous last dimension")
    row = boxes.stride(-2) if boxes.dim() > 1 else boxes.shape[-1]
    n = boxes.numel() // boxes.shape[-1]
    if boxes.dim() > 2 and not boxes.is_contiguous():
        raise ValueError("batched boxes must be contiguous")
    with torch.cuda.device(boxes.device):
        check(lib().ysp_scale_boxes(boxes.data_ptr(), n, row, float(gain), float(pad_x), float(pad_y), float(img0_shape[1]),
                                    float(img0_shape[0]), torch.cuda.current_stream(boxes.device).cuda_stream))
    return boxes
