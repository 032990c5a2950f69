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


INTER_NEAREST, INTER_LINEAR = 0, 1          # cv2's values


def resize_u8(src: torch.Tensor, size, interpolation: int = INTER_LINEAR, to_tensor: bool = False):
    """`cv2.resize(src, (W, H), interpolation=...)` for a BATCH of decoded uint8 slices on the device, bit-exact with
    OpenCV (/root/reference/dataset.py:59-65).  src uint8 [B,h,w,4] / [B,h,w,1] / [B,h,w]; `size` int or (H, W).
    Returns uint8 in the input's layout, or with `to_tensor=True` what `transforms.ToTensor()` makes of it
    (dataset.py:68-70): float32 [B,C,H,W] = value / 255."""
    require_cuda(src, "resize_u8")
    if src.dtype != torch.uint8:
        raise TypeError(f"resize_u8 expects uint8 (decoded PNG), got {src.dtype}")
    squeeze = src.dim() == 3
    s4 = (src.unsqueeze(-1) if squeeze else src).contiguous()
    if s4.dim() != 4 or s4.shape[-1] not in (1, 4):
        raise ValueError(f"expected [B,h,w,4], [B,h,w,1] or [B,h,w], got {tuple(src.shape)}")
    B, h, w, C = s4.shape
    H, W = (size, size) if isinstance(size, int) else size
    stream = torch.cuda.current_stream(src.device).cuda_stream
    with torch.cuda.device(src.device):
        if to_tensor:
            out = torch.empty(B, C, H, W, dtype=torch.float32, device=src.device)
            check(lib().ysp_resize_u8(s4.data_ptr(), B, h, w, C, H, W, interpolation, None, out.data_ptr(), stream))
            return out
        out = torch.empty(B, H, W, C, dtype=torch.uint8, device=src.device)
        check(lib().ysp_resize_u8(s4.data_ptr(), B, h, w, C, H, W, interpolation, out.data_ptr(), None, stream))
    return out.squeeze(-1) if squeeze else out


def ingest(img_u8: torch.Tensor, mask_u8: torch.Tensor, size: int):
    """dataset.py:59-70 for a batch of decoded slices: (img float32 [B,4,S,S], mask float32 [B,1,S,S])."""
    return (resize_u8(img_u8, size, INTER_LINEAR, to_tensor=True), resize_u8(mask_u8, size, INTER_NEAREST, to_tensor=True))
