"""Mask + metric step of the pipeline (evaluate_model.py:157-158, :166-174; monai DiceMetric as configured at
evaluate_model.py:49-56) on top of ysp_mask_dice, plus the cross-rank reduction."""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import check, lib, require_cuda


def mask_counts(pred_logits: torch.Tensor, target: Optional[torch.Tensor], want_mask: bool = False):
    """int32 [B,3] = (|P&T|, |P|, |T|), P = sigmoid(logit) > 0.5 (fp32).  Optionally also the uint8 mask."""
    require_cuda(pred_logits, "mask_counts")
    x = pred_logits if (pred_logits.dtype == torch.float32 and pred_logits.is_contiguous()) else pred_logits.float().contiguous()
    B = x.shape[0]
    HW = x.numel() // max(B, 1)
    t = None
    if target is not None:
        t = target if (target.dtype == torch.float32 and target.is_contiguous()) else target.float().contiguous()
        if t.numel() != x.numel():
            raise ValueError("target and logits differ in size")
    counts = torch.zeros(B, 3, dtype=torch.int32, device=x.device)
    mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_mask else None
    with torch.cuda.device(x.device):
        check(lib().ysp_mask_dice(x.data_ptr(), t.data_ptr() if t is not None else None, B, HW, counts.data_ptr(),
                                  mask.data_ptr() if mask is not None else None,
                                  torch.cuda.current_stream(x.device).cuda_stream))
    return (counts, mask) if want_mask else counts


def dice_from_counts(counts: torch.Tensor) -> torch.Tensor:
    """Per-slice Dice with monai's ignore_empty=False rule: |T|>0 -> 2|P&T|/(|P|+|T|); both empty -> 1; else 0.
    Pure integer->float arithmetic on B*3 numbers (host logic; works on any device)."""
    c = counts.to(torch.float64)
    inter, p, t = c[:, 0], c[:, 1], c[:, 2]
    d = torch.where(t > 0, 2 * inter / (p + t).clamp(min=1), (p == 0).to(torch.float64))
    return d.to(torch.float32)


class SegMetrics:
    """Accumulates what evaluate_model.py:166-187 reports: mean per-slice Dice, TP/FP/FN totals, precision, recall.
    `reduce()` sums the 5 counters across ranks with ONE all-reduce (the only collective of the inference path)."""

    def __init__(self):
        self.state = torch.zeros(5, dtype=torch.float64)      # [sum dice, n slices, TP, FP, FN]

    def update(self, counts: torch.Tensor):
        c = counts.detach().to("cpu", torch.int64)
        d = dice_from_counts(c)
        inter, p, t = (int(c[:, i].sum()) for i in range(3))
        self.state += torch.tensor([float(d.double().sum()), c.shape[0], inter, p - inter, t - inter], dtype=torch.float64)

    def reduce(self, group=None, device=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
            t = self.state.to(dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            self.state = t.cpu()
        return self

    def compute(self):
        sd, n, tp, fp, fn = self.state.tolist()
        # precision / recall exactly as evaluate_model.py:177-178 (the 1e-6 in the denominator included).  Masks are binary
        # here (T = target > 0.5, or PNG byte >= 128): for {0,1} masks TP/FP/FN equal the reference's float sums :166-168.
        # HD95 (evaluate_model.py:58-63, :160-163) is not computed: outside the hot path.
        return {"dice": sd / n if n else float("nan"), "slices": int(n), "TP": int(tp), "FP": int(fp), "FN": int(fn),
                "precision": tp / (tp + fp + 1e-6), "recall": tp / (tp + fn + 1e-6)}
