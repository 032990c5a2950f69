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
    """`replicas` > 1 builds that many engine handles over the same weights (own workspaces, own streams).  Consecutive
    batches then run on alternate handles (`submit_raw`, `HostPipeline`): the tail of batch i -- the last decoder stage,
    mask/Dice, NMS -- and the small-map detector layers, whose grids do not fill 148 SMs, run under the kernels of batch
    i+1.  Measured at B = 256: 7.84 -> 7.54 ms per batch (tc32), 4.80 -> 4.57 ms (bf16); a third handle adds nothing."""

    def __init__(self, det_state_dict: Mapping[str, torch.Tensor], seg_state_dict: Mapping[str, torch.Tensor],
                 device="cuda:0", mode: str = "tc32", replicas: int = 1):
        self.engine = Engine(device, mode)
        self.engine.load_state_dict("det", det_state_dict)
        self.engine.load_state_dict("seg", seg_state_dict)
        self.engine.finalize(det=True, seg=True)
        self._out = {}
        if replicas < 1:
            raise ValueError(f"replicas must be >= 1, got {replicas}")
        self.replicas = [self] + [Predictor(det_state_dict, seg_state_dict, device, mode) for _ in range(replicas - 1)]
        self._streams = None
        self._turn = 0

    @classmethod
    def from_modules(cls, predictor, segpp, device="cuda:0", mode: str = "tc32", replicas: int = 1):
        return cls(predictor.model.model.state_dict(), segpp.state_dict(), device, mode, replicas)

    @property
    def launches_total(self) -> int:
        """kernels launched so far by all replicas"""
        return sum(r.engine.launches_total for r in self.replicas)

    @torch.no_grad()
    def submit_raw(self, img: torch.Tensor, target: Optional[torch.Tensor] = None, after=None, **kw):
        """`predict_raw` of the next batch on the next replica, on that replica's own stream (ordered after the work already
        queued on the caller's current stream).  Returns (outputs, event): the replica's output dict -- valid until its next
        turn, i.e. for `len(replicas)` submissions -- and the event recorded behind the batch.  `after(outputs)`, if given,
        runs inside the replica's stream right behind the batch (e.g. to copy the counters away).  `join()` orders the
        current stream after everything submitted."""
        dev = self.engine.device
        if self._streams is None:
            self._streams = [torch.cuda.Stream(dev) for _ in self.replicas]
        r = self._turn % len(self.replicas)
        self._turn += 1
        st = self._streams[r]
        st.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(st):
            o = self.replicas[r].predict_raw(img, target, **kw)
            if after is not None:
                after(o)
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(st)
        for t in (img, target):
            if t is not None:
                t.record_stream(st)
        return o, ev

    def join(self):
        if self._streams is not None:
            cur = torch.cuda.current_stream(self.engine.device)
            for st in self._streams:
                cur.wait_stream(st)

    @torch.no_grad()
    def predict_raw(self, img: torch.Tensor, target: Optional[torch.Tensor] = None, conf_thres: float = 0.25,
                    iou_thres: float = 0.45, max_det: int = 300, want_mask: bool = False, want_bits: bool = False):
        """Device-side results, padded, no synchronisation (for benchmarking / graph capture)."""
        return self.engine.pipeline(img, target, conf_thres, iou_thres, max_det, out=self._out, want_mask=want_mask,
                                    want_bits=want_bits)

    @torch.no_grad()
    def predict(self, img: torch.Tensor, target: Optional[torch.Tensor] = None, conf_thres: float = 0.25,
                iou_thres: float = 0.45, max_det: int = 300, conf_gate: Optional[float] = None):
        """Returns (mask_logits [B,1,H,W], dets list[[n_i,6]], keep_idx list[int64 [n_i]], counts int32 [B,3]).
        `conf_gate` (off by default, like the commented-out block at evaluate_model.py:149-155): slices without a detection
        above it count as an empty predicted mask; the gated flags are left in `self.gated` (uint8 [B])."""
        o = self.predict_raw(img, target, conf_thres, iou_thres, max_det, want_mask=conf_gate is not None)
        if conf_gate is not None:
            from ._lib import check, lib
            B = o["det_count"].shape[0]
            self.gated = torch.empty(B, dtype=torch.uint8, device=o["counts"].device)
            mask = o.get("mask")
            with torch.cuda.device(o["counts"].device):
                check(lib().ysp_conf_gate(o["det_boxes"].data_ptr(), o["det_count"].data_ptr(), B, o["det_boxes"].shape[1],
                                          o["det_boxes"].shape[2], float(conf_gate), o["counts"].data_ptr(),
                                          mask.data_ptr() if mask is not None else None,
                                          o["mask_logits"].shape[-1] * o["mask_logits"].shape[-2], self.gated.data_ptr(),
                                          torch.cuda.current_stream(o["counts"].device).cuda_stream))
        n = o["det_count"].tolist()
        dets = [o["det_boxes"][b, :k] for b, k in enumerate(n)]
        keep = [o["det_idx"][b, :k] for b, k in enumerate(n)]
        return o["mask_logits"], dets, keep, o["counts"]


class HostPipeline:
    """End-to-end driver for HOST batches: pinned u8 [B,H,W,4] slices in, padded detections + Dice counters (+ optionally the
    bit-packed predicted mask) out (host).

    Double-buffered over an upload, a download and one compute stream per predictor replica, so the H2D copy of batch i+1
    and the D2H read of batch i-1 overlap the kernels of batch i (and, with `Predictor(replicas=2)`, consecutive batches
    compute on alternate engine handles and overlap each other's tails) (PCIe moves 59 MB of slices + 15 MB of u8 masks per 256 slices; the step itself is ~5-10 ms).
    `submit` is asynchronous; `results(i)` returns the host tensors of slot i after `synchronize()` (or after the slot's
    event completed).  Ground-truth masks go up as uint8 (the PNG bytes, dataset.py:55) -- fp32 masks are accepted too but
    cost four times the upload.  `return_mask=True` adds `mask_bits` (int32 [B, H*W/32]) to the results: the mask a
    predict() caller wants, at 1/32 of the size of the fp32 logits."""

    KEYS = ("counts", "det_count", "det_boxes", "det_idx")

    def __init__(self, predictor: "Predictor", B: int, H: int, W: int, max_det: int = 300, conf_thres: float = 0.25,
                 iou_thres: float = 0.45, return_mask: bool = False):
        self.P, self.B, self.H, self.W = predictor, B, H, W
        self.Ps = list(getattr(predictor, "replicas", [predictor]))[:2]      # slot k computes on replica k % len: two slots
        self.kw = dict(conf_thres=conf_thres, iou_thres=iou_thres, max_det=max_det, want_bits=return_mask)
        if return_mask:
            self.KEYS = self.KEYS + ("mask_bits",)
        dev = predictor.engine.device
        self.dev = dev
        self.s_in, self.s_out = (torch.cuda.Stream(dev) for _ in range(2))
        self.run_streams = [torch.cuda.Stream(dev) for _ in self.Ps]          # one compute stream per replica
        self.s_run = self.run_streams[0]
        self.d_img = [torch.empty(B, H, W, 4, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.d_tgt = [None, None]
        self.d_out = [dict(), dict()]
        self.h_out = [None, None]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_run = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.n = 0
        self.last_n = [B, B]
        self.h2d_bytes = B * H * W * 4
        self.d2h_bytes = 0

    def submit(self, h_img_u8: torch.Tensor, h_target: Optional[torch.Tensor] = None) -> int:
        """h_img_u8: pinned uint8 [n,H,W,4], n <= B (a ragged last batch keeps its own device / host result buffers);
        h_target: optional pinned ground-truth masks, uint8 [n,H,W] (PNG bytes) or fp32 [n,1,H,W] (both HOST).  Returns the
        slot (0/1) holding this batch's results."""
        k = self.n % 2
        self.n += 1
        n = int(h_img_u8.shape[0])
        if n > self.B:
            raise ValueError(f"batch of {n} slices exceeds the pipeline's capacity {self.B}")
        d_target = None
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_run[k])          # slot's previous compute has consumed d_img[k] / d_tgt[k]
            d_img = self.d_img[k][:n]
            d_img.copy_(h_img_u8, non_blocking=True)
            if h_target is not None:
                shape = (self.B,) + tuple(h_target.shape[1:])
                if self.d_tgt[k] is None or self.d_tgt[k].dtype != h_target.dtype or self.d_tgt[k].shape != shape:
                    self.d_tgt[k] = torch.empty(shape, dtype=h_target.dtype, device=self.dev)
                d_target = self.d_tgt[k][:n]
                d_target.copy_(h_target, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        self.h2d_bytes = h_img_u8.numel() + (h_target.numel() * h_target.element_size() if h_target is not None else 0)
        s_run, P = self.run_streams[k % len(self.Ps)], self.Ps[k % len(self.Ps)]
        with torch.cuda.stream(s_run):
            s_run.wait_event(self.ev_in[k])
            s_run.wait_event(self.ev_out[k])              # slot's previous results have left the device buffers
            o = P.engine.pipeline(d_img, d_target, out=self.d_out[k].setdefault(n, {}), **self.kw)
            self.ev_run[k].record(s_run)
        if self.h_out[k] is None:
            self.h_out[k] = {}
        if n not in self.h_out[k]:
            self.h_out[k][n] = {key: torch.empty(o[key].shape, dtype=o[key].dtype).pin_memory() for key in self.KEYS}
        h = self.h_out[k][n]
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in h.values())
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_run[k])
            for key in self.KEYS:
                h[key].copy_(o[key], non_blocking=True)
            self.ev_out[k].record(self.s_out)
        self.last_n[k] = n
        return k

    def results(self, slot: int):
        """Host tensors of the batch last submitted to `slot` (valid until that slot is submitted to again)."""
        self.ev_out[slot].synchronize()
        return self.h_out[slot][self.last_n[slot]]

    def streams(self):
        return [self.s_in, *self.run_streams, self.s_out]

    def synchronize(self):
        for s in self.streams():
            s.synchronize()


def predict(engine_or_predictor, img, **kw):
    return engine_or_predictor.predict(img, **kw)
