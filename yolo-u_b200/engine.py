"""Engine: one libysp handle (device + arithmetic mode) with its packed weights and a cached device workspace.
PyTorch is used only as plumbing here: device memory (torch.empty), streams, and tensors at the API boundary."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional

import torch

from . import _lib
from ._lib import PipelineIO, check, lib, require_cuda


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


class Engine:
    def __init__(self, device="cuda:0", mode="fp32"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.YspError(f"Engine needs a CUDA device, got {self.device}; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.mode = _lib.mode_id(mode)
        h = C.c_void_p()
        check(lib().ysp_create(C.byref(h), self.device.index, self.mode))
        self._h = h
        self._ws: Optional[torch.Tensor] = None
        self._which = 0
        self.launches_total = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ysp_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- weights -----------------------------------------------------------------------------------------------
    def load_state_dict(self, prefix: str, sd: Mapping[str, torch.Tensor]):
        """prefix 'det' (DetectionModel keys 'model.N...') or 'seg' (YOLOSegPlusPlus keys)."""
        L = lib()
        for k, v in sd.items():
            if not torch.is_tensor(v) or not v.dtype.is_floating_point:
                continue                                    # num_batches_tracked etc.
            t = v.detach().to("cpu", torch.float32).contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            check(L.ysp_load_weight(self._h, f"{prefix}.{k}".encode(), t.data_ptr(), t.dim(), shape))
        return self

    def finalize(self, det: bool, seg: bool):
        which = (1 if det else 0) | (2 if seg else 0)
        check(lib().ysp_finalize(self._h, which))
        self._which |= which
        return self

    def keep_intermediates(self, on: bool = True):
        check(lib().ysp_set_keep_intermediates(self._h, int(on)))

    # ---- workspace ---------------------------------------------------------------------------------------------
    def workspace(self, nbytes: int) -> torch.Tensor:
        """One cached arena per Engine.  An Engine serves ONE stream at a time (libysp's own lanes / aux stream fork from and
        join back into the caller's stream inside each call); before the arena is replaced by a larger one the device is
        synchronised, so no kernel of an earlier call -- on whatever stream it ran -- can still be using the old block
        when the caching allocator hands it out again."""
        if self._ws is None or self._ws.numel() < nbytes:
            if self._ws is not None:
                torch.cuda.synchronize(self.device)
            self._ws = None
            self._ws = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=self.device)
        return self._ws

    def _ws_for(self, B, H, W, max_det: int = 300) -> torch.Tensor:
        n = lib().ysp_pipeline_workspace_bytes(self._h, B, H, W, int(max_det))
        if n == 0:
            raise _lib.YspError("ysp_workspace_bytes failed: " + lib().ysp_last_error().decode())
        return self.workspace(n)

    def _count(self):
        self.launches_total += lib().ysp_last_launch_count(self._h)

    # ---- ops ---------------------------------------------------------------------------------------------------
    def detector_forward(self, img: torch.Tensor, want_raw: bool = True):
        require_cuda(img, "detector")
        if img.dim() != 4 or img.shape[1] != 4:
            raise ValueError(f"expected img [B,4,H,W], got {tuple(img.shape)}")
        img = _f32c(img)
        B, _, H, W = img.shape
        SH, SW = (H + 31) // 32 * 32, (W + 31) // 32 * 32
        hw = [(SH // s, SW // s) for s in (8, 16, 32)]
        A = sum(h * w for h, w in hw)
        y = torch.empty(B, 5, A, dtype=torch.float32, device=img.device)
        raws = [torch.empty(B, 65, h, w, dtype=torch.float32, device=img.device) for h, w in hw] if want_raw else [None] * 3
        ws = self._ws_for(B, H, W)
        with torch.cuda.device(self.device):
            check(lib().ysp_detector_forward(self._h, img.data_ptr(), B, H, W, y.data_ptr(),
                                             *[r.data_ptr() if r is not None else None for r in raws],
                                             ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
        self._count()
        return y, raws

    def segpp_forward(self, x: torch.Tensor, logits: torch.Tensor) -> torch.Tensor:
        require_cuda(x, "YOLOSegPlusPlus.forward")
        require_cuda(logits, "YOLOSegPlusPlus.forward")
        if x.dim() != 4 or x.shape[1] != 4:
            raise ValueError(f"expected x [B,4,H,W], got {tuple(x.shape)}")
        B, _, H, W = x.shape
        if H % 8 or W % 8:
            raise RuntimeError(f"H and W must be multiples of 8, got {H}x{W}")
        if tuple(logits.shape) != (B, 1, H // 8, W // 8):
            raise RuntimeError(f"Sizes of tensors must match: logits {tuple(logits.shape)} vs expected {(B, 1, H // 8, W // 8)}")
        x, logits = _f32c(x), _f32c(logits)
        out = torch.empty(B, 1, H, W, dtype=torch.float32, device=x.device)
        ws = self._ws_for(B, H, W)
        with torch.cuda.device(self.device):
            check(lib().ysp_segpp_forward(self._h, x.data_ptr(), logits.data_ptr(), out.data_ptr(), B, H, W,
                                          ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
        self._count()
        return out

    def pipeline(self, img: torch.Tensor, target: Optional[torch.Tensor] = None, conf_thres=0.25, iou_thres=0.45,
                 max_det=300, out: Optional[Dict[str, torch.Tensor]] = None, want_mask=False, want_bits=False):
        """evaluate_model.py:134-174 in one call.  img fp32 [B,4,H,W] or uint8 [B,H,W,4].  `target`: the ground-truth masks,
        fp32 [B,1,H,W] (T = t > 0.5) or uint8 [B,H,W] / [B,1,H,W] as stored in the mask PNG (T = v >= 128).  `want_bits`:
        also return the predicted mask bit-packed (int32 [B, H*W/32], pixel i = bit i%32 of word i/32).  Returns a dict of
        device tensors (padded detections + counts); nothing is synchronised."""
        require_cuda(img, "pipeline")
        u8 = img.dtype == torch.uint8
        if u8:
            B, H, W, Cc = img.shape
            img = img.contiguous()
        else:
            img = _f32c(img)
            B, Cc, H, W = img.shape
        if Cc != 4:
            raise ValueError(f"expected 4 channels, got {Cc}")
        dev = img.device
        SH, SW = (H + 31) // 32 * 32, (W + 31) // 32 * 32
        A = sum((SH // s) * (SW // s) for s in (8, 16, 32))
        o = out if out is not None else {}

        def buf(name, shape, dtype):
            t = o.get(name)
            if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
                t = torch.empty(shape, dtype=dtype, device=dev)
                o[name] = t
            return t

        ml = buf("mask_logits", (B, 1, H, W), torch.float32)
        y = buf("y", (B, 5, A), torch.float32)
        bt = buf("bottleneck", (B, 1, H // 8, W // 8), torch.float32)
        db = buf("det_boxes", (B, max_det, 6), torch.float32)
        di = buf("det_idx", (B, max_det), torch.int64)
        dc = buf("det_count", (B,), torch.int32)
        cnt = buf("counts", (B, 3), torch.int32)
        mk = buf("mask", (B, H, W), torch.uint8) if want_mask else None
        mb = buf("mask_bits", (B, H * W // 32), torch.int32) if want_bits else None
        t_f32 = t_u8 = None
        if target is not None:
            require_cuda(target, "pipeline target")
            if target.dtype == torch.uint8:
                t_u8 = target.contiguous()
            else:
                t_f32 = _f32c(target)
        io = PipelineIO(None if u8 else img.data_ptr(), img.data_ptr() if u8 else None,
                        t_f32.data_ptr() if t_f32 is not None else None, ml.data_ptr(), y.data_ptr(), bt.data_ptr(),
                        db.data_ptr(), di.data_ptr(), dc.data_ptr(), cnt.data_ptr(),
                        mk.data_ptr() if mk is not None else None, conf_thres, iou_thres, max_det,
                        t_u8.data_ptr() if t_u8 is not None else None, mb.data_ptr() if mb is not None else None)
        ws = self._ws_for(B, H, W, max_det)
        with torch.cuda.device(self.device):
            check(lib().ysp_pipeline(self._h, C.byref(io), B, H, W, ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
        self._count()
        return o

    def profile(self, enable: int):
        """1 = on, 2 = clear + on, 0 = off (per-step CUDA-event timing inside libysp; not for timed regions)."""
        check(lib().ysp_profile(self._h, int(enable)))

    def profile_report(self):
        import json
        buf = C.create_string_buffer(1 << 20)
        n = lib().ysp_profile_report(self._h, buf, len(buf))
        if n < 0:
            check(n)
        return json.loads(buf.value.decode())

    def debug_tensor(self, name: str) -> torch.Tensor:
        shape = (C.c_int64 * 4)()
        ws = self._ws
        check(lib().ysp_debug_tensor(self._h, name.encode(), ws.data_ptr(), None, shape, None))
        out = torch.empty(*shape, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().ysp_debug_tensor(self._h, name.encode(), ws.data_ptr(), out.data_ptr(), shape, _stream_ptr(self.device)))
        return out
