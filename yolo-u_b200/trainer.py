"""Seg-head training step (BASELINE cfg 4) -- host side of the reference's `Trainer.train`
loop body (/root/reference/train.py:302-360, both branches) over libysp's `ysp_train_*` entries.

What the reference does per batch                       what happens here
  img/mask/heatmaps .float().to(device)   :317-319      tensors arrive on the device (fp32)
  optimizer.zero_grad()                    :321          inside ysp_train_step
  pred = model(img, heatmaps)              :322          frozen encoder: ysp_encoder_forward (inference engine);
                                                         decoder in train() mode (BN batch stats): ysp_train_step
  loss = DiceLoss(sigmoid, soft_label, batch=True)(pred, mask)   :323, :98-104     same call
  loss.backward()                          :326          same call (hand-derived backward, flat gradient buffer)
  clip_grad_norm_(trainable_param, 1.0)    :328          a no-op in the reference (exhausted generator):
                                                         `max_norm=0` by default, the real clip is opt-in
  optimizer.step()  (AdamW lr, defaults)   :262, :329    ysp_adamw on the flat buffers
  -- data parallel (BASELINE cfg 4) --                   ONE all-reduce(SUM) of the flat gradient buffer, 1/world folded
                                                         into the optimiser's gradient scale; BN statistics stay local
  scheduler = CosineAnnealingLR(T_max=epochs), stepped per epoch :264, :398      `scheduler_step()`
  torch.save(model.state_dict(), best.pth) :428          `state_dict()` returns the reference's keys and layouts

`mixed_precision=True` follows the branch the reference runs by default (train.py:302-341): the loss gradient is multiplied
by the GradScaler's scale (`scaler.scale(loss).backward()` :325), the flat gradient buffer is checked for inf/NaN and
unscaled inside the optimiser (`scaler.unscale_` / `scaler.step` :328-336: a non-finite gradient SKIPS the step), and
`scaler.update()` (:339) halves the scale after a skipped step and doubles it after `growth_interval` clean ones
(torch.amp.GradScaler defaults: 2**16, x2, x0.5, 2000).  What autocast changes in the reference is the arithmetic (fp16
convs); here the 1x1 convs run on tcgen05 with split fp16 / bf16 operands and fp32 accumulation and every tensor is
stored in fp32, i.e. at or above the reference's precision, so the scaler's job reduces to its control flow -- power-of-two
scaling is exact in fp32, and a clean mixed-precision step equals the plain step up to the summation order of the
gradient atomics (tested).

There is no autograd and no PyTorch fallback: parameters, gradients, Adam moments and BN running statistics are flat
fp32 device buffers whose sub-tensors are views named by the reference's state_dict keys."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Mapping, Optional

import torch

from ._lib import check, lib, require_cuda
from .engine import Engine, _f32c, _stream_ptr

LOSS_KINDS = {"dice": 0, "dice_bce": 1, "dice+bce": 1}


class SegHeadTrainer:
    def __init__(self, seg_state_dict: Mapping[str, torch.Tensor], batch_size: int, image_size=240, lr: float = 1e-4,
                 epochs: int = 100, loss: str = "dice", device="cuda:0", encoder_mode="fp32", betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2, max_norm: float = 0.0, bn_momentum: float = 0.1,
                 process_group=None, mixed_precision: bool = False, init_scale: float = 2.0 ** 16, growth_factor: float = 2.0,
                 backoff_factor: float = 0.5, growth_interval: int = 2000):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            from ._lib import YspError
            raise YspError(f"SegHeadTrainer needs a CUDA device, got {self.device}; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if loss not in LOSS_KINDS:
            raise ValueError(f"loss must be one of {sorted(LOSS_KINDS)}, got {loss!r}")
        H, W = (image_size, image_size) if isinstance(image_size, int) else image_size
        self.B, self.H, self.W = int(batch_size), int(H), int(W)
        self.lr0 = self.lr = float(lr)
        self.epochs, self.epoch, self.step_count = int(epochs), 0, 0
        self.loss_kind = LOSS_KINDS[loss]
        self.betas, self.eps, self.weight_decay, self.max_norm = betas, eps, weight_decay, max_norm
        self.bn_momentum = bn_momentum
        self.pg = process_group
        # torch.amp.GradScaler state (train.py:266); inert unless mixed_precision
        self.mixed_precision = bool(mixed_precision)
        self.scale = float(init_scale) if self.mixed_precision else 1.0
        self.growth_factor, self.backoff_factor, self.growth_interval = float(growth_factor), float(backoff_factor), int(growth_interval)
        self._growth_tracker, self.skipped_steps = 0, 0
        # frozen encoder through the inference engine (its decoder weights are only needed to finalise the handle)
        self._frozen = {k: v.detach().clone() for k, v in seg_state_dict.items() if k.startswith("encoder.") or k == "param"}
        self.engine = Engine(self.device, encoder_mode)
        self.engine.load_state_dict("seg", seg_state_dict)
        self.engine.finalize(det=False, seg=True)
        # trainer handle + flat buffers
        L = lib()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(L.ysp_train_create(C.byref(h), self.device.index, self.B, self.H, self.W))
        self._t = h
        n_p, n_s = L.ysp_train_param_count(h), L.ysp_train_stat_count(h)
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=self.device)
        self.params, self.grads, self.adam_m, self.adam_v, self.stats = z(n_p), z(n_p), z(n_p), z(n_p), z(n_s)
        self._ws = torch.empty(L.ysp_train_workspace_bytes(h), dtype=torch.uint8, device=self.device)
        self._scratch = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._loss = z(3)
        self.layout: Dict[str, tuple] = {}
        name = C.create_string_buffer(256)
        kind, off, num = C.c_int(), C.c_int64(), C.c_int64()
        for i in range(L.ysp_train_num_tensors(h)):
            check(L.ysp_train_tensor_info(h, i, name, 256, C.byref(kind), C.byref(off), C.byref(num)))
            self.layout[name.value.decode()] = (kind.value, off.value, num.value)
        self._shapes: Dict[str, torch.Size] = {}
        self._nbt: Dict[str, int] = {}
        self.load_state_dict(seg_state_dict)

    def __del__(self):
        try:
            if getattr(self, "_t", None):
                lib().ysp_train_destroy(self._t)
                self._t = None
        except Exception:
            pass

    # ---- parameters <-> state_dict (reference keys and layouts; train.py:428 / evaluate_model.py:243) ------------------
    def _view(self, buf_p: torch.Tensor, buf_s: torch.Tensor, key: str) -> torch.Tensor:
        kind, off, num = self.layout[key]
        return (buf_p if kind == 0 else buf_s)[off:off + num]

    def load_state_dict(self, sd: Mapping[str, torch.Tensor]):
        missing = [k for k in self.layout if k not in sd]
        if missing:
            raise KeyError(f"state_dict is missing {missing[:4]}{'...' if len(missing) > 4 else ''}")
        for k, (kind, off, num) in self.layout.items():
            v = sd[k]
            if v.numel() != num:
                raise ValueError(f"{k}: expected {num} elements, got shape {tuple(v.shape)}")
            self._shapes[k] = v.shape
            self._view(self.params, self.stats, k).copy_(v.detach().reshape(-1).to(self.device, torch.float32))
        for k, v in sd.items():
            if k.endswith("num_batches_tracked"):
                self._nbt[k] = int(v)
        return self

    def named_parameters(self) -> Dict[str, torch.Tensor]:
        return {k: self._view(self.params, self.stats, k).view(self._shapes[k]) for k, (kind, _, _) in self.layout.items() if kind == 0}

    def named_grads(self) -> Dict[str, torch.Tensor]:
        return {k: self._view(self.grads, self.stats, k).view(self._shapes[k]) for k, (kind, _, _) in self.layout.items() if kind == 0}

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out = {k: v.clone() for k, v in self._frozen.items()}
        for k in self.layout:
            out[k] = self._view(self.params, self.stats, k).view(self._shapes[k]).clone()
        for k, n in self._nbt.items():
            out[k] = torch.tensor(n, dtype=torch.int64)
        return out

    # ---- one batch --------------------------------------------------------------------------------------------------
    def encode(self, img: torch.Tensor):
        """Frozen encoder (YOLOSegPlusPlus.py:255-259) -> (skipA [B,H/4,W/4,64], skipB [B,H/8,W/8,128]) NHWC fp32."""
        require_cuda(img, "SegHeadTrainer")
        if tuple(img.shape) != (self.B, 4, self.H, self.W):
            raise ValueError(f"expected img {(self.B, 4, self.H, self.W)}, got {tuple(img.shape)}")
        img = _f32c(img)
        B, H, W = self.B, self.H, self.W
        skipA = torch.empty(B, H // 4, W // 4, 64, dtype=torch.float32, device=self.device)
        skipB = torch.empty(B, H // 8, W // 8, 128, dtype=torch.float32, device=self.device)
        ws = self.engine._ws_for(B, H, W)
        with torch.cuda.device(self.device):
            check(lib().ysp_encoder_forward(self.engine._h, img.data_ptr(), skipA.data_ptr(), skipB.data_ptr(), B, H, W,
                                            ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
        self._enc_launches = lib().ysp_last_launch_count(self.engine._h)
        return skipA, skipB

    def _encode_ahead(self, img: torch.Tensor, ready: "torch.cuda.Event"):
        """The frozen encoder of a LATER batch on a side stream (it does not depend on the weights being trained), ordered
        after `ready`; the skips are handed to the step that trains on `img`."""
        if getattr(self, "_enc_stream", None) is None:
            self._enc_stream = torch.cuda.Stream(self.device)
        st = self._enc_stream
        st.wait_event(ready)
        with torch.cuda.stream(st):
            skips = self.encode(img)
            done = torch.cuda.Event()
            done.record(st)
        img.record_stream(st)
        self._ahead = (img, skips, done)

    def forward_backward(self, img: torch.Tensor, mask: torch.Tensor, heatmaps: torch.Tensor, grad_scale: float = 1.0,
                         want_pred: bool = True, next_img: Optional[torch.Tensor] = None):
        """zero_grad + forward (train mode) + loss + backward.  Returns (loss3 device tensor {total, dice, bce}, pred).
        `next_img`: the images of the NEXT batch (already resident on the device when this call starts): their frozen-encoder
        pass is launched on a side stream so it runs under this step's decoder kernels instead of in front of the next step
        (measured at B = 128: 15.16 -> 15.04 ms per step, with occasional 17 ms runs when the two streams' persistent kernels
        interleave badly -- an option, not the default of bench.py)."""
        cur = torch.cuda.current_stream(self.device)
        ahead = getattr(self, "_ahead", None)
        self._ahead = None
        if ahead is not None and ahead[0] is img:
            skipA, skipB = ahead[1]
            cur.wait_event(ahead[2])
            skipA.record_stream(cur)
            skipB.record_stream(cur)
        else:
            skipA, skipB = self.encode(img)
        ready = None
        if next_img is not None:
            # the side stream starts behind everything queued so far -- next_img's producer, and an encoder pass that had to
            # run on this stream (the two would share the encoder engine's workspace) -- but NOT behind this step's decoder
            ready = torch.cuda.Event()
            ready.record(cur)
        B, H, W = self.B, self.H, self.W
        if tuple(mask.shape) != (B, 1, H, W):
            raise ValueError(f"expected mask {(B, 1, H, W)}, got {tuple(mask.shape)}")
        if tuple(heatmaps.shape) != (B, 1, H // 8, W // 8):
            raise RuntimeError(f"Sizes of tensors must match: heatmaps {tuple(heatmaps.shape)} vs expected {(B, 1, H // 8, W // 8)}")
        require_cuda(mask, "SegHeadTrainer")
        require_cuda(heatmaps, "SegHeadTrainer")
        mask, heatmaps = _f32c(mask), _f32c(heatmaps)
        pred = torch.empty(B, 1, H, W, dtype=torch.float32, device=self.device) if want_pred else None
        with torch.cuda.device(self.device):
            check(lib().ysp_train_step(self._t, skipA.data_ptr(), skipB.data_ptr(), heatmaps.data_ptr(), mask.data_ptr(),
                                       self.params.data_ptr(), self.grads.data_ptr(), self.stats.data_ptr(),
                                       self.bn_momentum, self.loss_kind, grad_scale, self._loss.data_ptr(),
                                       pred.data_ptr() if want_pred else None, self._ws.data_ptr(), self._ws.numel(),
                                       _stream_ptr(self.device)))
        for k in self._nbt:
            self._nbt[k] += 1
        if next_img is not None:
            self._encode_ahead(next_img, ready)
        return self._loss, pred

    def optimizer_step(self) -> bool:
        """All-reduce (if a process group is up) + AdamW on the flat buffers.  With mixed_precision this is
        scaler.unscale_ + scaler.step + scaler.update (train.py:328-339): returns False when the step was skipped because a
        gradient was inf/NaN (one 8-byte D2H per step, the same host sync torch's scaler.step makes)."""
        import torch.distributed as dist
        world = 1
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.pg)
            if world > 1:
                dist.all_reduce(self.grads, op=dist.ReduceOp.SUM, group=self.pg)
        if self.mixed_precision:
            with torch.cuda.device(self.device):
                check(lib().ysp_grad_sqnorm(self.grads.data_ptr(), self.grads.numel(), self._scratch.data_ptr(),
                                            _stream_ptr(self.device)))
            if not math.isfinite(float(self._scratch.item())):       # every rank sees the same all-reduced buffer
                self.scale *= self.backoff_factor
                self._growth_tracker = 0
                self.skipped_steps += 1
                return False
        self.step_count += 1
        with torch.cuda.device(self.device):
            check(lib().ysp_adamw(self.params.data_ptr(), self.grads.data_ptr(), self.adam_m.data_ptr(), self.adam_v.data_ptr(),
                                  self.params.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                  self.step_count, 1.0 / (world * self.scale), self.max_norm, self._scratch.data_ptr(),
                                  _stream_ptr(self.device)))
        if self.mixed_precision:
            self._growth_tracker += 1
            if self._growth_tracker == self.growth_interval:
                self.scale *= self.growth_factor
                self._growth_tracker = 0
        return True

    def step(self, img, mask, heatmaps, want_pred: bool = False, next_img: Optional[torch.Tensor] = None):
        """One iteration of train.py:317-329.  Returns the device loss tensor {total, dice, bce} (no sync) and pred.
        `next_img` (optional): the next batch's images, see `forward_backward`."""
        loss, pred = self.forward_backward(img, mask, heatmaps, grad_scale=self.scale, want_pred=want_pred, next_img=next_img)
        self.optimizer_step()
        return loss, pred

    @torch.no_grad()
    def validate_batch(self, img: torch.Tensor, mask: torch.Tensor, heatmaps: torch.Tensor, mode: str = "fp32"):
        """One iteration of the validation loop (train.py:346-366): model.eval() forward (BN running statistics) through
        the inference engine, the loss value, and the integer counters (|P&T|, |P|, |T| per slice) that Dice, precision
        and recall are made of.  Returns (loss3 device tensor {total, dice, bce}, counts int32 [B,3], pred logits)."""
        if getattr(self, "_eval_step", None) != (self.step_count, mode):
            self._eval = self.eval_engine(mode)               # weights changed since the last validation pass
            self._eval_step = (self.step_count, mode)
        require_cuda(img, "SegHeadTrainer")
        pred = self._eval.segpp_forward(_f32c(img), _f32c(heatmaps))
        mask = _f32c(mask)
        B, _, H, W = pred.shape
        loss3 = torch.empty(3, dtype=torch.float32, device=self.device)
        counts = torch.empty(B, 3, dtype=torch.int32, device=self.device)
        ws = torch.empty(4, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            st = _stream_ptr(self.device)
            check(lib().ysp_seg_loss(pred.data_ptr(), mask.data_ptr(), pred.numel(), self.loss_kind, loss3.data_ptr(),
                                     ws.data_ptr(), st))
            check(lib().ysp_mask_dice(pred.data_ptr(), mask.data_ptr(), B, H * W, counts.data_ptr(), None, st))
        return loss3, counts, pred

    def scheduler_step(self):
        """CosineAnnealingLR(T_max=epochs, eta_min=0).step() (train.py:264, :398), closed form."""
        self.epoch += 1
        self.lr = 0.5 * self.lr0 * (1.0 + math.cos(math.pi * self.epoch / self.epochs))
        return self.lr

    def eval_engine(self, mode="fp32") -> Engine:
        """A fresh inference engine over the CURRENT weights (model.eval(): BN running statistics) -- the validation
        half of the epoch (train.py:343-372) and evaluate_model.py run through it."""
        e = Engine(self.device, mode)
        e.load_state_dict("seg", self.state_dict())
        e.finalize(det=False, seg=True)
        return e

    @property
    def bytes_per_step(self) -> float:
        """ALGORITHMIC HBM bytes of the decoder forward + backward of the last step (libysp's own accounting)"""
        return lib().ysp_train_last_step_bytes(self._t)

    @property
    def launches_per_step(self) -> int:
        """kernels of this library one `step()` launches: frozen encoder + decoder forward/backward + AdamW"""
        return lib().ysp_train_last_launch_count(self._t) + getattr(self, "_enc_launches", 0) + 1
