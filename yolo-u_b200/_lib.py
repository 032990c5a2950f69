"""ctypes binding of libysp.so (include/ysp.h).  There is NO fallback: if the library is missing or there is no CUDA
device, calls raise -- the product path never routes through PyTorch eager or the CPU oracle."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libysp.so")

MODE_FP32, MODE_BF16, MODE_TC32 = 0, 1, 2
# "fp32": fp32 storage, CUDA-core FFMA convs.  "tc32" / "parity": fp32 storage, convs on tcgen05 with fp16 hi/lo operand
# splits (3 MMAs per product, fp32-accurate: the mode that meets the 1e-3 logit tolerance ON tensor cores).
# "bf16" / "throughput": bf16 storage, single tcgen05 bf16 MMA.
_MODES = {"fp32": MODE_FP32, "float32": MODE_FP32, "ffma": MODE_FP32, "tc32": MODE_TC32, "parity": MODE_TC32,
          "fp16x3": MODE_TC32, "bf16": MODE_BF16, "bfloat16": MODE_BF16, "throughput": MODE_BF16}

vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t


class PipelineIO(C.Structure):
    _fields_ = [("d_img", vp), ("d_img_u8", vp), ("d_target", vp), ("d_mask_logits", vp), ("d_y", vp),
                ("d_bottleneck", vp), ("d_det_boxes", vp), ("d_det_idx", vp), ("d_det_count", vp), ("d_counts", vp),
                ("d_mask", vp), ("conf_thres", f32), ("iou_thres", f32), ("max_det", i32), ("d_target_u8", vp),
                ("d_mask_bits", vp)]


# name -> (restype, argtypes); every symbol include/ysp.h declares
PROTOTYPES = {
    "ysp_version": (i32, []),
    "ysp_last_error": (C.c_char_p, []),
    "ysp_create": (i32, [C.POINTER(vp), i32, i32]),
    "ysp_destroy": (None, [vp]),
    "ysp_load_weight": (i32, [vp, C.c_char_p, vp, i32, C.POINTER(i64)]),
    "ysp_finalize": (i32, [vp, i32]),
    "ysp_workspace_bytes": (sz, [vp, i32, i32, i32]),
    "ysp_pipeline_workspace_bytes": (sz, [vp, i32, i32, i32, i32]),
    "ysp_normalize_u8": (i32, [vp, vp, i32, i32, i32, vp]),
    "ysp_detector_forward": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp]),
    "ysp_bottleneck": (i32, [vp, i32, i32, i32, i32, vp, i32, i32, vp]),
    "ysp_segpp_forward": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, sz, vp]),
    "ysp_nms_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "ysp_nms": (i32, [vp, i32, i32, i32, i32, f32, f32, i32, i32, f32, i32, vp, i32, vp, vp, vp, vp, sz, vp]),
    "ysp_nms_core": (i32, [vp, vp, i32, f32, vp, vp, vp, sz, vp]),
    "ysp_xywh2xyxy_inplace": (i32, [vp, i32, i32, i32, vp]),
    "ysp_mask_dice": (i32, [vp, vp, i32, i32, vp, vp, vp]),
    "ysp_objectmap_transform": (i32, [vp, vp, i32, i32, vp]),
    "ysp_scale_boxes": (i32, [vp, C.c_longlong, i32, f32, f32, f32, f32, f32, vp]),
    "ysp_pipeline": (i32, [vp, C.POINTER(PipelineIO), i32, i32, i32, vp, sz, vp]),
    "ysp_conf_gate": (i32, [vp, vp, i32, i32, i32, f32, vp, vp, i32, vp, vp]),
    "ysp_resize_u8": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "ysp_train_create": (i32, [C.POINTER(vp), i32, i32, i32, i32]),
    "ysp_train_destroy": (None, [vp]),
    "ysp_train_num_tensors": (i32, [vp]),
    "ysp_train_tensor_info": (i32, [vp, i32, C.c_char_p, i32, C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]),
    "ysp_train_param_count": (i64, [vp]),
    "ysp_train_stat_count": (i64, [vp]),
    "ysp_train_workspace_bytes": (sz, [vp]),
    "ysp_train_last_launch_count": (i32, [vp]),
    "ysp_train_last_step_bytes": (C.c_double, [vp]),
    "ysp_encoder_forward": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, sz, vp]),
    "ysp_train_step": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, f32, i32, f32, vp, vp, vp, sz, vp]),
    "ysp_seg_loss": (i32, [vp, vp, i64, i32, vp, vp, vp]),
    "ysp_grad_sqnorm": (i32, [vp, i64, vp, vp]),
    "ysp_adamw": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, f32, f32, vp, vp]),
    "ysp_last_launch_count": (i32, [vp]),
    "ysp_set_keep_intermediates": (i32, [vp, i32]),
    "ysp_profile": (i32, [vp, i32]),
    "ysp_profile_report": (i32, [vp, C.c_char_p, sz]),
    "ysp_debug_tensor": (i32, [vp, C.c_char_p, vp, vp, C.POINTER(i64), vp]),
}

_lib = None


class YspError(RuntimeError):
    pass


def lib():
    """Load libysp.so (built in-tree by build.py / __graft_entry__.build()).  Raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise YspError(f"{SO} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                           "There is no CPU / PyTorch fallback for this path.")
        L = C.CDLL(SO)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def mode_id(mode) -> int:
    if isinstance(mode, int):
        return mode
    try:
        return _MODES[str(mode).lower()]
    except KeyError:
        raise ValueError(f"mode must be one of {sorted(_MODES)}, got {mode!r}")


def check(rc: int):
    """Turn a YSP_E* status into the Python exception the reference would have raised."""
    if rc == 0:
        return
    msg = lib().ysp_last_error().decode(errors="replace")
    if rc == -1:
        if msg.startswith("Invalid "):          # nms.py:59-60 are `assert`s
            raise AssertionError(msg)
        raise ValueError(msg)
    if rc == -4:
        raise KeyError(msg)
    raise YspError(f"libysp error {rc}: {msg}")


def require_cuda(t, what: str):
    if not t.is_cuda:
        raise YspError(f"{what}: tensor is on {t.device}; this B200-native path has no CPU fallback")
