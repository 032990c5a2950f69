"""ORACLE (test infrastructure) -- ctypes loader for the plain-C restatement oracle/nms_oracle.c."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libysp_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nms_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.ysp_oracle_nms.restype = ctypes.c_int
        _lib.ysp_oracle_mask_counts.restype = ctypes.c_int
    return _lib


def nms_batched(pred: torch.Tensor, conf_thres=0.25, iou_thres=0.45, max_det=300, max_nms=30000, max_wh=7680,
                agnostic=False, nc=0, nthreads=1):
    """pred [B,C,A] fp32 CPU (NOT modified) -> (list of [n,6] boxes, list of int64 [n] keep idx)."""
    p = np.ascontiguousarray(pred.detach().cpu().numpy().astype(np.float32))
    B, C, A = p.shape
    ob = np.zeros((B, max_det, 6), np.float32)
    oi = np.zeros((B, max_det), np.int64)
    oc = np.zeros((B,), np.int32)
    rc = lib().ysp_oracle_nms(p.ctypes.data_as(ctypes.c_void_p), B, C, A, int(nc), ctypes.c_float(conf_thres),
                              ctypes.c_float(iou_thres), int(max_det), int(max_nms), ctypes.c_float(max_wh),
                              int(bool(agnostic)), ob.ctypes.data_as(ctypes.c_void_p),
                              oi.ctypes.data_as(ctypes.c_void_p), oc.ctypes.data_as(ctypes.c_void_p), int(nthreads))
    assert rc == 0
    return ([torch.from_numpy(ob[b, :oc[b]].copy()) for b in range(B)],
            [torch.from_numpy(oi[b, :oc[b]].copy()) for b in range(B)])


def mask_counts(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    lg = np.ascontiguousarray(logits.detach().cpu().numpy().astype(np.float32))
    tg = np.ascontiguousarray(target.detach().cpu().numpy().astype(np.float32))
    B = lg.shape[0]
    HW = lg.size // B
    out = np.zeros((B, 3), np.int32)
    lib().ysp_oracle_mask_counts(lg.ctypes.data_as(ctypes.c_void_p), tg.ctypes.data_as(ctypes.c_void_p), B, HW,
                                 out.ctypes.data_as(ctypes.c_void_p))
    return torch.from_numpy(out)
