"""Shared helpers for the GPU parity tests (the oracle is the checker, never the path)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from oracle import weights as W
from oracle import yolo_nano_oracle as O
from yolo_nano_b200 import _lib
from yolo_nano_b200.engine import Engine
from yolo_nano_b200.topology import conv_table

DEV = torch.device("cuda", 0)
MODES = {"ffma": _lib.GEMM_FP32_FFMA, "3xtf32": _lib.GEMM_TC_3XTF32, "tf32": _lib.GEMM_TC_TF32,
         "bf16": _lib.GEMM_TC_BF16}


def make_engine(sd, size, classes, mode="3xtf32", max_batch=2, **kw) -> Engine:
    eng = Engine(DEV, size, classes, W.anchors_for(classes), gemm_mode=MODES[mode], max_batch=max_batch, **kw)
    eng.load_weights(O.fold_state_dict(sd, conv_table(classes)))
    return eng


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b| relative to the tensor's own scale (SURVEY §8c hazard 1)."""
    scale = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / scale)


def ptr(t):
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream(DEV).cuda_stream)


def run_nms(lib, boxes, scores, cls, classes, conf=0.001, thr=0.5, diou=False, grid_size=0):
    """boxes [B,N,4], scores [B,N], cls [B,N] numpy -> list of kept anchor index arrays + outputs.
    grid_size > 0: the anchor-grid NMS (ynb_nms_grid) for that input size instead of ynb_nms."""
    b, n = scores.shape
    d_boxes = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.float32)).to(DEV)
    d_scores = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32)).to(DEV)
    d_cls = torch.from_numpy(np.ascontiguousarray(cls, dtype=np.int32)).to(DEV)
    ob = torch.zeros((b, n, 4), device=DEV)
    os_ = torch.zeros((b, n), device=DEV)
    oc = torch.zeros((b, n), device=DEV, dtype=torch.int32)
    on = torch.zeros((b,), device=DEV, dtype=torch.int32)
    keep = torch.zeros((b, n), device=DEV, dtype=torch.uint8)
    if grid_size:
        assert n == sum(3 * (grid_size // s) ** 2 for s in (8, 16, 32))
        wsb = lib.ynb_nms_grid_workspace_bytes(b, grid_size)
        ws = torch.empty((wsb,), device=DEV, dtype=torch.uint8)
        rc = lib.ynb_nms_grid(ptr(d_boxes), ptr(d_scores), ptr(d_cls), b, grid_size, classes, conf, thr, int(diou),
                              ptr(ob), ptr(os_), ptr(oc), ptr(on), ptr(keep), ptr(ws), wsb, stream())
    else:
        wsb = lib.ynb_nms_workspace_bytes(b, n)
        ws = torch.empty((wsb,), device=DEV, dtype=torch.uint8)
        rc = lib.ynb_nms(ptr(d_boxes), ptr(d_scores), ptr(d_cls), b, n, classes, conf, thr, int(diou),
                         ptr(ob), ptr(os_), ptr(oc), ptr(on), ptr(keep), ptr(ws), wsb, stream())
    assert rc == 0, lib.ynb_last_error(None)
    torch.cuda.synchronize()
    keep_h = keep.cpu().numpy()
    counts = on.cpu().numpy()
    kept = [np.nonzero(keep_h[i])[0] for i in range(b)]
    return kept, counts, ob.cpu().numpy(), os_.cpu().numpy(), oc.cpu().numpy()
