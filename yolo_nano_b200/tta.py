"""Test-time augmentation driver — drop-in for `utils.misc.TestTimeAugmentation` (utils/misc.py:90-148).

Multi-scale + horizontal flip through the engine, then a second per-class NMS over the union.
Everything that computes runs in the library: the bilinear resize and flip (`ynb_resize_bilinear`),
the network / decode / NMS per scale (`YOLONano.detect`, original and flipped image as ONE batch of
two) and the merging NMS (`ynb_nms`, the same exact-IoU kernel, threshold only).  Host work is
concatenation and the box mirror `x -> 1 - x`, as in the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import EngineError, _ptr, _stream_ptr


def resize_bilinear(x: torch.Tensor, size: int, with_flip: bool = False) -> torch.Tensor:
    """F.interpolate(x, (size, size), mode='bilinear', align_corners=False) [+ torch.flip(.., [-1])] on the
    device: x [B,3,H,W] float32 CUDA -> [B or 2B, 3, size, size] (out[2b] resized, out[2b+1] flipped)."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3:
        raise EngineError("resize_bilinear wants a float32 CUDA tensor [B,3,H,W]")
    x = x.contiguous()
    b = x.shape[0]
    out = torch.empty((b * (2 if with_flip else 1), 3, size, size), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    rc = lib.ynb_resize_bilinear(_ptr(x), b, x.shape[2], x.shape[3], _ptr(out), int(size), int(with_flip),
                                 _stream_ptr(x.device))
    if rc != 0:
        raise EngineError(f"ynb_resize_bilinear failed ({rc}): {lib.ynb_last_error(None).decode()}")
    return out


def merge_nms(bboxes: np.ndarray, scores: np.ndarray, labels: np.ndarray, num_classes: int, nms_thresh: float,
              device: torch.device):
    """Per-class greedy NMS over the union of all augmented detections (utils/misc.py:131-146 with
    utils/misc.py:8-37): keeps rows in ascending index order, no score threshold."""
    n = len(scores)
    if n == 0:
        return bboxes, scores, labels
    if n > 65536:
        raise EngineError("merge_nms: more than 65536 detections to merge")
    lib = _lib.load()
    d_b = torch.from_numpy(np.ascontiguousarray(bboxes, dtype=np.float32)).to(device).view(1, n, 4)
    d_s = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32)).to(device).view(1, n)
    d_c = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int32)).to(device).view(1, n)
    ob = torch.empty_like(d_b)
    os_ = torch.empty_like(d_s)
    oc = torch.empty_like(d_c)
    on = torch.zeros((1,), dtype=torch.int32, device=device)
    wsb = lib.ynb_nms_workspace_bytes(1, n)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=device)
    lo = float(min(0.0, float(scores.min())))          # the reference applies no threshold here
    rc = lib.ynb_nms(_ptr(d_b), _ptr(d_s), _ptr(d_c), 1, n, int(num_classes), lo, float(nms_thresh), 0,
                     _ptr(ob), _ptr(os_), _ptr(oc), _ptr(on), None, _ptr(ws), wsb, _stream_ptr(device))
    if rc != 0:
        raise EngineError(f"ynb_nms failed ({rc}): {lib.ynb_last_error(None).decode()}")
    k = int(on.item())
    return (ob[0, :k].cpu().numpy(), os_[0, :k].cpu().numpy(), oc[0, :k].cpu().numpy().astype(np.int64))


class TestTimeAugmentation(object):
    """Same constructor and call as utils.misc.TestTimeAugmentation."""
    __test__ = False     # not a pytest class

    def __init__(self, num_classes=80, nms_thresh=0.4, scale_range=(320, 640, 32)):
        self.num_classes = num_classes
        self.nms_thresh = nms_thresh
        self.scales = np.arange(scale_range[0], scale_range[1] + 1, scale_range[2])

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, model):
        # x: [B, C, H, W]; like the reference model, only image 0 is detected (models/yolo_nano.py:365-367)
        bboxes_list, scores_list, labels_list = [], [], []
        x0 = x[:1].contiguous()
        for s in self.scales:
            s = int(s)
            # [2,3,s,s]: original, flipped.  At s == input size the reference skips the resize (:104-105);
            # the kernel's weights are then exactly (1, 0), i.e. a bit-exact copy.
            pair = resize_bilinear(x0, s, with_flip=True)
            model.set_grid(s)
            (b0, s0, l0), (b1, s1, l1) = model.detect(pair)
            b1 = b1.copy()
            b1[:, 0::2] = 1.0 - b1[:, 2::-2]                       # mirror back (:124-125)
            bboxes_list += [b0, b1]
            scores_list += [s0, s1]
            labels_list += [l0, l1]
        bboxes = np.concatenate(bboxes_list)
        scores = np.concatenate(scores_list)
        labels = np.concatenate(labels_list)
        return merge_nms(bboxes, scores, labels, self.num_classes, self.nms_thresh, x.device)
