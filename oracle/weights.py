"""Deterministic weight sets for parity tests (test infrastructure, see
yolo_nano_oracle.py header).

* `reference_init`: what `YOLONano(...)` of the reference produces under
  `torch.manual_seed(seed)` (backbone/shufflenetv2.py:131-154 + torch defaults).  Built
  through this repo's parameter containers, which consume the RNG in the same order;
  `oracle/gen_golden.py` asserts equality with the real reference and stores a digest.
* `calibrated`: same shapes, but BatchNorm affine parameters randomised and running
  statistics measured on a calibration batch, so that every layer has O(1) activations,
  BN folding is far from identity and scores are tie-free.  Needed because the reference
  init collapses deep activations to ~1e-5 (SURVEY §8c hazard 1), which would make an
  end-to-end tolerance blind to bugs in stage3/4.
"""
from __future__ import annotations

import hashlib
from typing import Dict

import torch

from . import yolo_nano_oracle as oracle

StateDict = Dict[str, torch.Tensor]


def anchors_for(num_classes: int):
    import yolo_nano_b200 as pkg
    return pkg.MULTI_ANCHOR_SIZE_COCO if num_classes == 80 else pkg.MULTI_ANCHOR_SIZE


def reference_init(num_classes: int, seed: int = 0) -> StateDict:
    import contextlib, io
    import yolo_nano_b200 as pkg
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = pkg.YOLONano(torch.device("cpu"), 416, num_classes, anchor_size=anchors_for(num_classes))
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def calibrated(num_classes: int, seed: int = 0, calib_size: int = 160, calib_batch: int = 2) -> StateDict:
    sd = reference_init(num_classes, seed)
    g = torch.Generator().manual_seed(seed + 1000)
    for k in sorted(sd.keys()):
        v = sd[k]
        if k.endswith("running_mean"):
            p = k[: -len("running_mean")]
            c = v.numel()
            sd[p + "weight"] = torch.empty(c).uniform_(0.5, 1.5, generator=g)
            sd[p + "bias"] = torch.randn(c, generator=g) * 0.3
        elif k.endswith(".convs.0.bias"):
            sd[k] = torch.randn(v.numel(), generator=g) * 0.1
    for hd in (1, 2, 3):      # final 1x1 convs have no BN: widen their logits directly
        sd[f"head_det_{hd}.4.weight"] = sd[f"head_det_{hd}.4.weight"] * 3.0
        sd[f"head_det_{hd}.4.bias"] = torch.randn(sd[f"head_det_{hd}.4.bias"].numel(), generator=g)
    x = torch.randn(calib_batch, 3, calib_size, calib_size, generator=g)
    oracle._CALIBRATE = True
    try:
        oracle.network(sd, x)
    finally:
        oracle._CALIBRATE = False
    return sd


def synthetic_input(batch: int, input_size: int, seed: int = 0) -> torch.Tensor:
    """ImageNet-normalised images are ~N(0,1) (data/transforms.py:65-68)."""
    g = torch.Generator().manual_seed(seed + 7)
    return torch.randn(batch, 3, input_size, input_size, generator=g)


def digest(sd_or_tensor) -> str:
    h = hashlib.sha256()
    if isinstance(sd_or_tensor, torch.Tensor):
        h.update(sd_or_tensor.contiguous().numpy().tobytes())
    else:
        for k in sorted(sd_or_tensor.keys()):
            h.update(k.encode())
            h.update(sd_or_tensor[k].contiguous().numpy().tobytes())
    return h.hexdigest()[:16]
