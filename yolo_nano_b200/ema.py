"""`ModelEMA` of the reference (utils/misc.py:67-86; `train.py --ema`, :147,233-235) on the device.

Same constructor, attributes (`ema`, `updates`, `decay`) and `update(model)` as the reference class; the update of
ALL floating-point state tensors (247 parameters + 148 BatchNorm statistics of YOLO-Nano) is ONE kernel launch
(`ynb_ema_update`) instead of two ATen launches per tensor, with the reference's arithmetic
(`v *= d; v += (1 - d) * m`, both scalars applied as float32) — bit-identical results.
No CPU path: the model must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import math
from copy import deepcopy

import torch

from . import _lib
from .engine import EngineError


def _is_parallel(model) -> bool:       # utils/misc.py: is_parallel
    return type(model) in (torch.nn.parallel.DataParallel, torch.nn.parallel.DistributedDataParallel)


class ModelEMA:
    def __init__(self, model, decay: float = 0.9999, updates: int = 0):
        self.ema = deepcopy(model.module if _is_parallel(model) else model).eval()     # FP32 EMA
        self.updates = updates
        self.decay = lambda x: decay * (1 - math.exp(-x / 2000.))
        for p in self.ema.parameters():
            p.requires_grad_(False)
        self._tables = None
        self._key = None

    def _build_tables(self, msd, esd):
        lib = _lib.load()
        chunk = int(lib.ynb_ema_chunk_elems())
        e_ptrs, m_ptrs, sizes, c_t, c_i = [], [], [], [], []
        dev = None
        for k, v in esd.items():
            if not v.dtype.is_floating_point:
                continue                                       # num_batches_tracked (utils/misc.py:84)
            m = msd[k]
            if v.dtype != torch.float32 or m.dtype != torch.float32 or not v.is_cuda or not m.is_cuda \
                    or v.device != m.device or not v.is_contiguous() or not m.is_contiguous() or v.shape != m.shape:
                raise EngineError(f"ModelEMA: {k} must be a contiguous float32 CUDA tensor in both models "
                                  "(no CPU fallback)")
            dev = v.device
            t = len(sizes)
            e_ptrs.append(v.data_ptr()); m_ptrs.append(m.data_ptr()); sizes.append(v.numel())
            for i in range((v.numel() + chunk - 1) // chunk):
                c_t.append(t); c_i.append(i)
        if dev is None:
            raise EngineError("ModelEMA: the model has no floating-point state")
        u64 = lambda a: torch.tensor(a, dtype=torch.int64).to(dev)      # addresses < 2^63
        self._tables = (u64(e_ptrs), u64(m_ptrs), u64(sizes), torch.tensor(c_t, dtype=torch.int32).to(dev),
                        torch.tensor(c_i, dtype=torch.int32).to(dev), len(c_t), dev)

    def update(self, model):
        self.updates += 1
        d = self.decay(self.updates)
        msd = (model.module if _is_parallel(model) else model).state_dict(keep_vars=True)
        esd = self.ema.state_dict(keep_vars=True)
        key = tuple(v.data_ptr() for v in esd.values()) + tuple(v.data_ptr() for v in msd.values())
        if key != self._key:                                   # storage moved (.to(), load_state_dict into new tensors)
            self._build_tables(msd, esd)
            self._key = key
        e, m, s, ct, ci, n, dev = self._tables
        f32 = lambda x: C.c_float(x)
        rc = _lib.load().ynb_ema_update(C.c_void_p(e.data_ptr()), C.c_void_p(m.data_ptr()), C.c_void_p(s.data_ptr()),
                                        C.c_void_p(ct.data_ptr()), C.c_void_p(ci.data_ptr()), n, f32(d), f32(1. - d),
                                        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        if rc != _lib.YNB_OK:
            raise EngineError(f"ynb_ema_update failed ({rc})")
        # the tensors were written behind autograd's back (no version bump): tell the drop-in detector to re-pack
        # its engine weights before the next forward of the EMA model (eval.py / train.py evaluate `ema.ema`)
        if hasattr(self.ema, "mark_weights_dirty"):
            self.ema.mark_weights_dirty()
