"""The chained training step of BASELINE config 5: `model.train()(x, target)` of the reference
(`models/yolo_nano.py:282-358` with BatchNorm on BATCH statistics) + `total_loss.backward()` (`train.py:222-229`),
every arithmetic step a kernel of libyolonano_b200.so:

    stem conv -> BN(batch) -> ReLU -> max-pool -> 16 ShuffleV2 units -> laterals -> FPN/PAN merges + 3x3 smooth convs
    -> three heads -> loss + d loss / d head maps (train_loss_kernel) -> the same chain backwards

A tape records one closure per forward op; `backward()` replays them in reverse, accumulating the gradient of a tensor
that feeds several consumers with `ynb_add`.  PyTorch moves data only (views, weight-layout transposes, padded
parameter copies of the 58-channel layers): no torch arithmetic on activations or gradients; `chunk` / `cat` /
`channel_shuffle` and their adjoints are one-launch copies (`ynb_shuffle_unit_move`).  The convs run on the
asynchronous tensor-core entries (weights split into hi / lo planes by a kernel on the same stream); the fused
inference kernels do not apply in training mode (batch statistics need the pre-BN conv outputs).  `Trainer` replays
the whole forward + backward from a CUDA graph.

    step = TrainStep(model)                       # model: yolo_nano_b200.YOLONano on a CUDA device
    losses, grads = step.forward_backward(x, target)          # grads: {parameter name: tensor, reference shapes}
    flat = step.flat_gradient(grads)              # parameter order = model.parameters() (train.py:167)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch

from . import _lib
from . import training as T
from .engine import EngineError, _ptr, _stream_ptr
from .topology import STAGE_CHANNELS, STAGE_REPEATS

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class _Tape:
    def __init__(self, dev):
        self.dev = dev
        self.ops: List = []
        self.grad: Dict[int, torch.Tensor] = {}
        self.keep: List[torch.Tensor] = []          # tensors whose id() keys self.grad must stay alive
        self.pgrad: Dict[str, torch.Tensor] = {}

    def add_grad(self, t: torch.Tensor, g: torch.Tensor):
        g = g.contiguous()
        k = id(t)
        if k in self.grad:
            a = self.grad[k]
            out = torch.empty_like(a)
            T._check(_lib.load().ynb_add(_ptr(a), _ptr(g), _ptr(out), a.numel(), _stream_ptr(self.dev)), "ynb_add")
            self.grad[k] = out
        else:
            self.grad[k] = g
        self.keep.append(t)

    def take(self, t: torch.Tensor) -> torch.Tensor:
        return self.grad.pop(id(t))

    def backward(self):
        for fn in reversed(self.ops):
            fn()


class TrainStep:
    """Forward + backward of the whole detector in training mode.  Holds padded working copies of the BatchNorm
    running statistics and writes them back to the model's buffers after every step (as nn.BatchNorm2d does)."""

    def __init__(self, model):
        self.model = model
        p = next(model.parameters())
        if p.device.type != "cuda":
            raise EngineError("TrainStep needs the model on a CUDA device (no CPU fallback)")
        self.dev = p.device
        self.lib = _lib.load()
        # debug: collect the pipeline-timeout flag of every tensor-core GEMM of a step and check them at its end
        self.check_flags = False
        self.flags = None

    # ---- kernel wrappers (NHWC float32, dense) ------------------------------------------------------------------
    def _st(self):
        return _stream_ptr(self.dev)

    # ---- ops: forward now, closure for backward -----------------------------------------------------------------
    def _param(self, name):
        return self.sd[name]

    def pw(self, tape, x, conv, bias: bool):
        """1x1 conv on [B,H,W,K] (K, N padded to multiples of 4: zero weights, packed by the GEMM's own pack kernel)."""
        w = self._param(conv + ".weight")
        n, k = w.shape[0], w.shape[1]
        w2 = w.reshape(n, k)
        np_, kp = _pad4(n), x.shape[-1]
        b = self._param(conv + ".bias") if bias else None
        shp = x.shape
        x2 = x.reshape(-1, kp)
        y = T.pwconv_forward(x2, w2, b, cout=np_, flags=self.flags).reshape(*shp[:-1], np_)

        def bw():
            dy = tape.take(y).reshape(-1, np_)
            dw, db = T.pwconv_backward_weight(dy, x2)
            tape.pgrad[conv + ".weight"] = dw[:n, :k].reshape(n, k, 1, 1).contiguous()
            if bias:
                tape.pgrad[conv + ".bias"] = db[:n].contiguous()
            tape.add_grad(x, T.pwconv_forward(dy, w2, None, transposed=True, cout=kp, flags=self.flags).reshape(shp))
        tape.ops.append(bw)
        return y

    def dw(self, tape, x, conv, stride: int, bias: bool):
        """depthwise 3x3 on [B,H,W,C] (C padded)."""
        w = self._param(conv + ".weight")                        # [C,1,3,3]
        c, cp = w.shape[0], x.shape[-1]
        w9 = torch.zeros((9, cp), device=self.dev)
        w9[:, :c] = w.reshape(c, 9).t()
        bp = torch.zeros(cp, device=self.dev)
        if bias:
            bp[:c] = self._param(conv + ".bias")
        b_, h, w_, _ = x.shape
        ho, wo = (h - 1) // stride + 1, (w_ - 1) // stride + 1
        y = torch.empty((b_, ho, wo, cp), device=self.dev)
        T._check(self.lib.ynb_dwconv3x3(_ptr(x), cp, 0, _ptr(y), cp, 0, 1, _ptr(w9), _ptr(bp), b_, h, w_, cp, stride, 0,
                                        self._st()), "ynb_dwconv3x3")

        def bw():
            dy = tape.take(y)
            dx, dw9, db = T.dwconv3x3_backward(dy, x, w9, stride)
            tape.pgrad[conv + ".weight"] = dw9[:, :c].t().reshape(c, 1, 3, 3).contiguous()
            if bias:
                tape.pgrad[conv + ".bias"] = db[:c].contiguous()
            tape.add_grad(x, dx)
        tape.ops.append(bw)
        return y

    def conv3(self, tape, x, conv):
        """dense 3x3, pad 1, stride 1, with bias (the smooth convs)."""
        w = self._param(conv + ".weight")                        # [N,K,3,3]
        n, k = w.shape[0], w.shape[1]
        wf = w.permute(0, 2, 3, 1).reshape(n, 9, k).contiguous()            # [n][t][k]
        bias = self._param(conv + ".bias").contiguous()
        b_, h, w_, _ = x.shape
        y = T.conv3x3_forward(x, wf, bias, flags=self.flags)

        def bw():
            dy = tape.take(y)
            dw9, db = T.conv3x3_backward_weight(dy, x)              # [9, N, K]
            tape.pgrad[conv + ".weight"] = dw9.permute(1, 2, 0).reshape(n, k, 3, 3).contiguous()
            tape.pgrad[conv + ".bias"] = db
            wd = w.permute(1, 2, 3, 0).reshape(k, 9, n).flip(1).contiguous()  # [k][8 - t][n]
            tape.add_grad(x, T.conv3x3_forward(dy, wd, torch.zeros(k, device=self.dev), flags=self.flags))
        tape.ops.append(bw)
        return y

    def bn(self, tape, x, bn_name, act: int):
        """nn.BatchNorm2d in training mode + activation; running statistics updated in the model's buffers."""
        g, b = self._param(bn_name + ".weight"), self._param(bn_name + ".bias")
        c, cp = g.shape[0], x.shape[-1]
        if c == cp:            # the kernel updates the model's running statistics in place
            gp, bp, rm, rv = g, b, self.sd[bn_name + ".running_mean"], self.sd[bn_name + ".running_var"]
        else:                  # 58-channel layers of stage 2 run on 60 channels: padded working copies
            gp = torch.ones(cp, device=self.dev); gp[:c] = g
            bp = torch.zeros(cp, device=self.dev); bp[:c] = b
            rm = torch.zeros(cp, device=self.dev); rm[:c] = self.sd[bn_name + ".running_mean"]
            rv = torch.ones(cp, device=self.dev); rv[:c] = self.sd[bn_name + ".running_var"]
            self.new_stats[bn_name] = (rm[:c], rv[:c])
        self.bn_names.append(bn_name)
        layer = T.BnActTrain(gp, bp, rm, rv, act=act)
        y = layer.forward(x)

        def bw():
            dx = layer.backward(tape.take(y))
            tape.pgrad[bn_name + ".weight"] = layer.grads["weight"][:c]
            tape.pgrad[bn_name + ".bias"] = layer.grads["bias"][:c]
            tape.add_grad(x, dx)
        tape.ops.append(bw)
        return y

    def _move(self, x, a, b, h, hp, op):
        rows = x.numel() // (2 * h)
        T._check(self.lib.ynb_shuffle_unit_move(_ptr(x), _ptr(a), _ptr(b), rows, h, hp, op, self._st()), "ynb_shuffle_unit_move")

    def split(self, tape, x, h, hp):
        """x.chunk(2, dim=C) with both halves zero-padded to hp channels (backbone/shufflenetv2.py:67)."""
        shp = x.shape[:-1]
        x1, x2 = torch.empty((*shp, hp), device=self.dev), torch.empty((*shp, hp), device=self.dev)
        self._move(x, x1, x2, h, hp, 0)

        def bw():
            d1, d2 = tape.take(x1), tape.take(x2)
            dx = torch.empty_like(x)
            self._move(dx, d1, d2, h, hp, 1)
            tape.add_grad(x, dx)
        tape.ops.append(bw)
        return x1, x2

    def shuffle_cat(self, tape, a, b, h, hp):
        """channel_shuffle(torch.cat((a, b), 1), 2): out[2i] = a[i], out[2i+1] = b[i] (backbone/shufflenetv2.py:14-28,76)."""
        y = torch.empty((*a.shape[:-1], 2 * h), device=self.dev)
        self._move(y, a, b, h, hp, 2)

        def bw():
            dy = tape.take(y)
            da, db = torch.empty_like(a), torch.empty_like(b)
            self._move(dy, da, db, h, hp, 3)
            tape.add_grad(a, da)
            tape.add_grad(b, db)
        tape.ops.append(bw)
        return y

    # ---- blocks ----------------------------------------------------------------------------------------------------
    def shuffle_unit(self, tape, x, prefix, cin, cout, stride):
        h = cout // 2
        hp = _pad4(h)

        def branch2(t):
            t = self.bn(tape, self.pw(tape, t, prefix + ".branch2.0", False), prefix + ".branch2.1", ACT_RELU)
            t = self.bn(tape, self.dw(tape, t, prefix + ".branch2.3", stride, False), prefix + ".branch2.4", ACT_NONE)
            return self.bn(tape, self.pw(tape, t, prefix + ".branch2.5", False), prefix + ".branch2.6", ACT_RELU)

        if stride == 1:
            x1, x2 = self.split(tape, x, h, hp)
            return self.shuffle_cat(tape, x1, branch2(x2), h, hp)
        t = self.bn(tape, self.dw(tape, x, prefix + ".branch1.0", stride, False), prefix + ".branch1.1", ACT_NONE)
        b1 = self.bn(tape, self.pw(tape, t, prefix + ".branch1.2", False), prefix + ".branch1.3", ACT_RELU)
        return self.shuffle_cat(tape, b1, branch2(x), h, hp)

    def conv_module(self, tape, x, prefix, k, groups):
        """`Conv` (utils/modules.py:8-18): Conv2d(bias) + BN + LeakyReLU(0.1)."""
        if k == 1:
            y = self.pw(tape, x, prefix + ".convs.0", True)
        elif groups > 1:
            y = self.dw(tape, x, prefix + ".convs.0", 1, True)
        else:
            y = self.conv3(tape, x, prefix + ".convs.0")
        return self.bn(tape, y, prefix + ".convs.1", ACT_LEAKY)

    def merge(self, tape, a, a2, mode):
        """a + F.interpolate(a2) (models/yolo_nano.py:291-296)."""
        b_, h, w_, c = a.shape
        y = torch.empty_like(a)
        T._check(self.lib.ynb_resample_add(_ptr(a), _ptr(a2), _ptr(y), b_, h, w_, c, mode, self._st()), "ynb_resample_add")

        def bw():
            dy = tape.take(y)
            da2 = torch.empty_like(a2)
            T._check(self.lib.ynb_resample_bwd(_ptr(dy), _ptr(da2), b_, h, w_, c, mode, self._st()), "ynb_resample_bwd")
            tape.add_grad(a, dy)
            tape.add_grad(a2, da2)
        tape.ops.append(bw)
        return y

    # ---- the step ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_backward(self, x: torch.Tensor, target: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        m = self.model
        dev = self.dev
        if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3]:
            raise EngineError("x must be a float32 CUDA tensor [B,3,S,S]")
        x = x.contiguous()
        bsz, s = int(x.shape[0]), int(x.shape[2])
        self.sd = {k: v.detach() for k, v in m.state_dict(keep_vars=True).items()}
        self.new_stats: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.bn_names: List[str] = []
        self.flags = [] if self.check_flags else None
        tape = _Tape(dev)
        lib = self.lib

        # stem: conv -> BN -> ReLU -> max-pool (backbone/shufflenetv2.py:109-116, 158-159)
        w0 = self.sd["backbone.conv1.0.weight"]
        w2724 = w0.permute(1, 2, 3, 0).reshape(27, 24).contiguous()
        c1 = torch.empty((bsz, s // 2, s // 2, 24), device=dev)
        T._check(lib.ynb_stem_conv_fwd(_ptr(x), _ptr(w2724), _ptr(c1), bsz, s, self._st()), "ynb_stem_conv_fwd")

        def stem_bw():
            dy = tape.take(c1)
            wsb = lib.ynb_stem_conv_bwd_weight_workspace_bytes(bsz, s)
            ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
            dwt = torch.empty((27, 24), device=dev)
            T._check(lib.ynb_stem_conv_bwd_weight(_ptr(dy), _ptr(x), _ptr(dwt), bsz, s, _ptr(ws), wsb, self._st()),
                     "ynb_stem_conv_bwd_weight")
            tape.pgrad["backbone.conv1.0.weight"] = dwt.reshape(3, 3, 3, 24).permute(3, 0, 1, 2).contiguous()
        tape.ops.append(stem_bw)
        a1 = self.bn(tape, c1, "backbone.conv1.1", ACT_RELU)
        hp_ = (s // 2 - 1) // 2 + 1
        pool = torch.empty((bsz, hp_, hp_, 24), device=dev)
        pidx = torch.empty((bsz, hp_, hp_, 24), device=dev, dtype=torch.uint8)
        T._check(lib.ynb_maxpool3x3s2_fwd_idx(_ptr(a1), _ptr(pool), _ptr(pidx), bsz, s // 2, s // 2, 24, self._st()),
                 "ynb_maxpool3x3s2_fwd_idx")

        def pool_bw():
            dy = tape.take(pool)
            din = torch.empty_like(a1)
            T._check(lib.ynb_maxpool3x3s2_bwd_idx(_ptr(dy), _ptr(pidx), _ptr(din), bsz, s // 2, s // 2, 24, self._st()),
                     "ynb_maxpool3x3s2_bwd_idx")
            tape.add_grad(a1, din)
        tape.ops.append(pool_bw)

        # backbone stages
        t = pool
        feats = []
        cin = STAGE_CHANNELS[0]
        for si, (rep, cout) in enumerate(zip(STAGE_REPEATS, STAGE_CHANNELS[1:])):
            for i in range(rep):
                t = self.shuffle_unit(tape, t, f"backbone.stage{si + 2}.{i}", cin if i == 0 else cout, cout, 2 if i == 0 else 1)
            cin = cout
            feats.append(t)
        c3, c4, c5 = feats
        # neck (models/yolo_nano.py:286-296)
        p3 = self.conv_module(tape, c3, "conv1x1_0", 1, 1)
        p4 = self.conv_module(tape, c4, "conv1x1_1", 1, 1)
        p5 = self.conv_module(tape, c5, "conv1x1_2", 1, 1)
        p4 = self.conv_module(tape, self.merge(tape, p4, p5, 1), "smooth_0", 3, 1)
        p3 = self.conv_module(tape, self.merge(tape, p3, p4, 1), "smooth_1", 3, 1)
        p4 = self.conv_module(tape, self.merge(tape, p4, p3, 2), "smooth_2", 3, 1)
        p5 = self.conv_module(tape, self.merge(tape, p5, p4, 2), "smooth_3", 3, 1)
        # heads (models/yolo_nano.py:50-70, 299-301); the raw maps keep the padded width the loss kernel reads (ld)
        raws = []
        for li, f in enumerate((p3, p4, p5)):
            hd = f"head_det_{li + 1}"
            u = self.conv_module(tape, f, hd + ".0", 3, 96)
            u = self.conv_module(tape, u, hd + ".1", 1, 1)
            u = self.conv_module(tape, u, hd + ".2", 3, 96)
            u = self.conv_module(tape, u, hd + ".3", 1, 1)
            raws.append(self.pw(tape, u, hd + ".4", True))
        raw2 = [r.reshape(bsz, -1, r.shape[-1]) for r in raws]
        losses, grads = T.train_loss(raw2, target.contiguous(), s, m.num_classes, m.anchor_size.view(-1, 2).tolist())
        nch = m.num_anchors * (1 + m.num_classes + 4)
        for r, g in zip(raws, grads):
            g[..., nch:] = 0.0                      # pad columns of the raw maps carry no gradient
            tape.add_grad(r, g.reshape(r.shape))
        tape.backward()
        if self.flags:
            bad = torch.cat(self.flags).view(torch.int32)
            if bool(bad.any()):
                raise EngineError("a tensor-core GEMM of the training step reported a pipeline timeout: %s"
                                  % bad.nonzero().flatten().tolist())

        # running statistics / counters, as nn.BatchNorm2d.forward leaves them
        full = m.state_dict(keep_vars=True)
        for name, (rm, rv) in self.new_stats.items():
            full[name + ".running_mean"].copy_(rm)
            full[name + ".running_var"].copy_(rv)
        torch._foreach_add_([full[name + ".num_batches_tracked"] for name in self.bn_names], 1)
        if hasattr(m, "mark_weights_dirty"):
            m.mark_weights_dirty()
        self.sd = None
        return losses, tape.pgrad

    def flat_gradient(self, grads: Dict[str, torch.Tensor], align: int = 1) -> torch.Tensor:
        """The gradients concatenated in `model.parameters()` order (what train.py:167's optimizer walks); with
        align > 1 every tensor starts at a multiple of `align` elements (zeros in between: `Trainer`'s layout)."""
        parts = []
        for name, p in self.model.named_parameters():
            g = grads[name]
            if tuple(g.shape) != tuple(p.shape):
                raise EngineError(f"gradient of {name} has shape {tuple(g.shape)}, parameter {tuple(p.shape)}")
            parts.append(g.reshape(-1))
            pad = -g.numel() % align
            if pad:
                parts.append(T._zeros(pad, self.dev))
        return torch.cat(parts)


class Trainer:
    """One iteration of the reference's loop body (train.py:219-235) per `step(images, targets)`:

        losses = model(images, target=targets); total = sum(losses)
        if isnan(total): continue                      # skip_nan (train.py:225-226; costs a host read, as there)
        total.backward(); optimizer.step(); optimizer.zero_grad()
        if ema: ema.update(model)

    The parameters are moved into ONE flat float32 vector (each `p.data` becomes a view of it, state_dict layout
    unchanged), so that the data-parallel step is one NCCL all-reduce of 5.3 MB + one fused SGD launch
    (`FlatSGD`; `DistributedDataParallel`'s gradient averaging = the 1 / world folded into that kernel), and the
    EMA one more launch.  `set_lr` = train.py:337-339."""

    def __init__(self, model, lr: float, momentum: float = 0.9, weight_decay: float = 5e-4, ema=None, group=None,
                 skip_nan: bool = True, cuda_graph: bool = False):
        self.model, self.ema, self.group, self.skip_nan = model, ema, group, skip_nan
        # cuda_graph: the ~1600 launches of forward + backward (+ the gradient flattening) are captured ONCE per input
        # shape into a CUDA graph and replayed; the first step of a shape runs eagerly (lazy initialisations),
        # the second is captured.  SGD / all-reduce / EMA stay outside (their scalars change every iteration).
        self.cuda_graph = cuda_graph
        self._gkey, self._graph = None, None
        self.fb = TrainStep(model)
        ps = list(model.parameters())
        for p in ps:
            if p.dtype != torch.float32 or not p.is_cuda:
                raise EngineError("Trainer: parameters must be float32 CUDA tensors (no CPU fallback)")
        # every tensor starts on a 16-byte boundary (the kernels read parameters with 16-byte loads); the <= 3 pad
        # elements after a tensor are zero, get zero gradient and stay zero under SGD with weight decay
        self.align = 4
        flat = torch.zeros(sum(p.numel() + (-p.numel() % self.align) for p in ps), device=ps[0].device)
        off = 0
        for p in ps:
            n = p.numel()
            flat[off:off + n] = p.detach().reshape(-1)
            p.data = flat[off:off + n].view(p.shape)
            off += n + (-n % self.align)
        self.flat = flat
        self.opt = T.FlatSGD(flat, lr, momentum, weight_decay)
        self.skipped = 0
        if hasattr(model, "mark_weights_dirty"):
            model.mark_weights_dirty()

    def set_lr(self, lr: float):
        self.opt.set_lr(lr)

    def _forward_backward(self, images, targets):
        if not self.cuda_graph:
            losses, grads = self.fb.forward_backward(images, targets)
            return losses, self.fb.flat_gradient(grads, self.align)
        key = (tuple(images.shape), tuple(targets.shape))
        if key != self._gkey:                         # new shape (multi-scale training): eager now, capture next time
            self._gkey, self._graph = key, None
            losses, grads = self.fb.forward_backward(images, targets)
            return losses, self.fb.flat_gradient(grads, self.align)
        if self._graph is None:
            sx, st = images.clone(), targets.clone()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                losses, grads = self.fb.forward_backward(sx, st)
                flat = self.fb.flat_gradient(grads, self.align)
            self._graph = (g, sx, st, losses, flat)
        else:
            g, sx, st, losses, flat = self._graph
            sx.copy_(images)
            st.copy_(targets)
        g.replay()
        if hasattr(self.model, "mark_weights_dirty"):
            self.model.mark_weights_dirty()
        return losses, flat

    @torch.no_grad()
    def step(self, images: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        losses, flat_grad = self._forward_backward(images, targets)
        if self.skip_nan and bool(torch.isnan(losses.sum()).item()):
            self.skipped += 1
            return losses
        self.opt.step(flat_grad, self.group)
        if hasattr(self.model, "mark_weights_dirty"):
            self.model.mark_weights_dirty()
        if self.ema is not None:
            self.ema.update(self.model)
        return losses
