"""Host side of the training branch at the head boundary (SURVEY §8 row a14, §8f row 4;
BASELINE config 5), over the C ABI of include/yolonano_b200.h.

Built: target assignment (`tools.multi_gt_creator`), the four losses + their gradient w.r.t. the raw
head maps (`models/yolo_nano.py:333-358`, `tools.py:12-34,219-276`), the SGD update
(`train.py:167-171`), the gradient all-reduce of the data-parallel config, and the
backward kernels of the depthwise / pointwise / dense 3x3 convolutions and of training-mode BatchNorm (batch
statistics, running-statistics update) + activation, composable per module (`SequentialTrain`).  NOT built: the
chaining of these modules through the whole network — `YOLONano.forward(x, target)` in `.train()` mode raises
NotImplementedError and the losses of the `.eval()` training branch carry no `grad_fn` (this package is a drop-in
for the detection forward path, not for `train.py`'s `total_loss.backward()`).

PyTorch supplies device memory, the stream and `torch.distributed`; every computation is in
libyolonano_b200.so.  No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .engine import EngineError, _ptr, _stream_ptr

STRIDES = (8, 16, 32)


def _check(rc: int, what: str):
    if rc != _lib.YNB_OK:
        msg = _lib.load().ynb_last_error(None)
        raise EngineError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def _dev(*ts: torch.Tensor) -> torch.device:
    for t in ts:
        if t is not None and (not t.is_cuda or t.dtype not in (torch.float32, torch.int32) or not t.is_contiguous()):
            raise EngineError("training kernels take contiguous float32 / int32 CUDA tensors (no CPU fallback)")
    return ts[0].device


def _anchors(anchor_size) -> "C.Array":
    flat = [float(v) for wh in anchor_size for v in wh]
    if len(flat) % 6:
        raise EngineError("anchor_size must hold 3 levels x A [w,h] pairs")
    return (C.c_float * len(flat))(*flat), len(flat) // 6


def build_targets(labels: torch.Tensor, counts: Optional[torch.Tensor], input_size: int, anchor_size) -> torch.Tensor:
    """`tools.multi_gt_creator(input_size, strides, label_lists, anchor_size)` (tools.py:97-216).
    labels [B, L, 5] float32 CUDA (xmin, ymin, xmax, ymax normalised, class), counts [B] int32 or None
    -> target [B, N, 11] float32 on the same device."""
    dev = _dev(labels, counts)
    anc, a = _anchors(anchor_size)
    b, l, five = labels.shape
    assert five == 5
    n = sum(a * (input_size // s) ** 2 for s in STRIDES)
    target = torch.empty((b, n, 11), device=dev, dtype=torch.float32)
    _check(_lib.load().ynb_build_targets(_ptr(labels), _ptr(counts), b, l, input_size, anc, a, _ptr(target),
                                         _stream_ptr(dev)), "ynb_build_targets")
    return target


def train_loss(raw: Sequence[torch.Tensor], target: torch.Tensor, input_size: int, num_classes: int, anchor_size,
               ) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
    """raw = three NHWC maps [B, H*W, ld] (reference channel map) -> (losses [4] = conf, cls, bbox, iou
    on the device, gradients of their sum w.r.t. the three maps)."""
    dev = _dev(*raw, target)
    lib = _lib.load()
    anc, a = _anchors(anchor_size)
    b, ld = raw[0].shape[0], raw[0].shape[-1]
    for r, s in zip(raw, STRIDES):
        if r.shape[0] != b or r.shape[-1] != ld or r.numel() != b * (input_size // s) ** 2 * ld:
            raise EngineError("raw head maps must be [B, (S/stride)^2, ld] with one ld")
    if tuple(target.shape) != (b, sum(a * (input_size // s) ** 2 for s in STRIDES), 11):
        raise EngineError("target must be [B, N, 11]")
    grads = tuple(torch.empty_like(r) for r in raw)
    losses = torch.empty(4, device=dev, dtype=torch.float32)
    wsb = lib.ynb_train_loss_workspace_bytes(b, input_size)
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    _check(lib.ynb_train_loss(_ptr(raw[0]), _ptr(raw[1]), _ptr(raw[2]), ld, _ptr(target), b, input_size, anc, a,
                              num_classes, _ptr(losses), _ptr(grads[0]), _ptr(grads[1]), _ptr(grads[2]),
                              _ptr(ws), wsb, _stream_ptr(dev)), "ynb_train_loss")
    return losses, grads


class FlatSGD:
    """`torch.optim.SGD(params, lr, momentum=0.9, weight_decay=5e-4)` (train.py:167-171) over ONE flat
    float32 device vector (the 1.33 M parameters of the model are 5.3 MB: one launch, one all-reduce).
    `step(grad)` with `world_size > 1` first all-reduces (sum) the gradient over the process group —
    NCCL over NVLink on GPUs — and folds the 1 / world_size into the update kernel."""

    def __init__(self, params: torch.Tensor, lr: float, momentum: float = 0.9, weight_decay: float = 5e-4):
        _dev(params)
        self.params = params
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        self.buf = torch.zeros_like(params)
        self.steps = 0

    def set_lr(self, lr: float):          # train.py:337-339
        self.lr = float(lr)

    def step(self, grad: torch.Tensor, group=None):
        dev = _dev(grad)
        scale = 1.0
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            allreduce_gradients(grad, group)
            scale = 1.0 / dist.get_world_size(group)
        _check(_lib.load().ynb_sgd_step(_ptr(self.params), _ptr(grad), _ptr(self.buf), self.params.numel(), self.lr,
                                        self.momentum, self.weight_decay, int(self.steps == 0), scale,
                                        _stream_ptr(dev)), "ynb_sgd_step")
        self.steps += 1


def allreduce_gradients(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """ONE all-reduce (sum) of the flat gradient per step; the division by the
    world size happens in the SGD kernel.  In place; returns its argument."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


# ---- backward of the conv kernels (individually callable, NHWC with (ld, off) channel windows) ----------
def dwconv3x3_backward(dout: torch.Tensor, x: torch.Tensor, w9c: torch.Tensor, stride: int):
    """dout [B,Ho,Wo,C], x [B,H,W,C], w9c [9,C] -> (dx [B,H,W,C], dw [9,C], db [C])."""
    dev = _dev(dout, x, w9c)
    lib = _lib.load()
    b, h, w, c = x.shape
    dx = torch.empty_like(x)
    _check(lib.ynb_dwconv3x3_bwd_data(_ptr(dout), c, 0, _ptr(dx), c, 0, _ptr(w9c), b, h, w, c, stride,
                                      _stream_ptr(dev)), "ynb_dwconv3x3_bwd_data")
    wsb = lib.ynb_dwconv3x3_bwd_weight_workspace_bytes(b, h, w, c, stride)
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    dwdb = torch.empty((10, c), device=dev, dtype=torch.float32)
    _check(lib.ynb_dwconv3x3_bwd_weight(_ptr(dout), c, 0, _ptr(x), c, 0, _ptr(dwdb), b, h, w, c, stride,
                                        _ptr(ws), wsb, _stream_ptr(dev)), "ynb_dwconv3x3_bwd_weight")
    return dx, dwdb[:9], dwdb[9]


def pwconv_backward_weight(dout: torch.Tensor, x: torch.Tensor):
    """dout [M, N], x [M, K] -> (dw [N, K], db [N])."""
    dev = _dev(dout, x)
    lib = _lib.load()
    m, n = dout.shape
    k = x.shape[1]
    wsb = lib.ynb_pwconv_bwd_weight_workspace_bytes(m, k, n)
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    dw = torch.empty((n, k), device=dev, dtype=torch.float32)
    db = torch.empty((n,), device=dev, dtype=torch.float32)
    _check(lib.ynb_pwconv_bwd_weight(_ptr(dout), n, 0, _ptr(x), k, 0, _ptr(dw), _ptr(db), m, k, n, _ptr(ws), wsb,
                                     _stream_ptr(dev)), "ynb_pwconv_bwd_weight")
    return dw, db


def conv3x3_backward_weight(dout: torch.Tensor, x: torch.Tensor):
    """Dense 3x3 conv, pad 1, stride 1 (the `smooth` convs): dout [B,H,W,N], x [B,H,W,K] ->
    (dw [9, N, K] tap-major, db [N])."""
    dev = _dev(dout, x)
    lib = _lib.load()
    b, h, w, n = dout.shape
    k = x.shape[-1]
    wsb = lib.ynb_pwconv_bwd_weight_workspace_bytes(b * h * w, k, n)
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    dw = torch.empty((9, n, k), device=dev, dtype=torch.float32)
    db = torch.empty((n,), device=dev, dtype=torch.float32)
    _check(lib.ynb_conv3x3_bwd_weight(_ptr(dout), n, 0, _ptr(x), k, 0, _ptr(dw), _ptr(db), b, h, w, k, n, _ptr(ws), wsb,
                                      _stream_ptr(dev)), "ynb_conv3x3_bwd_weight")
    return dw, db


_ZEROS = {}


def _zeros(n: int, dev) -> torch.Tensor:
    """A shared read-only zero vector (bias of bias-free convs): allocated once per (device, size bucket)."""
    key = (dev, (n + 511) // 512)
    z = _ZEROS.get(key)
    if z is None:
        z = _ZEROS[key] = torch.zeros(key[1] * 512, device=dev, dtype=torch.float32)
    return z[:n]


def _tc_workspace(cout: int, ktot: int, dev):
    n = int(_lib.load().ynb_tc_async_workspace_bytes(cout, ktot))
    return torch.empty(n, device=dev, dtype=torch.uint8), n


def pwconv_forward(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], act: int = 0, transposed: bool = False,
                   cout: Optional[int] = None, flags: Optional[list] = None) -> torch.Tensor:
    """y [M, N] = act(x [M, K] . W^T + b) on the tcgen05 GEMM (3xTF32), asynchronous (`ynb_pwconv_tc_async`: the weights
    are split into hi / lo planes by a kernel on the same stream).
    transposed = False: W = w [N, Kw] (Kw <= K; missing input channels have zero weight), N = w.shape[0] or `cout`
                        (rows beyond w.shape[0] are zero: channel padding).
    transposed = True:  W^T = w [Kw, Nw]: y = x . w — the input gradient of the layer whose weight is w."""
    dev = _dev(x, w, b)
    lib = _lib.load()
    m, k = x.shape
    if w.stride(-1) != 1:
        w = w.contiguous()
    ld = w.stride(0)
    if transposed:
        kw, n_real = w.shape
    else:
        n_real, kw = w.shape
    n = n_real if cout is None else int(cout)
    y = torch.empty((m, n), device=dev, dtype=torch.float32)
    if b is None:
        b = _zeros(n, dev)
    elif b.numel() < n:
        b = torch.cat((b, b.new_zeros(n - b.numel())))
    for n0 in range(0, n, 256):                       # the GEMM holds at most 256 output columns in TMEM
        nc = min(256, n - n0)
        rows = max(0, min(nc, n_real - n0))           # real weight rows in this slice (the rest: zero rows)
        if rows == 0:
            y[:, n0:n0 + nc] = b[n0:n0 + nc]
            continue
        ws, wsb = _tc_workspace(nc, k, dev)
        wp = w[:, n0:] if transposed else w[n0:]
        # rows < nc only when padding the channel count: the pack kernel zero-fills rows >= `rows`
        _check(lib.ynb_pwconv_tc_async(_ptr(x), k, 0, _ptr(y), n, n0, 1, _ptr(wp), ld, rows, kw, int(transposed), _ptr(b[n0:]),
                                       m, k, nc, act, _lib.GEMM_TC_3XTF32, _ptr(ws), wsb, _stream_ptr(dev)),
               "ynb_pwconv_tc_async")
        if flags is not None:
            flags.append(ws[-256:-252])
    return y


def pwconv_backward_data(dout: torch.Tensor, w_nk: torch.Tensor, tensor_cores: bool = True, flags: Optional[list] = None
                         ) -> torch.Tensor:
    """dx [M, K] = dout [M, N] . w [N, K]: the forward pointwise GEMM with the weights read transposed and a
    zero bias (tcgen05 3xTF32 by default).  N and K must be multiples of 4."""
    dev = _dev(dout, w_nk)
    if tensor_cores:
        return pwconv_forward(dout, w_nk, None, transposed=True, flags=flags)
    lib = _lib.load()
    m, n = dout.shape
    k = w_nk.shape[1]
    wt = w_nk.t().contiguous()                      # [K][N]: "cout" = K, "cin" = N
    zero = torch.zeros(k, device=dev, dtype=torch.float32)
    dx = torch.empty((m, k), device=dev, dtype=torch.float32)
    _check(lib.ynb_pwconv(_ptr(dout), n, 0, _ptr(dx), k, 0, 1, _ptr(wt), _ptr(zero), m, n, k, 0,
                          _stream_ptr(dev)), "ynb_pwconv")
    return dx


def conv3x3_forward(x: torch.Tensor, w_ntk: torch.Tensor, b: torch.Tensor, act: int = 0, flags: Optional[list] = None
                    ) -> torch.Tensor:
    """Dense 3x3, pad 1, stride 1 on NHWC x [B,H,W,K] (K % 32 == 0) with w [N][9][K] tap-major (asynchronous)."""
    dev = _dev(x, w_ntk, b)
    bsz, h, w_, k = x.shape
    n = w_ntk.shape[0]
    y = torch.empty((bsz, h, w_, n), device=dev, dtype=torch.float32)
    ws, wsb = _tc_workspace(n, 9 * k, dev)
    _check(_lib.load().ynb_conv3x3_tc_async(_ptr(x), k, _ptr(y), n, _ptr(w_ntk), _ptr(b), bsz, h, w_, k, n, act,
                                            _lib.GEMM_TC_3XTF32, _ptr(ws), wsb, _stream_ptr(dev)), "ynb_conv3x3_tc_async")
    if flags is not None:
        flags.append(ws[-256:-252])
    return y


def act_backward(dout: torch.Tensor, out: torch.Tensor, act: int) -> torch.Tensor:
    """dout, out [M, C]; act 1 = ReLU, 2 = LeakyReLU(0.1)."""
    dev = _dev(dout, out)
    m, c = dout.shape
    dpre = torch.empty_like(dout)
    _check(_lib.load().ynb_act_bwd(_ptr(dout), c, 0, _ptr(out), c, 0, _ptr(dpre), c, 0, m, c, act, _stream_ptr(dev)),
           "ynb_act_bwd")
    return dpre



class BatchNormTrain:
    """nn.BatchNorm2d in training mode + the following activation on an NHWC matrix [M, C] (C % 4 == 0):
    `forward` normalises with the batch statistics and updates the running ones in place, `backward`
    returns (dx, dgamma, dbeta).  utils/modules.py:13-14, backbone/shufflenetv2.py:47-62."""

    def __init__(self, gamma, beta, running_mean, running_var, eps: float = 1e-5, momentum: float = 0.1, act: int = 0):
        _dev(gamma, beta, running_mean, running_var)
        self.gamma, self.beta, self.running_mean, self.running_var = gamma, beta, running_mean, running_var
        self.eps, self.momentum, self.act = float(eps), float(momentum), int(act)
        self.saved = None

    def _ws(self, m, c, dev):
        wsb = _lib.load().ynb_bn_workspace_bytes(m, c)
        return torch.empty(wsb, device=dev, dtype=torch.uint8), wsb

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        dev = _dev(x)
        m, c = x.shape
        y = torch.empty_like(x)
        mean, rstd = torch.empty(c, device=dev), torch.empty(c, device=dev)
        ws, wsb = self._ws(m, c, dev)
        _check(_lib.load().ynb_bn_train_fwd(_ptr(x), c, 0, _ptr(y), c, 0, _ptr(self.gamma), _ptr(self.beta),
                                            _ptr(self.running_mean), _ptr(self.running_var), _ptr(mean), _ptr(rstd),
                                            m, c, self.eps, self.momentum, self.act, _ptr(ws), wsb, _stream_ptr(dev)),
               "ynb_bn_train_fwd")
        self.saved = (x, y, mean, rstd)
        return y

    def backward(self, dy: torch.Tensor):
        dev = _dev(dy)
        x, y, mean, rstd = self.saved
        m, c = x.shape
        dx = torch.empty_like(x)
        dgb = torch.empty(2 * c, device=dev)
        ws, wsb = self._ws(m, c, dev)
        _check(_lib.load().ynb_bn_train_bwd(_ptr(dy), c, 0, _ptr(x), c, 0, _ptr(y), c, 0, _ptr(self.gamma), _ptr(mean),
                                            _ptr(rstd), _ptr(dx), c, 0, _ptr(dgb), m, c, self.act, _ptr(ws), wsb,
                                            _stream_ptr(dev)), "ynb_bn_train_bwd")
        return dx, dgb[:c], dgb[c:]


# ---- layer objects: the kernels above composed into forward / backward of the reference's blocks ---------
class PwConvTrain:
    """1x1 conv (`nn.Conv2d(cin, cout, 1)`, optional bias) on [M, cin] -> [M, cout]; forward on the
    tcgen05 GEMM (3xTF32), backward = weight gradient kernel + the forward GEMM with W^T."""

    def __init__(self, weight_nk: torch.Tensor, bias: Optional[torch.Tensor] = None):
        _dev(weight_nk, bias)
        self.w, self.b = weight_nk, bias
        self.grads = {}

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        dev = _dev(x)
        m, k = x.shape
        n = self.w.shape[0]
        self.x = x
        return pwconv_forward(x, self.w, self.b)

    def backward(self, dy: torch.Tensor) -> torch.Tensor:
        dw, db = pwconv_backward_weight(dy, self.x)
        self.grads = {"weight": dw} if self.b is None else {"weight": dw, "bias": db}
        return pwconv_backward_data(dy, self.w)


class DwConvTrain:
    """Depthwise 3x3 (`nn.Conv2d(c, c, 3, stride, 1, groups=c, bias=False)`) on NHWC [B, H, W, C]."""

    def __init__(self, weight_9c: torch.Tensor, stride: int):
        _dev(weight_9c)
        self.w, self.stride = weight_9c, int(stride)
        self.grads = {}

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        dev = _dev(x)
        b, h, w, c = x.shape
        self.x = x
        ho, wo = (h - 1) // self.stride + 1, (w - 1) // self.stride + 1
        y = torch.empty((b, ho, wo, c), device=dev, dtype=torch.float32)
        zero = torch.zeros(c, device=dev)
        _check(_lib.load().ynb_dwconv3x3(_ptr(x), c, 0, _ptr(y), c, 0, 1, _ptr(self.w), _ptr(zero), b, h, w, c, self.stride,
                                         0, _stream_ptr(dev)), "ynb_dwconv3x3")
        return y

    def backward(self, dy: torch.Tensor) -> torch.Tensor:
        dx, dw, _db = dwconv3x3_backward(dy, self.x, self.w, self.stride)
        self.grads = {"weight": dw}
        return dx


class BnActTrain(BatchNormTrain):
    """BatchNormTrain on tensors of any leading shape [..., C] (NHWC maps or [M, C] matrices)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.shape = x.shape
        return super().forward(x.reshape(-1, x.shape[-1])).reshape(self.shape)

    def backward(self, dy: torch.Tensor) -> torch.Tensor:
        dx, dg, db = super().backward(dy.reshape(-1, dy.shape[-1]))
        self.grads = {"weight": dg, "bias": db}
        return dx.reshape(self.shape)


class SequentialTrain:
    """forward through the layers, backward in reverse; `grads()` lists the parameter gradients in
    layer order — what `total_loss.backward()` (train.py:229) leaves in `.grad` for this block."""

    def __init__(self, *layers):
        self.layers = list(layers)

    def forward(self, x):
        for layer in self.layers:
            if isinstance(layer, PwConvTrain) and x.dim() == 4:
                shp = x.shape
                x = layer.forward(x.reshape(-1, shp[-1])).reshape(*shp[:-1], -1)
            else:
                x = layer.forward(x)
        return x

    def backward(self, dy):
        for layer in reversed(self.layers):
            if isinstance(layer, PwConvTrain) and dy.dim() == 4:
                shp = dy.shape
                dy = layer.backward(dy.reshape(-1, shp[-1]).contiguous()).reshape(*shp[:-1], -1)
            else:
                dy = layer.backward(dy.contiguous())
        return dy

    def grads(self):
        return [layer.grads for layer in self.layers]
