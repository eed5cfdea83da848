"""TEST INFRASTRUCTURE (checker only) — CPU restatement of the reference's training branch
at the head boundary.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this.

Pinned: tests/golden/g7_train128.npz is recorded from the REAL reference
(`oracle/gen_golden.py train`): `tools.multi_gt_creator`, `YOLONano.forward(x, target)` with
`trainable=True`, `total_loss.backward()`; tests/test_oracle_golden.py checks this file against it.

Everything here is float32 torch on the CPU, the same operators in the same order as
  models/yolo_nano.py:333-358   (training branch of forward)
  tools.py:12-34                (MSEWithLogitsLoss)
  tools.py:219-233              (iou_score)
  tools.py:236-276              (loss)
  tools.py:97-216               (multi_gt_creator; float64 NumPy, cast to float32 at the end)
  train.py:167-171,222-231      (SGD momentum 0.9, weight decay 5e-4; total = sum of the four)
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .yolo_nano_oracle import STRIDES, grid_tensors, split_predictions

IGNORE_THRESH = 0.5  # data/config.py:3


# --------------------------------------------------------------------------------------
# tools.py:97-216 — target assignment (host, float64)
# --------------------------------------------------------------------------------------
def _anchor_iou(anchor_wh: np.ndarray, bw: float, bh: float) -> np.ndarray:
    """tools.py:37-77 for boxes centred on the origin: [0,0,aw,ah] against [0,0,bw,bh]."""
    aw, ah = anchor_wh[:, 0], anchor_wh[:, 1]
    iw = np.minimum(bw / 2, aw / 2) - np.maximum(-bw / 2, -aw / 2)
    ih = np.minimum(bh / 2, ah / 2) - np.maximum(-bh / 2, -ah / 2)
    si = ih * iw
    return si / (bw * bh + aw * ah - si + 1e-20)


def multi_gt_creator(input_size: int, label_lists: Sequence[Sequence[Sequence[float]]], anchor_size,
                     strides=STRIDES) -> torch.Tensor:
    """label_lists[b] = [[xmin, ymin, xmax, ymax, cls], ...] normalised to [0,1].
    Returns target [B, N, 11] float32 = obj(1/0/-1), cls, tx, ty, tw, th, weight, x1, y1, x2, y2.
    Later labels overwrite earlier ones that fall on the same (cell, anchor)."""
    B = len(label_lists)
    anchors = np.asarray(anchor_size, dtype=np.float64).reshape(-1, 2)
    A = len(anchors) // len(strides)
    gt = [np.zeros([B, input_size // s, input_size // s, A, 11]) for s in strides]
    w = h = input_size
    for b in range(B):
        for lab in label_lists[b]:
            cls = int(lab[-1])
            xmin, ymin, xmax, ymax = lab[:-1]
            cx, cy = (xmax + xmin) / 2 * w, (ymax + ymin) / 2 * h
            bw, bh = (xmax - xmin) * w, (ymax - ymin) * h
            if bw < 1.0 or bh < 1.0:
                continue
            iou = _anchor_iou(anchors, bw, bh)
            mask = iou > IGNORE_THRESH
            best = int(np.argmax(iou))
            for index in range(len(anchors)):
                if not (index == best or mask[index]):
                    continue
                lvl = index // A
                a = index - lvl * A
                s = strides[lvl]
                gx, gy = int(cx / s), int(cy / s)
                if index == best:
                    # the `mask.sum() == 0` branch (:131-157) and the best-index branch (:169-195)
                    # are the same assignment, including the bounds check
                    if gy < gt[lvl].shape[1] and gx < gt[lvl].shape[2]:
                        weight = 2.0 - (bw / w) * (bh / h)
                        gt[lvl][b, gy, gx, a] = [1.0, cls, cx / s - gx, cy / s - gy,
                                                 np.log(bw / anchors[index, 0]), np.log(bh / anchors[index, 1]),
                                                 weight, xmin, ymin, xmax, ymax]
                else:
                    # ignored anchors (:197-209): no bounds check in the reference (IndexError
                    # there when the centre sits exactly on the right / bottom border)
                    gt[lvl][b, gy, gx, a, 0] = -1.0
                    gt[lvl][b, gy, gx, a, 6] = -1.0
    out = np.concatenate([g.reshape(B, -1, 11) for g in gt], 1)
    return torch.from_numpy(out).float()


# --------------------------------------------------------------------------------------
# models/yolo_nano.py:333-358 + tools.py:12-34,219-276 — the four losses
# --------------------------------------------------------------------------------------
def decode_boxes_train(txtytwth: torch.Tensor, input_size: int, anchor_size, num_anchors: int = 3):
    """decode_boxes(...) / input_size as the training branch uses it (:336-337): NO clamp."""
    b = txtytwth.shape[0]
    grid, stride, anchors = grid_tensors(input_size, anchor_size, num_anchors)
    t = txtytwth.view(b, -1, num_anchors, 4)
    cxy = (torch.sigmoid(t[..., :2]) + grid) * stride
    wh = torch.exp(t[..., 2:]) * anchors
    xywh = torch.cat([cxy, wh], -1).view(b, -1, 4)
    box = torch.zeros_like(xywh)
    box[:, :, 0] = xywh[:, :, 0] - xywh[:, :, 2] / 2
    box[:, :, 1] = xywh[:, :, 1] - xywh[:, :, 3] / 2
    box[:, :, 2] = xywh[:, :, 0] + xywh[:, :, 2] / 2
    box[:, :, 3] = xywh[:, :, 1] + xywh[:, :, 3] / 2
    return box / input_size


def iou_score(a: torch.Tensor, b: torch.Tensor, batch_size: int) -> torch.Tensor:
    """tools.py:219-233."""
    tl = torch.max(a[:, :2], b[:, :2])
    br = torch.min(a[:, 2:], b[:, 2:])
    area_a = torch.prod(a[:, 2:] - a[:, :2], 1)
    area_b = torch.prod(b[:, 2:] - b[:, :2], 1)
    en = (tl < br).type(tl.type()).prod(dim=1)
    area_i = torch.prod(br - tl, 1) * en
    return (area_i / (area_a + area_b - area_i)).view(batch_size, -1, 1)


def losses_from_predictions(conf, cls, txtytwth, target, input_size: int, anchor_size, num_anchors: int = 3):
    """conf [B,N,1], cls [B,N,C], txtytwth [B,N,4] (models/yolo_nano.py:327-330), target [B,N,11]
    -> (conf_loss, cls_loss, bbox_loss, iou_loss), differentiable."""
    B = conf.shape[0]
    box = decode_boxes_train(txtytwth, input_size, anchor_size, num_anchors).view(-1, 4)
    iou_pred = iou_score(box, target[:, :, 7:].reshape(-1, 4), B)
    with torch.no_grad():
        gt_conf = iou_pred.clone()
    label = torch.cat([gt_conf, target[:, :, :7]], dim=2)
    # tools.loss
    pred_conf = conf[:, :, 0]
    pred_cls = cls.permute(0, 2, 1)
    pred_txty, pred_twth = txtytwth[:, :, :2], txtytwth[:, :, 2:]
    pred_iou = iou_pred[:, :, 0]
    gt_c = label[:, :, 0].float()
    gt_obj = label[:, :, 1].float()
    gt_cls = label[:, :, 2].long()
    gt_txty, gt_twth = label[:, :, 3:5].float(), label[:, :, 5:7].float()
    wgt = label[:, :, 7].float()
    gt_mask = (gt_obj > 0.).float()
    # MSEWithLogitsLoss (tools.py:12-34)
    p = torch.sigmoid(pred_conf)
    pos, neg = (gt_obj == 1.0).float(), (gt_obj == 0.0).float()
    conf_loss = torch.sum(5.0 * (pos * (p - gt_c) ** 2) + 1.0 * (neg * p ** 2)) / B
    cls_loss = torch.sum(F.cross_entropy(pred_cls, gt_cls, reduction="none") * gt_mask) / B
    txty = torch.sum(torch.sum(F.binary_cross_entropy_with_logits(pred_txty, gt_txty, reduction="none"), -1)
                     * wgt * gt_mask) / B
    twth = torch.sum(torch.sum(F.mse_loss(pred_twth, gt_twth, reduction="none"), -1) * wgt * gt_mask) / B
    iou_loss = torch.sum(F.smooth_l1_loss(pred_iou, gt_mask.clone(), reduction="none")) / B
    return conf_loss, cls_loss, txty + twth, iou_loss


def losses_and_grads(preds: Sequence[torch.Tensor], target: torch.Tensor, input_size: int, num_classes: int,
                     anchor_size, num_anchors: int = 3) -> Tuple[List[float], List[torch.Tensor]]:
    """preds: the three raw head maps NCHW (as the heads produce them).  Returns the four losses
    and d(total)/d(pred) per level, total = sum of the four (train.py:222)."""
    leaves = [p.detach().clone().requires_grad_(True) for p in preds]
    confs, clss, boxes = [], [], []
    a, c = num_anchors, num_classes
    for pred in leaves:  # models/yolo_nano.py:303-330
        b, ch, h, w = pred.shape
        q = pred.permute(0, 2, 3, 1).contiguous().view(b, h * w, ch)
        confs.append(q[:, :, :a].contiguous().view(b, h * w * a, 1))
        clss.append(q[:, :, a:(1 + c) * a].contiguous().view(b, h * w * a, c))
        boxes.append(q[:, :, (1 + c) * a:].contiguous().view(b, h * w * a, 4))
    ls = losses_from_predictions(torch.cat(confs, 1), torch.cat(clss, 1), torch.cat(boxes, 1), target,
                                 input_size, anchor_size, num_anchors)
    sum(ls).backward()
    return [float(v.detach()) for v in ls], [p.grad for p in leaves]


# --------------------------------------------------------------------------------------
# train.py:167-171,230 — torch.optim.SGD(lr, momentum=0.9, weight_decay=5e-4), one step
# --------------------------------------------------------------------------------------
def sgd_step(p: torch.Tensor, g: torch.Tensor, buf, lr: float, momentum: float = 0.9, weight_decay: float = 5e-4):
    """Returns (p_new, buf_new); buf None = first step (the buffer becomes the gradient)."""
    g = g.add(p, alpha=weight_decay)
    buf = g.clone() if buf is None else buf.mul(momentum).add(g)
    return p.add(buf, alpha=-lr), buf


_ = split_predictions  # re-exported for the tests


# --------------------------------------------------------------------------------------
# ModelEMA (utils/misc.py:67-86)
# --------------------------------------------------------------------------------------
def ema_decay(updates: int, decay: float = 0.9999) -> float:
    """utils/misc.py:72: decay * (1 - exp(-x / 2000)), Python double."""
    import math
    return decay * (1 - math.exp(-updates / 2000.))


@torch.no_grad()
def ema_update(ema_sd: dict, model_sd: dict, d: float) -> None:
    """ModelEMA.update body (utils/misc.py:82-86), in place on `ema_sd`: the same two in-place tensor ops per
    floating-point entry; integer entries (num_batches_tracked) are left alone."""
    for k, v in ema_sd.items():
        if v.dtype.is_floating_point:
            v *= d
            v += (1. - d) * model_sd[k].detach()


def ema_model(seed: int):
    """Small conv + BatchNorm stack (parameters, running statistics, an integer counter) for the ModelEMA fixture g8 —
    the class is model-agnostic (utils/misc.py:67-86); sizes chosen to cover vector tails and 4-byte-only alignment."""
    torch.manual_seed(seed)
    m = torch.nn.Sequential(torch.nn.Conv2d(3, 7, 3), torch.nn.BatchNorm2d(7), torch.nn.Conv2d(7, 129, 1),
                            torch.nn.BatchNorm2d(129), torch.nn.Conv2d(129, 33, 3, bias=False))
    with torch.no_grad():
        for b in m.buffers():
            if b.dtype.is_floating_point:
                b.uniform_(0.5, 1.5)
    return m


@torch.no_grad()
def ema_perturb(m, g: torch.Generator) -> None:
    """What an optimizer step + a BatchNorm forward would do to the model between two EMA updates (deterministic)."""
    for v in m.state_dict().values():
        if v.dtype.is_floating_point:
            v += (0.05 * torch.randn(v.shape, generator=g)).to(v.device)
        else:
            v += 1


# --------------------------------------------------------------------------------------
# the whole training step (train.py:219-231): model.train()(x, target) -> losses, backward, SGD
# --------------------------------------------------------------------------------------
def train_forward_backward(sd: dict, x: torch.Tensor, target: torch.Tensor, input_size: int, num_classes: int,
                           anchor_size, num_anchors: int = 3):
    """YOLONano in train() mode with trainable=True (models/yolo_nano.py:333-358): BatchNorm on batch statistics
    (running statistics updated, momentum 0.1), the four losses, `total_loss.backward()` (train.py:222-229).
    Returns (losses [4], grads {parameter name: tensor}, new_state {running_* / num_batches_tracked after the step}).
    `sd` is not modified."""
    from . import yolo_nano_oracle as O
    work = {}
    for k, v in sd.items():
        if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"):
            work[k] = v.detach().clone()
        else:
            work[k] = v.detach().clone().requires_grad_(True)
    O._BN_TRAIN = True
    try:
        preds = O.network_graph(work, x)
    finally:
        O._BN_TRAIN = False
    confs, clss, boxes = [], [], []
    a, c = num_anchors, num_classes
    for pred in preds:
        b, ch, h, w = pred.shape
        q = pred.permute(0, 2, 3, 1).contiguous().view(b, h * w, ch)
        confs.append(q[:, :, :a].contiguous().view(b, h * w * a, 1))
        clss.append(q[:, :, a:(1 + c) * a].contiguous().view(b, h * w * a, c))
        boxes.append(q[:, :, (1 + c) * a:].contiguous().view(b, h * w * a, 4))
    ls = losses_from_predictions(torch.cat(confs, 1), torch.cat(clss, 1), torch.cat(boxes, 1), target, input_size,
                                 anchor_size, num_anchors)
    sum(ls).backward()
    grads = {k: v.grad for k, v in work.items() if v.requires_grad}
    state = {k: v for k, v in work.items() if not v.requires_grad}
    for k in state:
        if k.endswith("num_batches_tracked"):
            state[k] = state[k] + 1          # F.batch_norm does not count; nn.BatchNorm2d.forward does
    return [float(v.detach()) for v in ls], grads, state

