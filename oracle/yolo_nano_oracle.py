"""CPU oracle for the YOLO-Nano-1.0x detection forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under `yolo_nano_b200/` may import this module; it
is used by `tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s CPU-baseline /
`--impl reference` legs as the checker and as the timed CPU restatement.

What it is: a functional (torch fp32 on CPU + NumPy) restatement of the reference
algorithm, working directly on a reference-layout `state_dict`, each function citing
the reference lines it follows (paths relative to the reference checkout).  The
arithmetic itself lives in third-party code the reference does not pin
(PyTorch ATen/oneDNN for the network, NumPy for NMS; installed here: torch 2.11.0,
numpy 2.3.5); the restatement calls the same library operators in the same order.

Parity pin: `tests/golden/*.npz` are produced by `oracle/gen_golden.py`, which imports
the REAL reference from /root/reference and records its outputs;
`tests/test_oracle_golden.py` checks this restatement against them (bit-exact on the
machine that generated them, tight tolerance elsewhere).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

STRIDES = (8, 16, 32)
STAGES = ((2, 4), (3, 8), (4, 4))
BN_EPS = 1e-5

StateDict = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def _conv(sd: StateDict, prefix: str, x, stride=1, padding=0, groups=1):
    """nn.Conv2d in eval (utils/modules.py:11, backbone/shufflenetv2.py:44-67)."""
    return F.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride, padding, 1, groups)


# oracle/weights.py flips this to measure activation statistics (calibrated test weights)
_CALIBRATE = False
# oracle/train_oracle.py flips this for the training branch: nn.BatchNorm2d in train() mode (batch statistics,
# running statistics updated in place with momentum 0.1)
_BN_TRAIN = False


def _bn(sd: StateDict, prefix: str, x):
    """nn.BatchNorm2d in eval with running statistics; identity once `fuse_conv_bn`
    (utils/fuse_conv_bn.py:47-48) has replaced the slot by nn.Identity (no keys left)."""
    if prefix + ".running_mean" not in sd:
        return x
    if _BN_TRAIN:
        return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                            sd[prefix + ".weight"], sd[prefix + ".bias"], True, 0.1, BN_EPS)
    if _CALIBRATE:   # batch statistics, written back into the running buffers (momentum 1)
        return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                            sd[prefix + ".weight"], sd[prefix + ".bias"], True, 1.0, BN_EPS)
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, BN_EPS)


def channel_shuffle(x, groups: int = 2):
    """backbone/shufflenetv2.py:14-28: out[2i] = x[i], out[2i+1] = x[C/2+i]."""
    b, c, h, w = x.shape
    return x.view(b, groups, c // groups, h, w).transpose(1, 2).contiguous().view(b, c, h, w)


def shuffle_block(sd: StateDict, p: str, x, stride: int):
    """ShuffleV2Block.forward (backbone/shufflenetv2.py:69-78)."""
    def branch2(t):
        t = F.relu(_bn(sd, p + ".branch2.1", _conv(sd, p + ".branch2.0", t)))
        t = _bn(sd, p + ".branch2.4", _conv(sd, p + ".branch2.3", t, stride, 1, t.shape[1]))
        return F.relu(_bn(sd, p + ".branch2.6", _conv(sd, p + ".branch2.5", t)))

    if stride == 1:
        x1, x2 = x.chunk(2, dim=1)
        out = torch.cat((x1, branch2(x2)), dim=1)
    else:
        t = _bn(sd, p + ".branch1.1", _conv(sd, p + ".branch1.0", x, stride, 1, x.shape[1]))
        b1 = F.relu(_bn(sd, p + ".branch1.3", _conv(sd, p + ".branch1.2", t)))
        out = torch.cat((b1, branch2(x)), dim=1)
    return channel_shuffle(out, 2)


def conv_module(sd: StateDict, p: str, x, k: int, groups: int = 1):
    """`Conv` = Conv2d(bias) + BN + LeakyReLU(0.1) (utils/modules.py:8-18)."""
    y = _bn(sd, p + ".convs.1", _conv(sd, p + ".convs.0", x, 1, k // 2, groups))
    return F.leaky_relu(y, 0.1)


def head(sd: StateDict, p: str, x):
    """head_det_* Sequential (models/yolo_nano.py:50-70)."""
    c = x.shape[1]
    x = conv_module(sd, p + ".0", x, 3, c)
    x = conv_module(sd, p + ".1", x, 1)
    x = conv_module(sd, p + ".2", x, 3, c)
    x = conv_module(sd, p + ".3", x, 1)
    return _conv(sd, p + ".4", x)


# --------------------------------------------------------------------------------------
# network
# --------------------------------------------------------------------------------------
@torch.no_grad()
def network(sd: StateDict, x: torch.Tensor, taps: Optional[dict] = None):
    """Backbone + neck + heads (backbone/shufflenetv2.py:157-167,
    models/yolo_nano.py:284-301).  Returns raw (pred_s, pred_m, pred_l) in NCHW.
    If `taps` is a dict it receives every stage boundary named as ynb_read_tap names."""
    return network_graph(sd, x, taps)


def network_graph(sd: StateDict, x: torch.Tensor, taps: Optional[dict] = None):
    """`network` without the no_grad guard (the training oracle differentiates through it)."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    x = F.relu(_bn(sd, "backbone.conv1.1", _conv(sd, "backbone.conv1.0", x, 2, 1)))
    tap("conv1", x)
    x = tap("pool", F.max_pool2d(x, 3, 2, 1))
    feats = []
    for stage, repeats in STAGES:
        for i in range(repeats):
            x = tap(f"stage{stage}.{i}", shuffle_block(sd, f"backbone.stage{stage}.{i}", x, 2 if i == 0 else 1))
        feats.append(x)
    c3, c4, c5 = feats
    tap("c3", c3), tap("c4", c4), tap("c5", c5)

    p3 = conv_module(sd, "conv1x1_0", c3, 1)
    p4 = conv_module(sd, "conv1x1_1", c4, 1)
    p5 = conv_module(sd, "conv1x1_2", c5, 1)
    tap("lat3", p3), tap("lat4", p4), tap("lat5", p5)
    # FPN, then PAN on the *updated* maps (models/yolo_nano.py:291-296); F.interpolate
    # default mode is nearest.
    p4 = tap("fpn4", conv_module(sd, "smooth_0", p4 + F.interpolate(p5, scale_factor=2.0), 3))
    p3 = tap("p3", conv_module(sd, "smooth_1", p3 + F.interpolate(p4, scale_factor=2.0), 3))
    p4 = tap("p4", conv_module(sd, "smooth_2", p4 + F.interpolate(p3, scale_factor=0.5), 3))
    p5 = tap("p5", conv_module(sd, "smooth_3", p5 + F.interpolate(p4, scale_factor=0.5), 3))

    return (tap("pred_s", head(sd, "head_det_1", p3)),
            tap("pred_m", head(sd, "head_det_2", p4)),
            tap("pred_l", head(sd, "head_det_3", p5)))


# --------------------------------------------------------------------------------------
# decode
# --------------------------------------------------------------------------------------
def grid_tensors(input_size: int, anchor_size, num_anchors: int = 3):
    """create_grid (models/yolo_nano.py:86-112): grid_xy [1,HW,1,2] (x=col, y=row),
    stride [1,HW,A,2], anchors [1,HW,A,2], concatenated over the 3 levels."""
    anchors = torch.tensor(anchor_size, dtype=torch.float32).view(3, num_anchors, 2)
    g, st, an = [], [], []
    for lvl, s in enumerate(STRIDES):
        n = input_size // s
        ys, xs = torch.meshgrid(torch.arange(n), torch.arange(n), indexing="ij")
        g.append(torch.stack([xs, ys], dim=-1).float().view(1, n * n, 1, 2))
        st.append(torch.ones([1, n * n, num_anchors, 2]) * s)
        an.append(anchors[lvl].repeat(n * n, 1, 1))
    return torch.cat(g, 1), torch.cat(st, 1), torch.cat(an, 0).unsqueeze(0)


@torch.no_grad()
def split_predictions(preds: Sequence[torch.Tensor], num_classes: int, num_anchors: int = 3):
    """Re-layout of models/yolo_nano.py:303-330: NCHW -> [B,HW,A(1+C+4)], then
    conf [B,N,1] | cls [B,N,C] | txtytwth [B,sumHW*A... ,4] concatenated over levels."""
    confs, clss, boxes = [], [], []
    a, c = num_anchors, num_classes
    for pred in preds:
        b, ch, h, w = pred.shape
        pred = pred.permute(0, 2, 3, 1).contiguous().view(b, h * w, ch)
        confs.append(pred[:, :, :a].contiguous().view(b, h * w * a, 1))
        clss.append(pred[:, :, a:(1 + c) * a].contiguous().view(b, h * w * a, c))
        boxes.append(pred[:, :, (1 + c) * a:].contiguous())
    return torch.cat(confs, 1), torch.cat(clss, 1), torch.cat(boxes, 1)


@torch.no_grad()
def decode(preds, input_size: int, num_classes: int, anchor_size, num_anchors: int = 3):
    """Eval branch of forward for EVERY image (the reference does image 0 only,
    models/yolo_nano.py:362-367): returns all_bbox [B,N,4] clamped to [0,1] and
    all_class [B,N,C] = softmax(cls) * sigmoid(conf)."""
    conf, cls, txtytwth = split_predictions(preds, num_classes, num_anchors)
    b = conf.shape[0]
    grid, stride, anchors = grid_tensors(input_size, anchor_size, num_anchors)
    t = txtytwth.view(b, -1, num_anchors, 4)
    # decode_xywh (models/yolo_nano.py:120-136)
    cxy = (torch.sigmoid(t[..., :2]) + grid) * stride
    wh = torch.exp(t[..., 2:]) * anchors
    xywh = torch.cat([cxy, wh], -1).view(b, -1, 4)
    # decode_boxes (models/yolo_nano.py:139-156)
    box = torch.zeros_like(xywh)
    box[:, :, 0] = xywh[:, :, 0] - xywh[:, :, 2] / 2
    box[:, :, 1] = xywh[:, :, 1] - xywh[:, :, 3] / 2
    box[:, :, 2] = xywh[:, :, 0] + xywh[:, :, 2] / 2
    box[:, :, 3] = xywh[:, :, 1] + xywh[:, :, 3] / 2
    all_bbox = torch.clamp(box / input_size, 0.0, 1.0)
    all_obj = torch.sigmoid(conf)
    all_class = torch.softmax(cls, dim=2) * all_obj
    return all_bbox, all_class


# --------------------------------------------------------------------------------------
# post-processing (NumPy, host) — models/yolo_nano.py:159-279
# --------------------------------------------------------------------------------------
def sort_order(scores: np.ndarray, tie: str) -> np.ndarray:
    """Visiting order of NMS.

    'numpy': exactly the reference expression `scores.argsort()[::-1]`
             (models/yolo_nano.py:167) — NumPy's default sort is unstable, the order of
             equal scores is implementation-defined (SURVEY §8c hazard 2).
    'index': the order this build defines and the CUDA kernels implement: score
             descending, equal scores by ascending index."""
    if tie == "numpy":
        return scores.argsort()[::-1]
    if tie == "index":
        return np.lexsort((np.arange(scores.shape[0]), -scores.astype(np.float64)))
    raise ValueError(tie)


def nms(dets: np.ndarray, scores: np.ndarray, nms_thresh: float, tie: str = "index") -> List[int]:
    """Greedy NMS, models/yolo_nano.py:159-188.  float32 arithmetic, clamp 1e-28,
    survivors are `ovr <= thresh` (NaN is suppressed)."""
    x1, y1, x2, y2 = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    order = sort_order(scores, tie)
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        rest = order[1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(1e-28, xx2 - xx1)
        h = np.maximum(1e-28, yy2 - yy1)
        inter = w * h
        ovr = inter / (areas[i] + areas[rest] - inter)
        order = rest[np.where(ovr <= nms_thresh)[0]]
    return keep


def diou_nms(dets: np.ndarray, scores: np.ndarray, nms_thresh: float, tie: str = "index") -> List[int]:
    """Greedy DIoU-NMS, models/yolo_nano.py:191-242: ovr = iou - D^2 / (C^2 + 1e-20)."""
    x1, y1, x2, y2 = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    order = sort_order(scores, tie)
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        rest = order[1:]
        n = len(rest)
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(1e-28, xx2 - xx1)
        h = np.maximum(1e-28, yy2 - yy1)
        inter = w * h
        iou = inter / (areas[i] + areas[rest] - inter)
        x1_, x2_ = x1[i].repeat(n), x2[i:i + 1].repeat(n)
        y1_, y2_ = y1[i].repeat(n), y2[i:i + 1].repeat(n)
        x1234 = np.stack([x1_, x2_, x1[rest], x2[rest]], axis=1)
        y1234 = np.stack([y1_, y2_, y1[rest], y2[rest]], axis=1)
        c = np.sqrt((np.max(x1234, axis=1) - np.min(x1234, axis=1)) ** 2
                    + (np.max(y1234, axis=1) - np.min(y1234, axis=1)) ** 2)
        p1x, p1y = (x1_ + x2_) / 2.0, (y1_ + y2_) / 2.0
        p2x, p2y = (x1[rest] + x2[rest]) / 2.0, (y1[rest] + y2[rest]) / 2.0
        d = np.sqrt((p2x - p1x) ** 2 + (p2y - p1y) ** 2)
        ovr = iou - d ** 2 / (c ** 2 + 1e-20)
        order = rest[np.where(ovr <= nms_thresh)[0]]
    return keep


def class_scores(all_class: np.ndarray):
    """models/yolo_nano.py:253-256: argmax over classes (first max wins) and its score."""
    cls_inds = np.argmax(all_class, axis=1)
    scores = all_class[(np.arange(all_class.shape[0]), cls_inds)].copy()
    return scores, cls_inds


def postprocess_flat(bbox: np.ndarray, scores: np.ndarray, cls_inds: np.ndarray, num_classes: int,
                     conf_thresh: float, nms_thresh: float, use_diou: bool = False,
                     tie: str = "index"):
    """Threshold + per-class NMS (models/yolo_nano.py:258-279) on already reduced
    (score, class) pairs.  Returns (bboxes, scores, cls_inds, kept_anchor_indices)."""
    sel = np.where(scores >= conf_thresh)[0]
    bbox, scores, cls_inds = bbox[sel], scores[sel], cls_inds[sel]
    keep = np.zeros(len(bbox), dtype=np.int64)
    fn = diou_nms if use_diou else nms
    for c in range(num_classes):
        inds = np.where(cls_inds == c)[0]
        if len(inds) == 0:
            continue
        c_keep = fn(bbox[inds], scores[inds], nms_thresh, tie)
        keep[inds[c_keep]] = 1
    k = np.where(keep > 0)[0]
    return bbox[k], scores[k], cls_inds[k].astype(np.int64), sel[k]


def postprocess(all_bbox: np.ndarray, all_class: np.ndarray, num_classes: int, conf_thresh: float,
                nms_thresh: float, use_diou: bool = False, tie: str = "index"):
    """YOLONano.postprocess (models/yolo_nano.py:245-279) for one image."""
    scores, cls_inds = class_scores(all_class)
    return postprocess_flat(all_bbox, scores, cls_inds, num_classes, conf_thresh, nms_thresh,
                            use_diou, tie)


# --------------------------------------------------------------------------------------
# whole path
# --------------------------------------------------------------------------------------
@torch.no_grad()
def detect(sd: StateDict, x: torch.Tensor, input_size: int, num_classes: int, anchor_size,
           conf_thresh: float = 0.001, nms_thresh: float = 0.5, use_diou: bool = False,
           tie: str = "index"):
    """YOLONano.forward in eval mode (models/yolo_nano.py:282-376), looped over the
    batch one image at a time — the defined batch extension (SURVEY §8c hazard 4).
    Returns a list of (bboxes [K,4] f32, scores [K] f32, cls_inds [K] i64)."""
    out = []
    for i in range(x.shape[0]):
        preds = network(sd, x[i:i + 1])
        all_bbox, all_class = decode(preds, input_size, num_classes, anchor_size)
        b, s, c, _ = postprocess(all_bbox[0].numpy(), all_class[0].numpy(), num_classes,
                                 conf_thresh, nms_thresh, use_diou, tie)
        out.append((b, s, c))
    return out


def fold_state_dict(sd: StateDict, table) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
    """_fuse_conv_bn (utils/fuse_conv_bn.py:6-22) applied to a state_dict: per conv
    (W', b').  `table` is yolo_nano_b200.topology.conv_table()."""
    out = {}
    for spec in table:
        w = sd[spec.name + ".weight"]
        b = sd.get(spec.name + ".bias")
        if spec.bn is not None and spec.bn + ".running_mean" in sd:
            mean, var = sd[spec.bn + ".running_mean"], sd[spec.bn + ".running_var"]
            if b is None:
                b = torch.zeros_like(mean)
            factor = sd[spec.bn + ".weight"] / torch.sqrt(var + BN_EPS)
            w = w * factor.reshape([w.shape[0], 1, 1, 1])
            b = (b - mean) * factor + sd[spec.bn + ".bias"]
        elif b is None:
            b = torch.zeros(w.shape[0])
        out[spec.name] = (w, b)
    return out


# ---------------------------------------------------------------------------------------------
# Pre-processing: the tail of ValTransforms, data/transforms.py:445-458
VAL_MEAN_BGR = (0.406, 0.456, 0.485)
VAL_STD_BGR = (0.225, 0.224, 0.229)


def preprocess_u8(canvas: np.ndarray, rect=None, mean=VAL_MEAN_BGR, std=VAL_STD_BGR) -> np.ndarray:
    """uint8 [S,S,3] BGR canvas (content already resized by cv2 on the host) -> float32 [3,S,S] RGB.
    Resize padding (:84-88,98-102: `np.ones(...) * mean*255`, float64), Normalize (:59-70: astype
    float32, /= 255., -= mean, /= std), ToTensor (:394-398: BGR->RGB, HWC->CHW)."""
    mean = np.array(mean, dtype=np.float32)
    std = np.array(std, dtype=np.float32)
    image = canvas.astype(np.float32)
    if rect is not None:
        x0, y0, w, h = (int(v) for v in rect)
        padded = np.ones(canvas.shape) * np.array([v * 255 for v in mean])
        padded[y0:y0 + h, x0:x0 + w, :] = image[y0:y0 + h, x0:x0 + w, :]
        image = padded
    image = image.astype(np.float32)
    image /= 255.
    image -= mean
    image /= std
    image = image[..., (2, 1, 0)]
    return np.ascontiguousarray(np.transpose(image, (2, 0, 1))).astype(np.float32)


def cv2_resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """`cv2.resize(src, (dw, dh))` for uint8 HWC images (default INTER_LINEAR), restated from OpenCV 4.x
    imgproc/src/resize.cpp (third-party code of the reference's call at data/transforms.py:80,94,108; not pinned by
    the reference — installed here: opencv 4.13): source index and weight per destination index in float32
    (`fx = (float)((dx + 0.5) * scale - 0.5)`), 11-bit fixed-point weights (cvRound), horizontal pass in int32,
    vertical pass `(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2`; an exact 2x downscale in both
    directions takes the INTER_AREA fast path (2x2 mean, rounded).  Checked bit-exact against cv2 in
    tests/test_oracle_golden.py."""
    sh, sw = src.shape[:2]

    def coeffs(dn, sn):
        scale = 1.0 / (np.float64(dn) / np.float64(sn))
        d = np.arange(dn, dtype=np.float64)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        return s, (f - s.astype(np.float32)).astype(np.float32), scale

    sx, fx, scale_x = coeffs(dw, sw)
    sy, fy, scale_y = coeffs(dh, sh)
    eps = np.finfo(np.float64).eps
    if abs(scale_x - 2) < eps and abs(scale_y - 2) < eps:
        s = src.astype(np.int32)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    neg = sx < 0
    fx = np.where(neg, np.float32(0), fx); sx = np.where(neg, 0, sx)
    big = sx >= sw - 1
    fx = np.where(big, np.float32(0), fx); sx = np.where(big, sw - 1, sx)
    rnd = lambda v: np.rint(v.astype(np.float32)).astype(np.int64)          # cvRound: half to even
    a0, a1 = rnd((np.float32(1.0) - fx) * np.float32(2048)), rnd(fx * np.float32(2048))
    b0, b1 = rnd((np.float32(1.0) - fy) * np.float32(2048)), rnd(fy * np.float32(2048))
    x1 = np.minimum(sx + 1, sw - 1)
    y0, y1 = np.clip(sy, 0, sh - 1), np.clip(sy + 1, 0, sh - 1)
    S = src.astype(np.int64)
    H = S[:, sx, :] * a0[None, :, None] + S[:, x1, :] * a1[None, :, None]
    out = (((b0[:, None, None] * (H[y0] >> 4)) >> 16) + ((b1[:, None, None] * (H[y1] >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def letterbox_geometry(h0: int, w0: int, size: int):
    """Resize (data/transforms.py:73-119): (new_w, new_h, left, top) of the content on the size x size canvas, and
    the (scale, offset) the evaluators use to map boxes back (float64, as the reference builds them)."""
    if h0 > w0:
        nw, nh = int(w0 / h0 * size), size
        left, top = (nh - nw) // 2, 0
        offset = np.array([[left / nh, 0., left / nh, 0.]])
        scale = np.array([[nw / nh, 1., nw / nh, 1.]])
    elif h0 < w0:
        nw, nh = size, int(h0 / w0 * size)
        left, top = 0, (nw - nh) // 2
        offset = np.array([[0., top / nw, 0., top / nw]])
        scale = np.array([1., nh / nw, 1., nh / nw])
    else:
        nw, nh, left, top = size, size, 0, 0
        offset = np.zeros([1, 4])
        scale = 1.
    return nw, nh, left, top, scale, offset


def val_transform(img: np.ndarray, size: int, mean=VAL_MEAN_BGR, std=VAL_STD_BGR):
    """ValTransforms (data/transforms.py:445-458) = Resize (letterbox, cv2 bilinear, padding mean*255) + Normalize +
    ToTensor on a uint8 BGR image of any shape: (x float32 [3,S,S] RGB, scale, offset)."""
    h0, w0 = img.shape[:2]
    nw, nh, left, top, scale, offset = letterbox_geometry(h0, w0, size)
    if h0 == w0 and h0 == size:
        content = img
    else:
        content = cv2_resize_linear_u8(img, nw, nh)
    canvas = np.zeros((size, size, 3), np.uint8)
    canvas[top:top + nh, left:left + nw] = content
    rect = None if h0 == w0 else (left, top, nw, nh)
    return preprocess_u8(canvas, rect, mean, std), scale, offset


def map_boxes_to_image(bboxes: np.ndarray, scale, offset, w: int, h: int) -> np.ndarray:
    """evaluator/cocoapi_evaluator.py:85-87 (test.py:133-135): in-place float32 ops with float64 operands."""
    b = np.array(bboxes, dtype=np.float32, copy=True)
    b -= offset
    b /= scale
    b *= np.array([[w, h, w, h]])
    return b


# ---------------------------------------------------------------------------------------------
# Test-time augmentation: utils/misc.py:90-148
def tta_merge(bboxes: np.ndarray, scores: np.ndarray, labels: np.ndarray, num_classes: int, nms_thresh: float,
              tie: str = "index"):
    """Second per-class NMS over the union (utils/misc.py:131-146; nms of utils/misc.py:8-37 is the
    same loop as models/yolo_nano.py:159-188)."""
    keep = np.zeros(len(bboxes), dtype=np.int64)
    for i in range(num_classes):
        inds = np.where(labels == i)[0]
        if len(inds) == 0:
            continue
        c_keep = nms(bboxes[inds], scores[inds], nms_thresh, tie)
        keep[inds[c_keep]] = 1
    keep = np.where(keep > 0)
    return bboxes[keep], scores[keep], labels[keep]


@torch.no_grad()
def tta(sd: StateDict, x: torch.Tensor, num_classes: int, anchor_size, scales, nms_thresh: float = 0.4,
        conf_thresh: float = 0.001, model_nms_thresh: float = 0.5, tie: str = "index"):
    """TestTimeAugmentation.__call__ (utils/misc.py:97-148): for every scale, the image and its
    horizontal flip through the model (image 0 only), boxes of the flip mirrored back, union, NMS."""
    import torch.nn.functional as F
    bl, sl, ll = [], [], []
    for s in scales:
        s = int(s)
        xs = x if (x.size(-1) == s and x.size(-2) == s) else F.interpolate(x, size=(s, s), mode="bilinear",
                                                                               align_corners=False)
        b, sc, lb = detect(sd, xs[:1], s, num_classes, anchor_size, conf_thresh, model_nms_thresh, False, tie)[0]
        bl.append(b); sl.append(sc); ll.append(lb)
        b, sc, lb = detect(sd, torch.flip(xs, [-1])[:1], s, num_classes, anchor_size, conf_thresh, model_nms_thresh,
                           False, tie)[0]
        b = b.copy()
        b[:, 0::2] = 1.0 - b[:, 2::-2]
        bl.append(b); sl.append(sc); ll.append(lb)
    return tta_merge(np.concatenate(bl), np.concatenate(sl), np.concatenate(ll), num_classes, nms_thresh, tie)
