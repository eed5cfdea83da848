"""Training branch at the head boundary through the C ABI on a real
B200, against the oracle restatement (oracle/train_oracle.py) and the fixture recorded from the REAL
reference (tests/golden/g7_train128.npz)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import os  # noqa: E402
os.environ.setdefault("YNB_SYNC_CHECK", "1")   # surface bounded-wait timeouts of the tcgen05 weight-gradient kernel

from oracle import train_oracle as T  # noqa: E402
from oracle import weights as W  # noqa: E402

# float32 tolerances of this stage: the losses are sums of 1e3..1e5 float32 terms (torch sums in
# float32, the kernel combines block partials in float64); gradients are per-element closed forms
# (expf / logf / division rounding differences only).
LOSS_RTOL = 2e-5
GRAD_RTOL, GRAD_ATOL = 2e-5, 2e-8


@pytest.fixture(scope="module")
def G():
    import gpu_util
    return gpu_util


@pytest.fixture(scope="module")
def TR():
    from yolo_nano_b200 import training
    return training


@pytest.fixture(scope="module")
def g7(golden):
    return golden("g7_train128.npz")


def _nhwc_maps(preds, ld, dev):
    out = []
    for p in preds:
        b, ch, h, w = p.shape
        m = torch.zeros(b, h * w, ld)
        m[:, :, :ch] = p.permute(0, 2, 3, 1).reshape(b, h * w, ch)
        out.append(m.to(dev))
    return out


def _nchw_grads(grads, preds):
    out = []
    for g, p in zip(grads, preds):
        b, ch, h, w = p.shape
        out.append(g.cpu()[:, :, :ch].reshape(b, h, w, ch).permute(0, 3, 1, 2).numpy())
    return out


def _labels_f32(g7):
    lab = g7["labels"].astype(np.float32)
    return lab, g7["n_labels"].astype(np.int32)


def test_build_targets_matches_reference_fixture(G, TR, g7):
    """tools.multi_gt_creator on the device: positives, ignored (-1), dirty boxes, overwrite order."""
    lab, cnt = _labels_f32(g7)
    size, classes = int(g7["size"]), int(g7["classes"])
    got = TR.build_targets(torch.from_numpy(lab).to(G.DEV), torch.from_numpy(cnt).to(G.DEV), size,
                           W.anchors_for(classes)).cpu().numpy()
    assert np.array_equal(lab.astype(np.float64), g7["labels"])      # the fixture's labels are float32 values
    lists = [[[float(v) for v in lab[b, i]] for i in range(cnt[b])] for b in range(len(cnt))]
    want = T.multi_gt_creator(size, lists, W.anchors_for(classes)).numpy()
    assert (want[:, :, 0] > 0).sum() >= 16 and (want[:, :, 0] < 0).sum() > 0
    np.testing.assert_array_equal(got[:, :, [0, 1]], want[:, :, [0, 1]])           # obj / class: exact
    np.testing.assert_allclose(got, want, rtol=2e-7, atol=1e-7)                      # double log -> float32
    # and against the REAL reference's tools.multi_gt_creator on the same labels
    np.testing.assert_array_equal(got[:, :, [0, 1]], g7["target"][:, :, [0, 1]])
    np.testing.assert_allclose(got, g7["target"], rtol=2e-7, atol=1e-7)


@pytest.mark.parametrize("ld", [76, 80])
def test_train_loss_matches_reference_fixture(G, TR, g7, ld):
    """Losses and d(total)/d(raw head maps) against the REAL reference's forward(x, target) + backward."""
    size, classes = int(g7["size"]), int(g7["classes"])
    preds = [torch.from_numpy(g7[k]) for k in ("pred_s", "pred_m", "pred_l")]
    raw = _nhwc_maps(preds, ld, G.DEV)
    losses, grads = TR.train_loss(raw, torch.from_numpy(g7["target"]).to(G.DEV), size, classes, W.anchors_for(classes))
    np.testing.assert_allclose(losses.cpu().numpy(), g7["losses"], rtol=LOSS_RTOL)
    for k, got in zip(("pred_s", "pred_m", "pred_l"), _nchw_grads(grads, preds)):
        np.testing.assert_allclose(got, g7["grad_" + k], rtol=GRAD_RTOL, atol=GRAD_ATOL, err_msg=k)
    for g in grads:                                   # padding channels of the gradient map are zero
        if ld > preds[0].shape[1]:
            assert float(g[:, :, preds[0].shape[1]:].abs().max()) == 0.0


@pytest.mark.parametrize("size,classes,batch", [(416, 80, 4), (320, 20, 3)])
def test_train_loss_matches_oracle_at_baseline_shape(G, TR, size, classes, batch):
    """Config-5 shape (416^2, COCO-80): random logits, targets from random labels; oracle = autograd."""
    torch.manual_seed(size + classes)
    rng = np.random.RandomState(size)
    anchors = W.anchors_for(classes)
    labels = []
    for b in range(batch):
        labs = []
        for _ in range(12):
            cx, cy = rng.uniform(0.1, 0.9, 2)
            w, h = rng.uniform(0.02, 0.7, 2)
            labs.append([float(np.float32(v)) for v in (max(cx - w / 2, 0), max(cy - h / 2, 0), min(cx + w / 2, 1),
                                                         min(cy + h / 2, 1))] + [float(rng.randint(0, classes))])
        labels.append(labs)
    target = T.multi_gt_creator(size, labels, anchors)
    ch = 3 * (1 + classes + 4)
    preds = [torch.randn(batch, ch, size // s, size // s) * 1.5 for s in (8, 16, 32)]
    want_l, want_g = T.losses_and_grads(preds, target, size, classes, anchors)
    raw = _nhwc_maps(preds, (ch + 3) // 4 * 4, G.DEV)
    losses, grads = TR.train_loss(raw, target.to(G.DEV), size, classes, anchors)
    np.testing.assert_allclose(losses.cpu().numpy(), np.array(want_l, dtype=np.float32), rtol=LOSS_RTOL)
    for lvl, (got, want) in enumerate(zip(_nchw_grads(grads, preds), want_g)):
        np.testing.assert_allclose(got, want.numpy(), rtol=GRAD_RTOL, atol=GRAD_ATOL, err_msg=f"level {lvl}")
    # deterministic: a second run gives bit-identical losses and gradients
    losses2, grads2 = TR.train_loss(raw, target.to(G.DEV), size, classes, anchors)
    assert torch.equal(losses, losses2) and all(torch.equal(a, b) for a, b in zip(grads, grads2))


def test_sgd_step_bit_exact(G, TR, g7):
    """torch.optim.SGD(momentum 0.9, weight decay 5e-4), two steps, vs the recorded CPU optimiser."""
    p = torch.from_numpy(g7["sgd_p0"]).to(G.DEV)
    opt = TR.FlatSGD(p, lr=1e-3)
    opt.step(torch.from_numpy(g7["sgd_g1"]).to(G.DEV))
    np.testing.assert_array_equal(p.cpu().numpy(), g7["sgd_p1"])
    opt.step(torch.from_numpy(g7["sgd_g2"]).to(G.DEV))
    np.testing.assert_array_equal(p.cpu().numpy(), g7["sgd_p2"])
    with pytest.raises(Exception):
        TR.FlatSGD(torch.zeros(8), lr=0.1)            # CPU tensor: no fallback


@pytest.mark.parametrize("batch,ch,hw,stride", [(2, 58, 52, 1), (2, 58, 104, 2), (3, 116, 26, 1), (2, 232, 13, 1),
                                                (2, 24, 104, 2), (4, 96, 13, 1), (1, 4, 3, 2), (1, 5, 7, 1)])
def test_dwconv3x3_backward(G, TR, batch, ch, hw, stride):
    torch.manual_seed(ch + hw)
    x = torch.randn(batch, ch, hw, hw, requires_grad=True)
    w = torch.randn(ch, 1, 3, 3, requires_grad=True)
    b = torch.randn(ch, requires_grad=True)
    y = F.conv2d(x, w, b, stride, 1, 1, ch)
    dy = torch.randn_like(y)
    y.backward(dy)
    nhwc = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().to(G.DEV)  # noqa: E731
    dx, dw, db = TR.dwconv3x3_backward(nhwc(dy), nhwc(x), w.detach().view(ch, 9).t().contiguous().to(G.DEV), stride)
    torch.testing.assert_close(dx.cpu().permute(0, 3, 1, 2), x.grad, rtol=1e-5, atol=1e-5)
    scale = float(w.grad.abs().max())
    torch.testing.assert_close(dw.cpu().t().reshape(ch, 1, 3, 3), w.grad, rtol=1e-4, atol=1e-5 * scale)
    torch.testing.assert_close(db.cpu(), b.grad, rtol=1e-4, atol=1e-5 * float(b.grad.abs().max()))


@pytest.mark.parametrize("m,k,n", [(2 * 2704, 116, 116), (2 * 676, 232, 232), (338, 464, 96), (5408, 96, 256),
                                   (1000, 24, 60), (33, 8, 4), (64 * 2704, 60, 60)])
def test_pwconv_backward(G, TR, m, k, n):
    torch.manual_seed(m + k + n)
    x = torch.randn(m, k, requires_grad=True)
    w = (torch.randn(n, k) / k ** 0.5).requires_grad_(True)
    b = torch.randn(n, requires_grad=True)
    y = F.linear(x, w, b)
    dy = torch.randn_like(y)
    y.backward(dy)
    dw, db = TR.pwconv_backward_weight(dy.to(G.DEV), x.detach().to(G.DEV))
    sw, sb = float(w.grad.abs().max()), float(b.grad.abs().max())
    torch.testing.assert_close(dw.cpu(), w.grad, rtol=1e-4, atol=2e-5 * sw)
    torch.testing.assert_close(db.cpu(), b.grad, rtol=1e-4, atol=2e-5 * sb)
    for tc in (False, True):
        dx = TR.pwconv_backward_data(dy.to(G.DEV), w.detach().to(G.DEV), tensor_cores=tc)
        torch.testing.assert_close(dx.cpu(), x.grad, rtol=1e-4, atol=1e-4 * float(x.grad.abs().max()))


@pytest.mark.parametrize("act", [1, 2])
def test_act_backward(G, TR, act):
    torch.manual_seed(act)
    pre = torch.randn(777, 58, requires_grad=True)
    out = F.relu(pre) if act == 1 else F.leaky_relu(pre, 0.1)
    dy = torch.randn_like(out)
    out.backward(dy)
    got = TR.act_backward(dy.to(G.DEV), out.detach().to(G.DEV), act)
    np.testing.assert_array_equal(got.cpu().numpy(), pre.grad.numpy())


def test_forward_with_target_matches_reference_end_to_end(G, g7):
    """model(x, target) with trainable=True on running-statistics BatchNorm: input images -> four losses
    and head gradients, against the REAL reference run in the same state (fixture g7).  The network part
    is the tensor-core path (3xTF32, raw head outputs within 1e-3 + 1e-4 rel), hence the looser bar."""
    import contextlib, io
    import yolo_nano_b200 as pkg
    size, classes, seed = int(g7["size"]), int(g7["classes"]), int(g7["seed"])
    sd = W.calibrated(classes, seed=seed)
    x = W.synthetic_input(2, size, seed=seed)
    assert W.digest(sd) == str(g7["sd_digest"]) and W.digest(x) == str(g7["x_digest"])
    with contextlib.redirect_stdout(io.StringIO()):
        m = pkg.YOLONano(G.DEV, size, classes, anchor_size=W.anchors_for(classes))
    m.load_state_dict(sd)
    m = m.to(G.DEV)
    m.trainable = True
    keep = {k: v.clone() for k, v in m.state_dict().items()}
    lt = m.train()(x.to(G.DEV), target=torch.from_numpy(g7["target"]).to(G.DEV))      # batch statistics + backward chain
    assert len(lt) == 4 and all(bool(torch.isfinite(v)) for v in lt)
    assert len(m.gradients) == 247 and all(tuple(m.gradients[k].shape) == tuple(p.shape) for k, p in m.named_parameters())
    m.load_state_dict(keep)                            # the train() call moved the running statistics
    m.eval()
    ls = m(x.to(G.DEV), target=torch.from_numpy(g7["target"]).to(G.DEV))
    got = np.array([float(v) for v in ls], dtype=np.float32)
    np.testing.assert_allclose(got, g7["losses"], rtol=2e-3)
    ch = 3 * (1 + classes + 4)
    for k, g in zip(("pred_s", "pred_m", "pred_l"), m.head_gradients):
        want = g7["grad_" + k]
        b, _, h, w = want.shape
        gg = g.cpu()[:, :, :ch].reshape(b, h, w, ch).permute(0, 3, 1, 2).numpy()
        assert np.abs(gg - want).max() <= 3e-3 * np.abs(want).max(), k
    # same weights, same input, the eval branch still works on the same model object
    m.trainable = False
    bboxes, scores, cls_inds = m(x[:1].to(G.DEV))
    assert bboxes.shape[1] == 4 and len(scores) == len(cls_inds)


@pytest.mark.parametrize("m,c,act", [(2 * 2704, 116, 1), (3 * 676, 232, 0), (4 * 169, 464, 1), (5408, 96, 2), (7, 4, 2),
                                     (64 * 2704, 60, 1)])
def test_batchnorm_training_mode(G, TR, m, c, act):
    """nn.BatchNorm2d(train) + activation, forward / running statistics / backward vs torch autograd."""
    torch.manual_seed(m + c)
    x = (torch.randn(m, c) * 1.7 + 0.6).requires_grad_(True)
    bn = torch.nn.BatchNorm1d(c)          # same arithmetic as BatchNorm2d on [M, C] = (N*H*W, C)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.3)
        bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 2.0)
    mine = TR.BatchNormTrain(bn.weight.detach().clone().to(G.DEV), bn.bias.detach().clone().to(G.DEV),
                             bn.running_mean.clone().to(G.DEV), bn.running_var.clone().to(G.DEV), act=act)
    bn.train()
    pre = bn(x)
    y = pre if act == 0 else (F.relu(pre) if act == 1 else F.leaky_relu(pre, 0.1))
    dy = torch.randn_like(y)
    y.backward(dy)
    got = mine.forward(x.detach().to(G.DEV))
    torch.testing.assert_close(got.cpu(), y.detach(), rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(mine.running_mean.cpu(), bn.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(mine.running_var.cpu(), bn.running_var, rtol=1e-5, atol=1e-6)
    dx, dg, db = mine.backward(dy.to(G.DEV))
    # elements whose pre-activation sits within rounding of 0 may take the other branch of act'
    near0 = pre.detach().abs() < 1e-5
    scale = float(x.grad.abs().max())
    diff = (dx.cpu() - x.grad).abs()
    assert float(diff[~near0].max()) <= 2e-4 * scale
    torch.testing.assert_close(dg.cpu(), bn.weight.grad, rtol=2e-4, atol=2e-4 * float(bn.weight.grad.abs().max()))
    torch.testing.assert_close(db.cpu(), bn.bias.grad, rtol=2e-4, atol=2e-4 * float(bn.bias.grad.abs().max()))


def _bn_layer(TR, G, bn, act):
    return TR.BnActTrain(bn.weight.detach().clone().to(G.DEV), bn.bias.detach().clone().to(G.DEV),
                         bn.running_mean.clone().to(G.DEV), bn.running_var.clone().to(G.DEV), eps=bn.eps,
                         momentum=bn.momentum, act=act)


def _check_block(G, TR, ref, layers, x, tol=3e-4):
    """ref: torch module in train mode (NCHW); layers: SequentialTrain (NHWC).  Output, input gradient and
    every parameter gradient, each within `tol` of its own scale."""
    x = x.clone().requires_grad_(True)
    ref.train()
    y = ref(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    nhwc = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().to(G.DEV)  # noqa: E731
    got = layers.forward(nhwc(x))
    err = lambda a, b: float((a - b).abs().max() / b.abs().max())  # noqa: E731
    assert err(got.cpu().permute(0, 3, 1, 2), y.detach()) <= tol
    dx = layers.backward(nhwc(dy))
    assert err(dx.cpu().permute(0, 3, 1, 2), x.grad) <= tol
    return layers.grads(), err


def test_conv_module_train_step_matches_autograd(G, TR):
    """utils/modules.py:8-18 `Conv(c1, c2, k=1)` in TRAINING mode: conv(+bias) -> BN(batch stats) -> LeakyReLU,
    forward + backward composed from the kernels (tcgen05 GEMM, BN, weight-gradient, W^T GEMM)."""
    torch.manual_seed(11)
    c1, c2 = 232, 96
    conv, bn = torch.nn.Conv2d(c1, c2, 1), torch.nn.BatchNorm2d(c2)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.3)
    ref = torch.nn.Sequential(conv, bn, torch.nn.LeakyReLU(0.1))
    layers = TR.SequentialTrain(TR.PwConvTrain(conv.weight.detach().view(c2, c1).contiguous().to(G.DEV),
                                               conv.bias.detach().clone().to(G.DEV)), _bn_layer(TR, G, bn, 2))
    grads, err = _check_block(G, TR, ref, layers, torch.randn(3, c1, 26, 26))
    assert err(grads[0]["weight"].cpu().view_as(conv.weight), conv.weight.grad) <= 3e-4
    assert err(grads[1]["weight"].cpu(), bn.weight.grad) <= 3e-4 and err(grads[1]["bias"].cpu(), bn.bias.grad) <= 3e-4
    # the conv bias gradient is ~0 by construction (BatchNorm removes the mean): compare on the weight-grad scale
    assert float((grads[0]["bias"].cpu() - conv.bias.grad).abs().max()) <= 3e-4 * float(conv.weight.grad.abs().max()) * 26


def test_shuffle_branch2_train_step_matches_autograd(G, TR):
    """backbone/shufflenetv2.py:53-63 branch2 of a stride-1 stage-3 unit in TRAINING mode:
    pw -> BN -> ReLU -> dw3x3 -> BN -> pw -> BN -> ReLU, forward + backward, all parameter gradients."""
    torch.manual_seed(12)
    c = 116
    pw1, bn1 = torch.nn.Conv2d(c, c, 1, bias=False), torch.nn.BatchNorm2d(c)
    dw, bn2 = torch.nn.Conv2d(c, c, 3, 1, 1, groups=c, bias=False), torch.nn.BatchNorm2d(c)
    pw2, bn3 = torch.nn.Conv2d(c, c, 1, bias=False), torch.nn.BatchNorm2d(c)
    with torch.no_grad():
        for bn in (bn1, bn2, bn3):
            bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.3)
    ref = torch.nn.Sequential(pw1, bn1, torch.nn.ReLU(), dw, bn2, pw2, bn3, torch.nn.ReLU())
    dev = G.DEV
    layers = TR.SequentialTrain(
        TR.PwConvTrain(pw1.weight.detach().view(c, c).contiguous().to(dev)), _bn_layer(TR, G, bn1, 1),
        TR.DwConvTrain(dw.weight.detach().view(c, 9).t().contiguous().to(dev), 1), _bn_layer(TR, G, bn2, 0),
        TR.PwConvTrain(pw2.weight.detach().view(c, c).contiguous().to(dev)), _bn_layer(TR, G, bn3, 1))
    grads, err = _check_block(G, TR, ref, layers, torch.randn(4, c, 26, 26), tol=1e-3)
    assert err(grads[0]["weight"].cpu().view_as(pw1.weight), pw1.weight.grad) <= 1e-3
    assert err(grads[2]["weight"].cpu().t().reshape(c, 1, 3, 3), dw.weight.grad) <= 1e-3
    assert err(grads[4]["weight"].cpu().view_as(pw2.weight), pw2.weight.grad) <= 1e-3
    for i, bn in ((1, bn1), (3, bn2), (5, bn3)):
        assert err(grads[i]["weight"].cpu(), bn.weight.grad) <= 1e-3
        # bn2's bias gradient is analytically 0 (a constant shift goes through the linear pw2 and is removed
        # by bn3): both sides are rounding noise there, so the bias gradients are compared on the scale of the
        # layer's weight gradient
        assert float((grads[i]["bias"].cpu() - bn.bias.grad).abs().max()) <= 1e-3 * float(bn.weight.grad.abs().max())
    # running statistics moved exactly as torch moves them
    torch.testing.assert_close(layers.layers[1].running_var.cpu(), bn1.running_var, rtol=1e-4, atol=1e-6)


def test_train_loss_full_config5_batch_is_additive(G, TR):
    """BASELINE config 5's per-GPU shape (32 images, 416^2, COCO-80), size-independent property: the loss sums
    are additive over images and the gradient of an anchor depends on its own image only.  With power-of-two
    batch sizes the 1/B factors are exact, so the B=32 gradients equal 1/4 of the B=8 gradients of the same
    images BIT FOR BIT, and 32 * loss(B=32) equals the sum of 8 * loss(B=8) over the four sub-batches."""
    torch.manual_seed(5)
    size, classes, batch, ld = 416, 80, 32, 256
    anchors = W.anchors_for(classes)
    grids = [size // s for s in (8, 16, 32)]
    labels = torch.rand(batch, 24, 5, device=G.DEV)
    xy, wh = labels[..., :2] * 0.6, labels[..., 2:4] * 0.35 + 0.02
    labels = torch.cat([xy, xy + wh, (labels[..., 4:] * classes).floor()], -1).contiguous()
    target = TR.build_targets(labels, None, size, anchors)
    assert int((target[:, :, 0] > 0).sum()) > 300
    raw = [torch.randn(batch, g * g, ld, device=G.DEV) * 1.5 for g in grids]
    losses, grads = TR.train_loss(raw, target, size, classes, anchors)
    total = torch.zeros(4, dtype=torch.float64)
    for b0 in range(0, batch, 8):
        l8, g8 = TR.train_loss([r[b0:b0 + 8].contiguous() for r in raw], target[b0:b0 + 8].contiguous(), size, classes,
                               anchors)
        total += l8.cpu().double() * 8
        for full, part in zip(grads, g8):
            assert torch.equal(full[b0:b0 + 8] * 4.0, part)
    np.testing.assert_allclose(losses.cpu().double().numpy() * batch, total.numpy(), rtol=1e-6)
    assert torch.isfinite(losses).all() and all(torch.isfinite(g).all() for g in grads)


@pytest.mark.parametrize("b,hw,k,n", [(2, 26, 96, 96), (3, 13, 96, 96), (1, 52, 96, 96), (2, 7, 8, 4)])
def test_conv3x3_backward_weight(G, TR, b, hw, k, n):
    """models/yolo_nano.py:44-47 `smooth_*` = Conv(96, 96, k=3, p=1): weight / bias gradient of the dense 3x3 conv
    (nine tap-shifted tcgen05 weight gradients, zero padding by predication) vs autograd."""
    torch.manual_seed(b + hw + k)
    x = torch.randn(b, k, hw, hw)
    w = (torch.randn(n, k, 3, 3) / (3 * k ** 0.5)).requires_grad_(True)
    bias = torch.randn(n, requires_grad=True)
    y = F.conv2d(x, w, bias, 1, 1)
    dy = torch.randn_like(y)
    y.backward(dy)
    nhwc = lambda t_: t_.detach().permute(0, 2, 3, 1).contiguous().to(G.DEV)  # noqa: E731
    dw, db = TR.conv3x3_backward_weight(nhwc(dy), nhwc(x))
    got = dw.cpu().view(3, 3, n, k).permute(2, 3, 0, 1)
    torch.testing.assert_close(got, w.grad, rtol=1e-4, atol=2e-5 * float(w.grad.abs().max()))
    torch.testing.assert_close(db.cpu(), bias.grad, rtol=1e-4, atol=2e-5 * float(bias.grad.abs().max()))


@pytest.mark.parametrize("first_size", [96, 192])
def test_training_branch_follows_set_grid(G, g7, first_size):
    """Multi-scale training calls model.set_grid(s) between steps (train.py:205,264): the loss kernel must take its
    level geometry from the CURRENT grid, not from the size the engine was created with (advisor finding) —
    created smaller and larger than the fixture's size, then set_grid, then the fixture's losses / gradients."""
    import contextlib, io
    import yolo_nano_b200 as pkg
    size, classes, seed = int(g7["size"]), int(g7["classes"]), int(g7["seed"])
    sd = W.calibrated(classes, seed=seed)
    x = W.synthetic_input(2, size, seed=seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = pkg.YOLONano(G.DEV, first_size, classes, anchor_size=W.anchors_for(classes))
    m.load_state_dict(sd)
    m = m.to(G.DEV).eval()
    m(W.synthetic_input(1, first_size, 1).to(G.DEV))            # engine + workspace exist at the first size
    m.set_grid(size)
    m.trainable = True
    ls = m(x.to(G.DEV), target=torch.from_numpy(g7["target"]).to(G.DEV))
    np.testing.assert_allclose(np.array([float(v) for v in ls], dtype=np.float32), g7["losses"], rtol=2e-3)
    ch = 3 * (1 + classes + 4)
    for k, g in zip(("pred_s", "pred_m", "pred_l"), m.head_gradients):
        want = g7["grad_" + k]
        b, _, h, w = want.shape
        gg = g.cpu()[:, :, :ch].reshape(b, h, w, ch).permute(0, 3, 1, 2).numpy()
        assert np.abs(gg - want).max() <= 3e-3 * np.abs(want).max(), k


def test_model_ema_one_launch_bit_exact(G, golden):
    """yolo_nano_b200.ModelEMA (one multi-tensor kernel launch per update) against the REAL reference class over three
    updates (fixture g8: parameters, BatchNorm statistics, vector tails, an integer counter) — bit-identical."""
    from yolo_nano_b200.ema import ModelEMA
    from oracle import train_oracle as T
    g8 = golden("g8_ema.npz")
    m = T.ema_model(int(g8["seed"])).to(G.DEV)
    ema = ModelEMA(m, decay=0.9999, updates=int(g8["start_updates"]))
    assert not ema.ema.training and all(not p.requires_grad for p in ema.ema.parameters())
    g = torch.Generator().manual_seed(int(g8["seed"]) + 1)
    for step in range(3):
        T.ema_perturb(m, g)
        ema.update(m)
        assert ema.decay(ema.updates) == float(g8[f"decay{step}"])
    for k, v in ema.ema.state_dict().items():
        np.testing.assert_array_equal(v.cpu().numpy(), g8["ema." + k], err_msg=k)


def test_model_ema_of_the_detector_feeds_its_engine(G):
    """train.py:147,233-235 + eval of `ema.ema`: the EMA copy of the drop-in detector is a working detector whose engine
    picks up the averaged weights (the kernel writes them behind autograd's version counters)."""
    import contextlib, io
    import yolo_nano_b200 as pkg
    from yolo_nano_b200.ema import ModelEMA
    sd = W.calibrated(20, seed=3)
    with contextlib.redirect_stdout(io.StringIO()):
        m = pkg.YOLONano(G.DEV, 128, 20, anchor_size=W.anchors_for(20))
    m.load_state_dict(sd)
    m = m.to(G.DEV).eval()
    x = W.synthetic_input(1, 128, 3).to(G.DEV)
    ema = ModelEMA(m, updates=100000)                       # decay ~ 0.9999
    before = ema.ema(x)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(0.5)
    for _ in range(3):
        ema.update(m)
    d = 1.0
    for u in (100001, 100002, 100003):
        d *= ema.decay(u)
    want = {k: (v * (d + (1 - d) * 0.5) if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd.items()}
    for k, v in ema.ema.state_dict().items():
        if v.dtype.is_floating_point:
            torch.testing.assert_close(v.cpu(), want[k], rtol=1e-5, atol=1e-7, msg=k)
    after = ema.ema(x)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = pkg.YOLONano(G.DEV, 128, 20, anchor_size=W.anchors_for(20))
    ref.load_state_dict(ema.ema.state_dict())
    ref = ref.to(G.DEV).eval()
    want_out = ref(x)
    for a, b in zip(after, want_out):
        np.testing.assert_array_equal(a, b)
    assert len(before[1]) > 0


# ---- round 2: the pieces of the chained training step ---------------------------------------------------------------
def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("batch,size", [(2, 64), (3, 96)])
def test_stem_conv_forward_and_weight_gradient(G, batch, size):
    """Unfused stem conv (backbone/shufflenetv2.py:109-113) and its weight gradient against autograd."""
    import torch.nn.functional as F
    lib = G._lib.load()
    g = torch.Generator().manual_seed(size)
    x = torch.randn(batch, 3, size, size, generator=g)
    w = (torch.randn(24, 3, 3, 3, generator=g) / 3).requires_grad_(True)
    y = F.conv2d(x, w, None, 2, 1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    xd = x.to(G.DEV)
    wp = w.detach().permute(1, 2, 3, 0).reshape(27, 24).contiguous().to(G.DEV)
    out = torch.empty(batch, size // 2, size // 2, 24, device=G.DEV)
    assert lib.ynb_stem_conv_fwd(G.ptr(xd), G.ptr(wp), G.ptr(out), batch, size, G.stream()) == 0
    torch.testing.assert_close(out.cpu(), _nhwc(y.detach()), rtol=1e-5, atol=1e-5)
    wsb = lib.ynb_stem_conv_bwd_weight_workspace_bytes(batch, size)
    ws = torch.empty(wsb, device=G.DEV, dtype=torch.uint8)
    dw = torch.empty(27, 24, device=G.DEV)
    dyd = _nhwc(dy).to(G.DEV)
    assert lib.ynb_stem_conv_bwd_weight(G.ptr(dyd), G.ptr(xd), G.ptr(dw), batch, size, G.ptr(ws), wsb, G.stream()) == 0
    want = w.grad.permute(1, 2, 3, 0).reshape(27, 24)
    torch.testing.assert_close(dw.cpu(), want, rtol=1e-4, atol=1e-4 * float(want.abs().max()))


@pytest.mark.parametrize("batch,h,w,c", [(2, 32, 32, 24), (1, 17, 23, 8)])
def test_maxpool_forward_backward(G, batch, h, w, c):
    """MaxPool2d(3, 2, 1) (backbone/shufflenetv2.py:116): forward exact, backward = ATen's (first maximum takes the
    gradient; ties included: the input is quantised so that windows hold equal values)."""
    import torch.nn.functional as F
    lib = G._lib.load()
    g = torch.Generator().manual_seed(h)
    x = (torch.randn(batch, c, h, w, generator=g) * 2).round() / 2          # many exact ties
    x.requires_grad_(True)
    y = F.max_pool2d(x, 3, 2, 1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    xd = _nhwc(x.detach()).to(G.DEV)
    out = torch.empty(batch, y.shape[2], y.shape[3], c, device=G.DEV)
    assert lib.ynb_maxpool3x3s2_fwd(G.ptr(xd), G.ptr(out), batch, h, w, c, G.stream()) == 0
    assert torch.equal(out.cpu(), _nhwc(y.detach()))
    din = torch.empty_like(xd)
    dyd = _nhwc(dy).to(G.DEV)
    assert lib.ynb_maxpool3x3s2_bwd(G.ptr(dyd), G.ptr(xd), G.ptr(din), batch, h, w, c, G.stream()) == 0
    torch.testing.assert_close(din.cpu(), _nhwc(x.grad), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("mode", [1, 2])
def test_resample_add_forward_backward(G, mode):
    """p + F.interpolate(q) (models/yolo_nano.py:291-296) and the gradient w.r.t. q, against autograd."""
    import torch.nn.functional as F
    lib = G._lib.load()
    g = torch.Generator().manual_seed(mode)
    b, h, w, c = 2, 12, 12, 96
    a = torch.randn(b, c, h, w, generator=g)
    a2 = torch.randn(b, c, h // 2 if mode == 1 else h * 2, w // 2 if mode == 1 else w * 2, generator=g, requires_grad=True)
    y = a + F.interpolate(a2, scale_factor=2.0 if mode == 1 else 0.5)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    ad, a2d = _nhwc(a).to(G.DEV), _nhwc(a2.detach()).to(G.DEV)
    out = torch.empty(b, h, w, c, device=G.DEV)
    assert lib.ynb_resample_add(G.ptr(ad), G.ptr(a2d), G.ptr(out), b, h, w, c, mode, G.stream()) == 0
    torch.testing.assert_close(out.cpu(), _nhwc(y.detach()), rtol=0, atol=0)
    da2 = torch.empty_like(a2d)
    dyd = _nhwc(dy).to(G.DEV)
    assert lib.ynb_resample_bwd(G.ptr(dyd), G.ptr(da2), b, h, w, c, mode, G.stream()) == 0
    torch.testing.assert_close(da2.cpu(), _nhwc(a2.grad), rtol=1e-6, atol=1e-6)
    s = torch.empty_like(ad)
    assert lib.ynb_add(G.ptr(ad), G.ptr(out), G.ptr(s), ad.numel(), G.stream()) == 0
    assert torch.equal(s.cpu(), (ad + out).cpu())


@pytest.mark.parametrize("b,hw", [(2, 13), (1, 26)])
def test_conv3x3_dense_forward_and_input_gradient(G, b, hw):
    """The standalone dense 3x3 entry (ynb_conv3x3_tc, 3xTF32): forward of a `smooth` conv and — with
    w'[k][8 - t][n] = w[n][t][k] — its input gradient, against autograd."""
    import torch.nn.functional as F
    lib = G._lib.load()
    g = torch.Generator().manual_seed(hw)
    c = 96
    x = torch.randn(b, c, hw, hw, generator=g, requires_grad=True)
    w = (torch.randn(c, c, 3, 3, generator=g) / (9 * c) ** 0.5)
    bias = torch.randn(c, generator=g) * 0.1
    y = F.conv2d(x, w, bias, 1, 1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    xd = _nhwc(x.detach()).to(G.DEV)
    wt = w.permute(0, 2, 3, 1).reshape(c, 9, c).contiguous().to(G.DEV)                  # [n][t][k]
    out = torch.empty(b, hw, hw, c, device=G.DEV)
    bd = bias.to(G.DEV)
    assert lib.ynb_conv3x3_tc(G.ptr(xd), c, G.ptr(out), c, G.ptr(wt), G.ptr(bd), b, hw, hw, c, c, 0, 1, G.stream()) == 0, \
        lib.ynb_last_error(None)
    torch.testing.assert_close(out.cpu(), _nhwc(y.detach()), rtol=1e-4, atol=1e-4)
    wd = w.permute(1, 2, 3, 0).reshape(c, 9, c).flip(1).contiguous().to(G.DEV)          # [k][8 - t][n]
    dx = torch.empty(b, hw, hw, c, device=G.DEV)
    dyd = _nhwc(dy).to(G.DEV)
    zero = torch.zeros(c, device=G.DEV)
    assert lib.ynb_conv3x3_tc(G.ptr(dyd), c, G.ptr(dx), c, G.ptr(wd), G.ptr(zero), b, hw, hw, c, c, 0, 1, G.stream()) == 0
    torch.testing.assert_close(dx.cpu(), _nhwc(x.grad), rtol=1e-4, atol=1e-4)


def test_chained_training_step_matches_reference_train_mode(G, golden):
    """BASELINE config 5's step, chained: TrainStep.forward_backward = model.train()(x, target) + total.backward() of
    the reference (BatchNorm on batch statistics, 77 convs forward and backward, every kernel from the library) against
    the train-mode oracle, itself pinned to the REAL reference by golden g10.

    What can be asserted: ReLU / LeakyReLU / max-pool make the gradient piecewise linear in the activations, and the
    reference's own forward amplifies a 1-ulp input change to ~1e-4 at the deep layers, so a few units change side and
    the REFERENCE's gradients move by per cent (g10 `grad_sens`: the largest relative change of each gradient over four
    1e-6 input perturbations, median 4e-2).  Hence three checks:
      1. the four losses (rtol 2e-4) and the running statistics after the step (rtol 2e-3);
      2. every backward op of the chain, in place, against torch.autograd of the same op on the same inputs and incoming
         gradient: forward, d input and parameter gradients within 2e-5 of their max (float32 / 3xTF32 rounding) — no
         systematic error in any of the 154 conv / BatchNorm / merge ops (stem conv, max-pool: own unit tests);
      3. every one of the 247 parameter gradients against the oracle within 2 x the reference's own sensitivity + 2e-4
         of the largest gradient of its module (composition: residual sums, chunk / cat / shuffle, resampling, pooling)."""
    import contextlib, io
    import yolo_nano_b200 as pkg
    from train_step_checker import CheckedTrainStep
    from oracle import train_oracle as T
    g = golden("g10_trainstep128.npz")
    size, classes, seed, batch = int(g["size"]), int(g["classes"]), int(g["seed"]), int(g["batch"])
    sd = W.calibrated(classes, seed=seed)
    x = W.synthetic_input(batch, size, seed=seed)
    target = torch.from_numpy(g["target"])
    want_l, want_g, want_state = T.train_forward_backward(sd, x, target, size, classes, W.anchors_for(classes))
    np.testing.assert_allclose(np.array(want_l, dtype=np.float32), g["losses"], rtol=1e-5)      # oracle == real reference
    with contextlib.redirect_stdout(io.StringIO()):
        m = pkg.YOLONano(G.DEV, size, classes, anchor_size=W.anchors_for(classes))
    m.load_state_dict(sd)
    m = m.to(G.DEV)
    m.trainable = True
    m.train()
    step = CheckedTrainStep(m)
    losses, grads = step.forward_backward(x.to(G.DEV), target.to(G.DEV))
    # 1. losses
    np.testing.assert_allclose(losses.cpu().numpy(), g["losses"], rtol=2e-4)
    # 2. every op of the chain against autograd of the same op
    assert len(step.records) >= 150, len(step.records)
    worst_op = max(step.records, key=lambda r: max([r["fwd"], r["dx"]] + list(r["params"].values())))
    print("[report] chained training step: %d ops checked in place, worst %s" % (len(step.records), worst_op))
    for r in step.records:
        assert r["fwd"] < 2e-5 and r["dx"] < 2e-5 and all(v < 2e-5 for v in r["params"].values()), r
    # 3. all parameter gradients against the oracle, bounded by the reference's own rounding sensitivity
    assert set(grads) == set(want_g) and len(grads) == 247
    sens = dict(zip([str(n) for n in g["grad_names"]], g["grad_sens"]))
    module_scale = {}
    for k, wg in want_g.items():
        mod = k.rsplit(".", 1)[0]
        module_scale[mod] = max(module_scale.get(mod, 0.0), float(wg.abs().max()))
    bad, ratios = {}, []
    for k, wg in want_g.items():
        got = grads[k].cpu()
        assert tuple(got.shape) == tuple(wg.shape), k
        err = float((got - wg).abs().max())
        bound = 2.0 * sens[k] * float(wg.abs().max()) + 2e-4 * module_scale[k.rsplit(".", 1)[0]]
        ratios.append(err / max(float(wg.abs().max()), 1e-30))
        if not err <= bound:
            bad[k] = (err, bound)
    flat_got = step.flat_gradient(grads).cpu().double()
    flat_want = torch.cat([want_g[k].reshape(-1) for k, _ in m.named_parameters()]).double()
    cos = float(torch.dot(flat_got, flat_want) / (flat_got.norm() * flat_want.norm()))
    print("[report] chained training step: gradient error / max, median %.2e (reference's own sensitivity: median %.2e); "
          "cosine of the flat gradient %.6f" % (float(np.median(ratios)), float(np.median(g["grad_sens"])), cos))
    assert not bad, bad
    assert cos > 0.999, cos
    msd = m.state_dict()
    for k, v in want_state.items():
        if k.endswith("num_batches_tracked"):
            assert int(msd[k]) == int(v), k
        else:
            torch.testing.assert_close(msd[k].cpu(), v, rtol=2e-3, atol=2e-4, msg=k)
    assert flat_got.numel() == sum(p.numel() for p in m.parameters())


def test_trainer_iterations_are_sgd_on_the_chained_gradients(G, golden, monkeypatch):
    """`Trainer.step` = train.py:219-235's loop body: three iterations (momentum buffer live from the second; the first
    runs eagerly, the second is captured into a CUDA graph and replayed, the third replays it) against the
    chained gradients of an identical second model + the CPU SGD oracle (`torch.optim.SGD` restated, train_oracle.py)
    — parameters equal after each iteration, state_dict layout unchanged, BatchNorm counters advanced, and the
    inference path of the trained model sees the new weights."""
    import contextlib, io
    import yolo_nano_b200 as pkg
    from yolo_nano_b200.train_step import TrainStep, Trainer
    from oracle import train_oracle as T
    monkeypatch.delenv("YNB_SYNC_CHECK", raising=False)      # a host read of a kernel flag cannot be captured into a graph
    g = golden("g10_trainstep128.npz")
    size, classes, seed, batch = int(g["size"]), int(g["classes"]), int(g["seed"]), int(g["batch"])
    sd = W.calibrated(classes, seed=seed)
    target = torch.from_numpy(g["target"]).to(G.DEV)
    models = []
    for _ in range(2):
        with contextlib.redirect_stdout(io.StringIO()):
            m = pkg.YOLONano(G.DEV, size, classes, anchor_size=W.anchors_for(classes))
        m.load_state_dict(sd)
        m = m.to(G.DEV)
        m.trainable = True
        m.train()
        models.append(m)
    a, b = models
    lr = 1e-5
    ema = pkg.ModelEMA(a)
    trainer = Trainer(a, lr=lr, ema=ema, cuda_graph=True)      # iteration 0 eager, 1 captured + replayed, 2 replayed
    assert list(a.state_dict().keys()) == list(sd.keys())
    fb = TrainStep(b)
    buf = None
    for it in range(3):
        x = W.synthetic_input(batch, size, seed=seed + it).to(G.DEV)
        la = trainer.step(x, target)
        lb, gb = fb.forward_backward(x, target)
        assert torch.equal(la, lb), (it, la, lb)
        flat_p = torch.cat([p.detach().reshape(-1) for p in b.parameters()]).cpu()
        new_p, buf = T.sgd_step(flat_p, fb.flat_gradient(gb).cpu(), buf, lr)
        off = 0
        for p in b.parameters():
            p.data.copy_(new_p[off:off + p.numel()].view(p.shape))
            off += p.numel()
        b.mark_weights_dirty()
        got = torch.cat([p.detach().reshape(-1) for p in a.parameters()]).cpu()
        exact = bool(torch.equal(got, new_p))
        print("[report] trainer iteration %d: parameters bit-identical to CPU SGD on the chained gradients: %s" % (it, exact))
        torch.testing.assert_close(got, new_p, rtol=1e-6, atol=1e-7)
        assert float((got - flat_p).abs().max()) > 0
    assert trainer.flat.data_ptr() == next(a.parameters()).data_ptr()
    assert int(a.state_dict()["backbone.conv1.1.num_batches_tracked"]) == 3 and ema.updates == 3
    assert int(a.state_dict()["backbone.stage2.1.branch2.4.num_batches_tracked"]) == 3
    # the trained weights reach the inference plan
    a.eval(); a.trainable = False
    with contextlib.redirect_stdout(io.StringIO()):
        fresh = pkg.YOLONano(G.DEV, size, classes, anchor_size=W.anchors_for(classes))
    fresh.load_state_dict({k: v.detach().cpu() for k, v in a.state_dict().items()})
    fresh = fresh.to(G.DEV).eval()
    xe = W.synthetic_input(2, size, seed=3).to(G.DEV)
    ra, rf = a.engine(2).forward_raw(xe), fresh.engine(2).forward_raw(xe)
    for u, v in zip(ra, rf):
        assert torch.equal(u, v)


@pytest.mark.parametrize("m,k,kw,n,nreal,trans", [(1000, 60, 58, 60, 58, False), (333, 116, 116, 464, 464, False),
                                                  (4096, 96, 96, 256, 255, False), (777, 60, 58, 116, 116, True),
                                                  (64, 256, 255, 96, 96, True)])
def test_async_pointwise_gemm_pads_and_transposes_on_the_device(G, m, k, kw, n, nreal, trans):
    """`ynb_pwconv_tc_async` (the training step's GEMM): weights packed by a kernel — channel padding (zero rows /
    columns) and the transposed read for the input gradient — against float64 matmul; 3xTF32 tolerance 2e-5 of max."""
    from yolo_nano_b200 import training as TR
    g = torch.Generator().manual_seed(m + n)
    x = torch.randn(m, k, generator=g)
    w = torch.randn(nreal, kw, generator=g) * 0.2          # logical W [out x in]
    b = torch.randn(nreal, generator=g)
    flags = []
    wd = (w.t().contiguous() if trans else w).to(G.DEV)
    y = TR.pwconv_forward(x.to(G.DEV), wd, b.to(G.DEV), act=0, transposed=trans, cout=n, flags=flags)
    torch.cuda.synchronize()
    assert not bool(torch.cat(flags).view(torch.int32).any())
    want = torch.zeros(m, n, dtype=torch.float64)
    want[:, :nreal] = x[:, :kw].double() @ w.double().t() + b.double()
    err = float((y.cpu().double() - want).abs().max() / want.abs().max())
    assert tuple(y.shape) == (m, n) and err < 2e-5, err
    assert float(y[:, nreal:].abs().max()) == 0.0 if n > nreal else True
