"""Pins oracle/train_oracle.py (training branch at the head boundary) against
g7_train128.npz recorded from the REAL reference (oracle/gen_golden.py train).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import train_oracle as T
from oracle import weights as W


@pytest.fixture(scope="module")
def g7(golden):
    return golden("g7_train128.npz")


def _labels(g7):
    return [[list(r) for r in g7["labels"][b][: int(g7["n_labels"][b])]] for b in range(len(g7["n_labels"]))]


def test_target_assignment_matches_reference(g7):
    """tools.multi_gt_creator: positives, ignored (-1) anchors, dirty boxes, overwrite order."""
    t = T.multi_gt_creator(int(g7["size"]), _labels(g7), W.anchors_for(int(g7["classes"])))
    want = g7["target"]
    assert (want[:, :, 0] < 0).sum() > 0 and (want[:, :, 0] > 0).sum() > 0
    np.testing.assert_array_equal(t.numpy(), want)


def test_losses_and_head_gradients_match_reference(g7):
    preds = [torch.from_numpy(g7[k]) for k in ("pred_s", "pred_m", "pred_l")]
    ls, grads = T.losses_and_grads(preds, torch.from_numpy(g7["target"]), int(g7["size"]), int(g7["classes"]),
                                   W.anchors_for(int(g7["classes"])))
    np.testing.assert_allclose(np.array(ls, dtype=np.float32), g7["losses"], rtol=1e-6)
    for k, g in zip(("pred_s", "pred_m", "pred_l"), grads):
        np.testing.assert_allclose(g.numpy(), g7["grad_" + k], rtol=1e-5, atol=1e-8, err_msg=k)


def test_sgd_step_matches_torch_optim(g7):
    p0, g1, g2 = (torch.from_numpy(g7[k]) for k in ("sgd_p0", "sgd_g1", "sgd_g2"))
    p1, buf = T.sgd_step(p0, g1, None, 1e-3)
    np.testing.assert_array_equal(p1.numpy(), g7["sgd_p1"])
    p2, _ = T.sgd_step(p1, g2, buf, 1e-3)
    np.testing.assert_array_equal(p2.numpy(), g7["sgd_p2"])


def test_ema_update_matches_reference_class(golden):
    """oracle ema_update / ema_decay against the REAL ModelEMA (utils/misc.py:67-86) over three updates (g8)."""
    from copy import deepcopy
    g8 = golden("g8_ema.npz")
    m = T.ema_model(int(g8["seed"]))
    ema_sd = deepcopy(m).eval().state_dict()
    g = torch.Generator().manual_seed(int(g8["seed"]) + 1)
    updates = int(g8["start_updates"])
    for step in range(3):
        T.ema_perturb(m, g)
        updates += 1
        d = T.ema_decay(updates)
        assert d == float(g8[f"decay{step}"])
        T.ema_update(ema_sd, m.state_dict(), d)
    for k, v in ema_sd.items():
        np.testing.assert_array_equal(v.numpy(), g8["ema." + k], err_msg=k)
    assert int(ema_sd["1.num_batches_tracked"]) == 0            # integer state is not averaged


def test_train_step_oracle_matches_reference_in_train_mode(golden):
    """oracle train_forward_backward (BatchNorm on batch statistics, the four losses, total.backward()) against the
    REAL reference in train() mode (golden g10): losses, every parameter gradient by (sum, abs-sum, max), eleven
    gradients in full, running statistics after the step."""
    g = golden("g10_trainstep128.npz")
    size, classes, seed, batch = int(g["size"]), int(g["classes"]), int(g["seed"]), int(g["batch"])
    sd = W.calibrated(classes, seed=seed)
    x = W.synthetic_input(batch, size, seed=seed)
    assert W.digest(sd) == str(g["sd_digest"]) and W.digest(x) == str(g["x_digest"])
    ls, grads, state = T.train_forward_backward(sd, x, torch.from_numpy(g["target"]), size, classes, W.anchors_for(classes))
    np.testing.assert_allclose(np.array(ls, dtype=np.float32), g["losses"], rtol=1e-5)
    names = [str(n) for n in g["grad_names"]]
    assert len(names) == 247 and set(names) == set(grads)
    for n, (s, a, mx) in zip(names, g["grad_stats"]):
        gr = grads[n]
        assert abs(float(gr.double().abs().sum()) - a) <= 2e-4 * a + 1e-9, n
        assert abs(float(gr.abs().max()) - mx) <= 2e-4 * mx + 1e-9, n
    for k in g.files:
        if k.startswith("grad."):
            want = g[k]
            np.testing.assert_allclose(grads[k[5:]].numpy(), want, rtol=2e-3, atol=2e-5 * float(np.abs(want).max()), err_msg=k)
        elif k.startswith("rm."):
            np.testing.assert_allclose(state[k[3:] + ".running_mean"].numpy(), want := g[k], rtol=1e-5, atol=1e-6)
        elif k.startswith("rv."):
            np.testing.assert_allclose(state[k[3:] + ".running_var"].numpy(), g[k], rtol=1e-5, atol=1e-6)
        elif k.startswith("nbt."):
            assert int(state[k[4:] + ".num_batches_tracked"]) == int(g[k])
