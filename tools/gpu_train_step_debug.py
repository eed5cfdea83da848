#!/usr/bin/env python
"""Debug aid for the chained training step: every backward closure of `TrainStep` is checked IN PLACE against
torch.autograd of the same op on the same inputs and incoming gradient (torch is the checker here, never the product),
then the parameter gradients are listed against the train-mode oracle from the last layer backwards.

    python tools/gpu_train_step_debug.py
"""
import contextlib
import io
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import train_oracle as T  # noqa: E402
from oracle import weights as W  # noqa: E402
import yolo_nano_b200 as pkg  # noqa: E402
from yolo_nano_b200.train_step import TrainStep  # noqa: E402
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from train_step_checker import CheckedTrainStep, rel  # noqa: E402

def main():
    g = np.load(Path(__file__).resolve().parent.parent / "tests/golden/g10_trainstep128.npz")
    size, classes, seed, batch = int(g["size"]), int(g["classes"]), int(g["seed"]), int(g["batch"])
    sd = W.calibrated(classes, seed=seed)
    x = W.synthetic_input(batch, size, seed=seed)
    target = torch.from_numpy(g["target"])
    want_l, want_g, _ = T.train_forward_backward(sd, x, target, size, classes, W.anchors_for(classes))
    sens = dict(zip([str(n) for n in g["grad_names"]], g["grad_sens"]))
    dev = torch.device("cuda", 0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = pkg.YOLONano(dev, size, classes, anchor_size=W.anchors_for(classes))
    m.load_state_dict(sd)
    m = m.to(dev)
    m.trainable = True
    m.train()
    step = CheckedTrainStep(m) if "--ops" in sys.argv else TrainStep(m)
    losses, grads = step.forward_backward(x.to(dev), target.to(dev))
    for r in getattr(step, "records", []):
        print("%-36s fwd %.1e  dx %.1e  %s" % (r["op"], r["fwd"], r["dx"],
                                              "  ".join("%s %.1e" % (k.split(".")[-1], v) for k, v in r["params"].items())))
    print("losses", losses.cpu().numpy(), want_l)
    for k, p in reversed(list(m.named_parameters())):
        wg = want_g[k]
        got = grads[k].cpu()
        print("%-44s max|ref| %.3e  err/max %.3e  (reference's own sensitivity %.1e)"
              % (k, float(wg.abs().max()), rel(got, wg), sens[k]))


if __name__ == "__main__":
    main()
