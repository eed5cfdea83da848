"""Accuracy of the raw head outputs against the fp32 oracle AND an fp64 evaluation of the same network:
how much of the distance to the oracle is fp32 re-association noise (which the oracle has too)?
  SWEEP_MODE=ffma|3xtf32|tf32 python tools/gpu_accuracy_sweep.py"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import gpu_util as G  # noqa: E402
from oracle import weights as W, yolo_nano_oracle as O  # noqa: E402

torch.set_num_threads(16)
MODE = os.environ.get("SWEEP_MODE", "3xtf32")


def worst(a, b):
    w, v = 0.0, 0
    for r, g in zip(a, b):
        err = np.abs(g - r)
        tol = 1e-3 + 1e-4 * np.abs(r)
        w = max(w, float((err / tol).max()))
        v += int((err > tol).sum())
    return w, v


for size, classes in ((320, 20), (416, 80)):
    for seed in (11, 12, 14):
        sd = W.calibrated(classes, seed=seed)
        x = W.synthetic_input(1, size, seed)
        ref32 = [t.numpy() for t in O.network(sd, x)]
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        ref64 = [t.numpy() for t in O.network(sd64, x.double())]
        eng = G.make_engine(sd, size, classes, MODE)
        got = [t.cpu().numpy() for t in eng.forward_raw(x.to(G.DEV))]
        eng.close()
        a = worst(ref32, got)
        b = worst(ref64, got)
        c = worst(ref64, ref32)
        print(f"{MODE} S={size} C={classes} seed={seed}: vs fp32 oracle {a[0]:.2f} tol ({a[1]} over) | "
              f"vs fp64 truth: ours {b[0]:.2f} tol ({b[1]} over), fp32 oracle {c[0]:.2f} tol ({c[1]} over)")
