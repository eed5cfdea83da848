"""First-contact diagnostics on a real B200: prints per-stage errors of every arithmetic
mode against the oracle instead of asserting, plus rough timings.  Not a test, not a bench.
    python tools/gpu_probe.py [--size 128] [--batch 2]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import weights as W  # noqa: E402
from oracle import yolo_nano_oracle as O  # noqa: E402
import gpu_util as G  # noqa: E402


def probe_pw(lib):
    print("== unit: pointwise conv, FFMA vs tcgen05 ==")
    torch.manual_seed(0)
    for (m, cin, cout) in [(1000, 24, 58), (4321, 60, 58), (2704 * 2, 116, 116), (700, 232, 232),
                           (338, 464, 96), (5408, 96, 255), (128, 32, 16), (129, 96, 96)]:
        x = torch.randn(m, cin, device=G.DEV)
        w = torch.randn(cout, cin, device=G.DEV) / cin ** 0.5
        b = torch.randn(cout, device=G.DEV)
        ref = (x.double() @ w.double().t() + b.double()).float()
        for name, fn, mode in (("ffma", lib.ynb_pwconv, None), ("3xtf32", lib.ynb_pwconv_tc, 1),
                               ("tf32", lib.ynb_pwconv_tc, 2)):
            out = torch.full((m, cout), float("nan"), device=G.DEV)
            args = [G.ptr(x), cin, 0, G.ptr(out), cout if cout % 4 == 0 else cout, 0, 1, G.ptr(w), G.ptr(b), m, cin, cout, 0]
            if cout % 4:   # tc path wants ld multiple of 4
                ld = (cout + 3) // 4 * 4
                out = torch.full((m, ld), float("nan"), device=G.DEV)
                args[3], args[4] = G.ptr(out), ld
            if mode is not None:
                args.append(mode)
            args.append(G.stream())
            rc = fn(*args)
            torch.cuda.synchronize()
            if rc != 0:
                print(f"  M={m} K={cin} N={cout} {name}: rc={rc} {lib.ynb_last_error(None)}")
                continue
            o = out[:, :cout]
            err = (o - ref).abs().max().item()
            print(f"  M={m:6d} K={cin:3d} N={cout:3d} {name:7s} max|err|={err:.3e} nan={int(torch.isnan(o).sum())}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--classes", type=int, default=80)
    ap.add_argument("--skip-unit", action="store_true")
    a = ap.parse_args()
    from yolo_nano_b200 import _lib
    lib = _lib.load()
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    if not a.skip_unit:
        probe_pw(lib)

    sd = W.calibrated(a.classes, seed=1)
    x = W.synthetic_input(a.batch, a.size, seed=1)
    taps = {}
    t0 = time.time()
    preds = O.network(sd, x, taps)
    bbox, cls = O.decode(preds, a.size, a.classes, W.anchors_for(a.classes))
    print(f"oracle network+decode: {time.time() - t0:.2f}s")
    xd = x.to(G.DEV)
    names = ["pool"] + [f"stage{s}.{i}" for s, n in ((2, 4), (3, 8), (4, 4)) for i in range(n)] + \
            ["lat3", "lat4", "lat5", "fpn4", "p3", "p4", "p5", "pred_s", "pred_m", "pred_l"]
    for mode in ("ffma", "3xtf32", "tf32"):
        print(f"== engine mode {mode} ==")
        try:
            eng = G.make_engine(sd, a.size, a.classes, mode, max_batch=a.batch)
            raw = eng.forward_raw(xd)
            torch.cuda.synchronize()
        except Exception as ex:  # noqa: BLE001
            print("  FAILED:", ex)
            continue
        for n in names:
            got = eng.read_tap(n, a.batch).cpu().numpy()
            print(f"  {n:10s} rel_err={G.rel_err(got, taps[n].numpy()):.3e} nan={int(np.isnan(got).sum())}")
        for i, k in enumerate(("pred_s", "pred_m", "pred_l")):
            d = np.abs(raw[i].cpu().numpy() - taps[k].numpy())
            tol = 1e-3 + 1e-4 * np.abs(taps[k].numpy())
            print(f"  raw {k}: max|err|={d.max():.3e}  violations of 1e-3+1e-4*|ref|: {int((d > tol).sum())}")
        boxes, scores, cl = eng.forward_decode(xd)
        torch.cuda.synchronize()
        sc_ref, ci_ref = zip(*[O.class_scores(cls[i].numpy()) for i in range(a.batch)])
        print(f"  decode: box max|err|={np.abs(boxes.cpu().numpy() - bbox.numpy()).max() * a.size:.3e} px, "
              f"score max|err|={np.abs(scores.cpu().numpy() - np.stack(sc_ref)).max():.3e}, "
              f"cls mismatches={int((cl.cpu().numpy() != np.stack(ci_ref)).sum())}")
        ob, os_, oc, on = eng.forward_detect(xd)
        torch.cuda.synchronize()
        bh, sh, ch = boxes.cpu().numpy(), scores.cpu().numpy(), cl.cpu().numpy()
        for i in range(a.batch):
            _, _, _, idx = O.postprocess_flat(bh[i], sh[i], ch[i].astype(np.int64), a.classes, 0.001, 0.5)
            k = int(on[i])
            same = k == len(idx) and np.array_equal(ob[i, :k].cpu().numpy(), bh[i][idx])
            print(f"  detect img{i}: kept {k} (oracle on same candidates {len(idx)}) identical={same}")
        # timing (rough)
        for _ in range(3):
            eng.forward_detect(xd)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(10):
            eng.forward_detect(xd)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / 10
        print(f"  forward_detect: {dt * 1e3:.3f} ms / batch of {a.batch} -> {a.batch / dt:.0f} img/s")
        eng.close()


if __name__ == "__main__":
    main()
