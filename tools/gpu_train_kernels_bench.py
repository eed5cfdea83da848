"""Throughput of the training-branch kernels at BASELINE config 5's per-GPU shape
(416^2, COCO-80, 32 images): CUDA-event times over rotating buffers larger than L2, algorithmic bytes
/ time against MEASURED_PEAKS.json.  Run on the GPU box:  python tools/gpu_train_kernels_bench.py [out.json]"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import yolo_nano_b200 as pkg  # noqa: E402
from yolo_nano_b200 import training as TR  # noqa: E402

DEV = torch.device("cuda", 0)
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]


def timed(fn, sets, iters=20, warmup=3):
    """fn(i) uses buffer set i % sets; returns ms per call."""
    for i in range(warmup):
        fn(i % sets)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i % sets)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def row(name, ms, nbytes, note=""):
    gbs = nbytes / ms / 1e6
    return {"kernel": name, "ms": round(ms, 4), "algorithmic_mb": round(nbytes / 1e6, 2), "gbs": round(gbs, 1),
            "frac_of_hbm_peak": round(gbs / PEAK, 3), "note": note}


def main():
    out = []
    S, C, B, A = 416, 80, 32, 3
    anchors = pkg.MULTI_ANCHOR_SIZE_COCO
    grids = [S // s for s in (8, 16, 32)]
    N = sum(A * g * g for g in grids)
    ld = 256
    torch.manual_seed(0)
    # ---- targets: 32 labels per image
    labels = torch.rand(B, 32, 5, device=DEV)
    xy = labels[..., :2] * 0.6
    wh = labels[..., 2:4] * 0.35 + 0.02
    labels = torch.cat([xy, xy + wh, (labels[..., 4:] * C).floor()], -1).contiguous()
    ms = timed(lambda i: TR.build_targets(labels, None, S, anchors), 1)
    target = TR.build_targets(labels, None, S, anchors)
    out.append(row("build_targets (memset + 1 thread/image)", ms, B * N * 11 * 4, "latency-bound: 32 threads"))
    # ---- loss + head gradient: 4 rotating sets of raw maps (3 x 4 x 116 MB > L2)
    sets = 4
    raws = [[torch.randn(B, g * g, ld, device=DEV) for g in grids] for _ in range(sets)]
    lib = TR._lib.load()
    anc, _a = TR._anchors(anchors)
    grads = [torch.empty_like(r) for r in raws[0]]          # preallocated: the timed region is launches only
    losses = torch.empty(4, device=DEV)
    wsb = lib.ynb_train_loss_workspace_bytes(B, S)
    ws = torch.empty(wsb, device=DEV, dtype=torch.uint8)
    st = TR._stream_ptr(DEV)
    ms = timed(lambda i: lib.ynb_train_loss(TR._ptr(raws[i][0]), TR._ptr(raws[i][1]), TR._ptr(raws[i][2]), ld,
                                            TR._ptr(target), B, S, anc, A, C, TR._ptr(losses), TR._ptr(grads[0]),
                                            TR._ptr(grads[1]), TR._ptr(grads[2]), TR._ptr(ws), wsb, st), sets)
    cells = B * sum(g * g for g in grids)
    out.append(row("train_loss + finalize", ms, cells * (A * 5 * 4 + A * 11 * 4 + ld * 4),
                   "reads obj/box logits + targets, class logits of positives only; writes the gradient map once"))
    del raws, grads
    # ---- SGD: the model (1.326 M parameters) and a bandwidth-sized vector
    for n, tag in ((1326305, "model-sized: launch-latency bound"), (64 * 1024 * 1024, "bandwidth-sized")):
        ps = [torch.randn(n, device=DEV) for _ in range(2)]
        gs = [torch.randn(n, device=DEV) for _ in range(2)]
        opts = [TR.FlatSGD(p, lr=1e-3) for p in ps]
        for o, g in zip(opts, gs):
            o.step(g)
        ms = timed(lambda i: opts[i].step(gs[i]), 2)
        out.append(row(f"sgd_step n={n}", ms, n * 4 * 5, tag))
        del ps, gs, opts
    # ---- conv backward at stage-2 / head shapes
    for (b, c, hw, s) in ((B, 60, 52, 1), (B, 60, 104, 2), (B, 96, 52, 1)):
        sets = 3
        xs = [torch.randn(b, hw, hw, c, device=DEV) for _ in range(sets)]
        ho = (hw - 1) // s + 1
        dys = [torch.randn(b, ho, ho, c, device=DEV) for _ in range(sets)]
        w = torch.randn(9, c, device=DEV)
        dx = torch.empty_like(xs[0])
        ms = timed(lambda i: lib.ynb_dwconv3x3_bwd_data(TR._ptr(dys[i]), c, 0, TR._ptr(dx), c, 0, TR._ptr(w), b, hw, hw,
                                                        c, s, st), sets)
        out.append(row(f"dwconv3x3_bwd_data B={b} C={c} {hw}^2 s{s}", ms, 4 * c * b * (hw * hw + ho * ho)))
        wsb = lib.ynb_dwconv3x3_bwd_weight_workspace_bytes(b, hw, hw, c, s)
        ws = torch.empty(wsb, device=DEV, dtype=torch.uint8)
        dwdb = torch.empty(10, c, device=DEV)
        ms = timed(lambda i: lib.ynb_dwconv3x3_bwd_weight(TR._ptr(dys[i]), c, 0, TR._ptr(xs[i]), c, 0, TR._ptr(dwdb), b,
                                                          hw, hw, c, s, TR._ptr(ws), wsb, st), sets)
        out.append(row(f"dwconv3x3_bwd_weight B={b} C={c} {hw}^2 s{s}", ms, 4 * c * b * (hw * hw + ho * ho)))
        del xs, dys
    for (m, k, n) in ((B * 2704, 116, 116), (B * 2704, 96, 96), (B * 676, 232, 232), (B * 2704, 96, 256)):
        sets = 3
        xs = [torch.randn(m, k, device=DEV) for _ in range(sets)]
        dys = [torch.randn(m, n, device=DEV) for _ in range(sets)]
        wsb = lib.ynb_pwconv_bwd_weight_workspace_bytes(m, k, n)
        ws = torch.empty(wsb, device=DEV, dtype=torch.uint8)
        dw, db = torch.empty(n, k, device=DEV), torch.empty(n, device=DEV)
        ms = timed(lambda i: lib.ynb_pwconv_bwd_weight(TR._ptr(dys[i]), n, 0, TR._ptr(xs[i]), k, 0, TR._ptr(dw), TR._ptr(db),
                                                       m, k, n, TR._ptr(ws), wsb, st), sets)
        r = row(f"pwconv_bwd_weight M={m} K={k} N={n}", ms, 4 * m * (k + n))
        r["tflops"] = round(2 * m * k * n / ms / 1e9, 1)
        out.append(r)
        del xs, dys
    # ---- BatchNorm (training mode) + activation, stage-2 / neck shapes
    for (m, c, act) in ((B * 2704, 60, 1), (B * 2704, 96, 2), (B * 676, 232, 1)):
        sets = 3
        xs = [torch.randn(m, c, device=DEV) for _ in range(sets)]
        dys = [torch.randn(m, c, device=DEV) for _ in range(sets)]
        gamma, beta = torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV)
        rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
        y, dx = torch.empty_like(xs[0]), torch.empty_like(xs[0])
        mean, rstd, dgb = torch.empty(c, device=DEV), torch.empty(c, device=DEV), torch.empty(2 * c, device=DEV)
        wsb = lib.ynb_bn_workspace_bytes(m, c)
        ws = torch.empty(wsb, device=DEV, dtype=torch.uint8)
        P = TR._ptr
        ms = timed(lambda i: lib.ynb_bn_train_fwd(P(xs[i]), c, 0, P(y), c, 0, P(gamma), P(beta), P(rm), P(rv), P(mean),
                                                  P(rstd), m, c, 1e-5, 0.1, act, P(ws), wsb, st), sets)
        out.append(row(f"bn_train_fwd M={m} C={c} act={act}", ms, 4 * m * c * 2,
                       "x is read twice (statistics, apply): 3 passes of traffic for 2 algorithmic"))
        ms = timed(lambda i: lib.ynb_bn_train_bwd(P(dys[i]), c, 0, P(xs[i]), c, 0, P(y), c, 0, P(gamma), P(mean), P(rstd),
                                                  P(dx), c, 0, P(dgb), m, c, act, P(ws), wsb, st), sets)
        out.append(row(f"bn_train_bwd M={m} C={c} act={act}", ms, 4 * m * c * 4,
                       "reads dy, x, y, writes dx; the reduction pass re-reads dy, x, y"))
        del xs, dys
    for r in out:
        print(json.dumps(r))
    if len(sys.argv) > 1:
        Path(sys.argv[1]).write_text("\n".join(json.dumps(r) for r in out) + "\n")


if __name__ == "__main__":
    main()
