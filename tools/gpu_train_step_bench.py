#!/usr/bin/env python
"""BASELINE config 5 as one bench line: the data-parallel TRAINING step of YOLO-Nano-1.0x (416x416, COCO-80,
32 images per GPU): multi_gt_creator on the device, forward in train() mode (BatchNorm on batch statistics), the four
losses, the backward chain through all 77 convolutions, ONE NCCL all-reduce of the flat gradient over NVLink and the
fused SGD(momentum 0.9, wd 5e-4) kernel — `yolo_nano_b200.train_step.Trainer.step`, every arithmetic step a kernel of
libyolonano_b200.so (or NCCL).  Random-init weights, synthetic images and labels.

    python tools/gpu_train_step_bench.py [--steps K --warmup W --batch 32 --size 416] [out.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/gpu_train_step_bench.py ...

`value` times the step with inputs resident in HBM; `e2e` adds, per step, the H2D copy of the images and labels from
pinned host memory and the D2H read of the four losses (what train.py:215-226 does).  CUDA events, max over ranks.
This path is correct (tests/test_gpu_training.py) but NOT tuned: see DESIGN.md §8 a14.
"""
import argparse
import contextlib
import io
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import yolo_nano_b200 as pkg  # noqa: E402
from yolo_nano_b200 import training as TR  # noqa: E402
from yolo_nano_b200.train_step import Trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--no-graph", action="store_true", help="launch the ~1600 kernels of a step one by one")
    ap.add_argument("--cpu-baseline", action="store_true",
                    help="also time the train-mode oracle (= the reference's PyTorch CPU path restated) on the host cores, "
                         "one 8-image step at the same size")
    ap.add_argument("--profile", action="store_true", help="host enqueue time + kernel table (torch.profiler / CUPTI)")
    ap.add_argument("out", nargs="?", default="")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    classes, anchors = 80, pkg.MULTI_ANCHOR_SIZE_COCO
    torch.manual_seed(0)                               # identical initial weights on every rank (DDP broadcasts them)
    with contextlib.redirect_stdout(io.StringIO()):
        m = pkg.YOLONano(dev, a.size, classes, anchor_size=anchors, trainable=True)
    m = m.to(dev).train()
    trainer = Trainer(m, lr=1e-3, skip_nan=False, cuda_graph=not a.no_graph)
    g = torch.Generator().manual_seed(100 + rank)
    x_h = torch.randn(a.batch, 3, a.size, a.size, generator=g).pin_memory()
    nlab = 8
    c = torch.rand(a.batch, nlab, 2, generator=g) * 0.6 + 0.2
    wh = torch.rand(a.batch, nlab, 2, generator=g) * 0.3 + 0.05
    cls = torch.randint(0, classes, (a.batch, nlab, 1), generator=g).float()
    lab_h = torch.cat((c - wh / 2, c + wh / 2, cls), dim=-1).clamp_(0, classes).pin_memory()
    lab_h[..., :4].clamp_(0.0, 1.0)
    x_d, lab_d = x_h.to(dev), lab_h.to(dev)
    xb, lb = torch.empty_like(x_d), torch.empty_like(lab_d)
    loss_h = torch.empty(4).pin_memory()

    def step_dev():
        return trainer.step(x_d, TR.build_targets(lab_d, None, a.size, anchors))

    def step_host():
        xb.copy_(x_h, non_blocking=True)
        lb.copy_(lab_h, non_blocking=True)
        ls = trainer.step(xb, TR.build_targets(lb, None, a.size, anchors))
        loss_h.copy_(ls, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # train.py reads the losses on the host (isnan check)
        return ls

    def timed(fn):
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            ls = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), ls

    if a.profile:
        import time
        from torch.profiler import ProfilerActivity, profile
        for _ in range(3):
            step_dev()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            step_dev()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("host enqueue %.1f ms/step, drained after %.1f ms more" % ((t1 - t0) / 5 * 1e3, (t2 - t1) * 1e3))
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step_dev()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
        return
    ms, ls = timed(step_dev)
    ms_e2e, _ = timed(step_host)
    res = {"metric": "images/sec YOLO-Nano-1.0x training step %d bs%d/GPU (train-mode BN, backward, all-reduce, SGD)" % (a.size, a.batch),
           "value": world * a.batch / (ms * 1e-3), "unit": "images/sec", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (3xTF32 GEMMs)",
           "data": "synthetic images + labels, random-init weights",
           "config": {"workload": "BASELINE config 5: 416x416 COCO-80 training step, %d images/GPU, data-parallel" % a.batch,
                      "parameters": int(trainer.flat.numel()), "optimizer": "SGD momentum 0.9 wd 5e-4, one flat all-reduce"},
           "e2e": {"value": world * a.batch / (ms_e2e * 1e-3), "unit": "images/sec",
                   "h2d_bytes_per_step": int(x_h.numel() * 4 + lab_h.numel() * 4), "d2h_bytes_per_step": 16},
           "losses_last_step": [float(v) for v in ls.cpu()], "cuda_graph": not a.no_graph}
    if a.cpu_baseline and rank == 0:
        import time
        from oracle import train_oracle as TO          # the checker, used here only as the reported CPU baseline
        sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
        nb = 8
        tgt = TR.build_targets(lab_d[:nb], None, a.size, anchors).cpu()
        TO.train_forward_backward(sd, x_h[:2].clone(), tgt[:2], a.size, classes, anchors)       # warm-up
        t0 = time.perf_counter()
        TO.train_forward_backward(sd, x_h[:nb].clone(), tgt, a.size, classes, anchors)
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": nb / dt, "unit": "images/sec", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "one forward + backward of %d images at %d (train-mode oracle, PyTorch CPU, %d threads)"
                                         % (nb, a.size, torch.get_num_threads())}
    if rank == 0:
        print(json.dumps(res), flush=True)
        if a.out:
            Path(a.out).write_text(json.dumps(res) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
