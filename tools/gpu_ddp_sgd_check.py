#!/usr/bin/env python
"""Config 5's data-parallel step tail on REAL GPUs: ONE NCCL all-reduce (sum) of the flat 1.33 M-element
gradient over NVLink + the fused SGD kernel with the 1 / world folded in — checked against the CPU oracle update on the
mean gradient, on every rank, and timed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/gpu_ddp_sgd_check.py [out.json]
"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import train_oracle as T  # noqa: E402  (the checker)
from yolo_nano_b200 import training as TR  # noqa: E402

N_PARAMS = 1326305       # YOLO-Nano-1.0x, COCO-80


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    p0 = torch.randn(N_PARAMS, generator=torch.Generator().manual_seed(7))
    grads = [[torch.randn(N_PARAMS, generator=torch.Generator().manual_seed(100 * step + r)) for r in range(world)]
             for step in range(2)]
    params = p0.to(dev)
    opt = TR.FlatSGD(params, lr=1e-3)
    for step in range(2):
        opt.step(grads[step][rank].to(dev))
    torch.cuda.synchronize()
    # CPU oracle: sum in rank order (exact for 2 ranks; NCCL's order may differ by an ulp for more), times 1 / world
    ref, buf = p0, None
    for step in range(2):
        gsum = grads[step][0].clone()
        for r in range(1, world):
            gsum += grads[step][r]
        ref, buf = T.sgd_step(ref, gsum * (1.0 / world), buf, 1e-3)
    got = params.cpu()
    exact = bool(torch.equal(got, ref))
    close = bool(torch.allclose(got, ref, rtol=2e-6, atol=1e-6))      # a few ulp: NCCL's summation order for > 2 ranks
    max_err = float((got - ref).abs().max())
    # every rank must hold identical parameters
    chk = torch.tensor([float(got.double().sum())], device=dev, dtype=torch.float64)
    allchk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allchk, chk)
    same = all(float(c) == float(allchk[0]) for c in allchk)
    # time the all-reduce + SGD tail
    g = torch.randn(N_PARAMS, device=dev)
    for _ in range(5):
        opt.step(g)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 50
    for _ in range(iters):
        opt.step(g)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = {"world": world, "params": N_PARAMS, "bit_exact_vs_cpu_oracle": exact, "within_few_ulp": close, "max_abs_err": max_err,
           "ranks_identical": same, "allreduce_plus_sgd_ms": float(t[0]),
           "allreduce_bus_gbs": 2 * (world - 1) / world * N_PARAMS * 4 / (float(t[0]) * 1e-3) / 1e9}
    ok = same and (exact if world == 2 else close)
    if rank == 0:
        print(json.dumps(res), flush=True)
        if len(sys.argv) > 1:
            Path(sys.argv[1]).write_text(json.dumps(res) + "\n")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
