"""Host<->device copy bandwidth of the box (pinned memory), alone and while a forward runs."""
import time
import torch

x = torch.empty(64 * 3 * 416 * 416, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
s = torch.cuda.Stream()
for name, src, dst in (("H2D", x, d), ("D2H", d, x)):
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {x.numel() * 4 / 1e6:.0f} MB in {ms:.3f} ms = {x.numel() * 4 / ms / 1e6:.1f} GB/s")
