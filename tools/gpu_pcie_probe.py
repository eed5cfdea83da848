"""Host<->device copy bandwidth of the box (pinned memory): one stream, N concurrent streams, chunked."""
import torch

n = 64 * 3 * 416 * 416
x = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: d.copy_(x, non_blocking=True))
print(f"H2D 1 stream : {n * 4 / 1e6:.0f} MB in {ms:.3f} ms = {n * 4 / ms / 1e6:.1f} GB/s")
ms = timed(lambda: x.copy_(d, non_blocking=True))
print(f"D2H 1 stream : {n * 4 / ms / 1e6:.1f} GB/s")
for k in (2, 4):
    streams = [torch.cuda.Stream() for _ in range(k)]
    parts = [(i * n // k, (i + 1) * n // k) for i in range(k)]
    cur = torch.cuda.current_stream()

    def fn():
        ev = torch.cuda.Event()
        ev.record(cur)
        for s, (a, b) in zip(streams, parts):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                d[a:b].copy_(x[a:b], non_blocking=True)
        for s in streams:
            cur.wait_stream(s)
    ms = timed(fn)
    print(f"H2D {k} streams: {n * 4 / ms / 1e6:.1f} GB/s")
# host-side page placement: first-touch by this thread vs interleaved is not controllable here; report NUMA
import os
print("cpus", len(os.sched_getaffinity(0)))
