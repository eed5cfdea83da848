"""Host-side multi-GPU logic on CPU: shard arithmetic and the gloo gather of variable-length
detections with world_size 2 (the data path itself has no collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolo_nano_b200 import sharding as S


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 255, 1024):
        for world in (1, 2, 3, 4, 8):
            spans = [S.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        S.shard_range(4, 2, 2)


def _fake(i):
    rng = np.random.default_rng(i)
    k = int(rng.integers(0, 6)) if i % 3 else 0          # some images keep nothing
    return (rng.random((k, 4), dtype=np.float32), rng.random(k, dtype=np.float32),
            rng.integers(0, 80, k).astype(np.int64))


def test_pack_roundtrip():
    dets = [_fake(i) for i in range(9)]
    back = S.unpack_detections(*S.pack_detections(dets))
    for a, b in zip(dets, back):
        for u, v in zip(a, b):
            np.testing.assert_array_equal(u, v)
    assert S.unpack_detections(*S.pack_detections([])) == []


def _worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = S.shard_range(n_images, world, rank)
    local = [_fake(i) for i in range(lo, hi)]
    full = S.gather_detections(local)
    ok = len(full) == n_images and all(
        np.array_equal(a, b) for i, d in enumerate(full) for a, b in zip(d, _fake(i)))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [7, 2, 1])
def test_gather_world2_gloo(n_images):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _grad_worker(rank, world, port, q):
    """Config 5: one all-reduce (sum) of the flat gradient; the 1/world is folded into
    the SGD kernel on the device, here into the oracle update."""
    from oracle import train_oracle as T
    from yolo_nano_b200 import training as TR
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    grad = torch.randn(1001, generator=g)
    TR.allreduce_gradients(grad)
    want = sum(torch.randn(1001, generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
    p0 = torch.randn(1001, generator=torch.Generator().manual_seed(7))
    p1, _ = T.sgd_step(p0, grad * (1.0 / world), None, 1e-3)
    ref, _ = T.sgd_step(p0, want * (1.0 / world), None, 1e-3)
    q.put((rank, bool(torch.equal(grad, want)) and bool(torch.equal(p1, ref))))
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
