"""Batch sharding across GPUs: images are independent, so a batch is split
into contiguous ranges, one per rank (one process per GPU), every rank runs the whole
path on its range with replicated weights, and nothing crosses NVLink on the data path.
Only the (tiny, variable-length) detections are gathered to the caller — host side,
through `torch.distributed` (NCCL between GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

Detection = Tuple[np.ndarray, np.ndarray, np.ndarray]   # bboxes [K,4] f32, scores [K] f32, cls [K] i64


def shard_range(n_images: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced range of images for `rank` (first `n % world` ranks get one more)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world / rank")
    base, extra = divmod(n_images, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pack_detections(dets: Sequence[Detection]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Variable-length per-image results -> (counts [n] i64, rows [sum K, 6] f32: box, score, cls)."""
    counts = torch.tensor([len(d[1]) for d in dets], dtype=torch.int64)
    if int(counts.sum()) == 0:
        return counts, torch.zeros((0, 6), dtype=torch.float32)
    rows = [np.concatenate([d[0].reshape(-1, 4), d[1].reshape(-1, 1), d[2].reshape(-1, 1).astype(np.float32)], 1)
            for d in dets if len(d[1])]
    return counts, torch.from_numpy(np.concatenate(rows, 0).astype(np.float32))


def unpack_detections(counts: torch.Tensor, rows: torch.Tensor) -> List[Detection]:
    out, pos = [], 0
    for k in counts.tolist():
        r = rows[pos:pos + k].numpy()
        out.append((np.ascontiguousarray(r[:, :4], dtype=np.float32), np.ascontiguousarray(r[:, 4], dtype=np.float32),
                    r[:, 5].astype(np.int64)))
        pos += k
    return out


def gather_detections(local: Sequence[Detection], device: torch.device = torch.device("cpu")) -> List[Detection]:
    """All ranks call this with the detections of their shard (in shard order); every rank
    gets the detections of the whole batch in image order.  Two collectives: image/row
    counts, then rows padded to the largest shard."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local)
    world = dist.get_world_size()
    counts, rows = pack_detections(local)
    meta = torch.tensor([len(counts), rows.shape[0]], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    max_img = max(int(m[0]) for m in metas)
    max_rows = max(int(m[1]) for m in metas)
    cpad = torch.zeros(max_img, dtype=torch.int64, device=device)
    cpad[:len(counts)] = counts.to(device)
    rpad = torch.zeros((max_rows, 6), dtype=torch.float32, device=device)
    rpad[:rows.shape[0]] = rows.to(device)
    call = [torch.zeros_like(cpad) for _ in range(world)]
    rall = [torch.zeros_like(rpad) for _ in range(world)]
    dist.all_gather(call, cpad)
    dist.all_gather(rall, rpad)
    out: List[Detection] = []
    for r in range(world):
        n_img, n_rows = int(metas[r][0]), int(metas[r][1])
        out.extend(unpack_detections(call[r][:n_img].cpu(), rall[r][:n_rows].cpu()))
    return out


def detect_sharded(model, x_host: torch.Tensor) -> List[Detection]:
    """Every rank passes the SAME host batch; each runs `model.detect` on its shard on its
    own GPU and all ranks return the full batch's detections."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(x_host.shape[0], world, rank)
    dev = next(model.parameters()).device
    local = model.detect(x_host[lo:hi].to(dev, non_blocking=True)) if hi > lo else []
    return gather_detections(local, device=dev if dist.is_initialized() and dist.get_backend() == "nccl"
                             else torch.device("cpu"))
