"""Tracklet sharding across GPUs and the final gather of results.

The reference shards the annotation job over worker processes with no communication at all: each worker
takes whole segments, picks a GPU with ``wid % ngpus`` and leaves its results in files
(tools/occ/occ_annotate.py:320-322, 649-671).  Here the unit is the *tracklet* (independent by
construction: a tracklet's grid depends only on its own points / poses and read-only range images), one
process per GPU, balanced by predicted cost, and ONE collective at the very end: a gather of the
flattened labels to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests).  Nothing is exchanged on
the data path.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def tracklet_cost(trk, voxel_size: float, num_lidars: int = 5) -> float:
    """Predicted work of a tracklet: grid voxels x frames x LiDARs (upper bound of the visibility tests)
    plus its candidate points."""
    if len(trk) == 0:
        return 0.0
    d = np.ceil(trk.boxes[:, 3:6].astype(np.float32).max(0) / np.float32(voxel_size))
    return float(d[0] * d[1] * d[2]) * len(trk) * num_lidars + float(sum(len(p) for p in trk.points))


def shard_indices(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time greedy: heaviest tracklet first onto the lightest rank.  Deterministic."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(v) for v in out]


def shard_batch(batch, rank: int, world_size: int):
    """The sub-batch of ``rank`` (segments are shared read-only) and the global tracklet indices it holds."""
    L = len(batch.segments[0].inclinations) if batch.segments else 5
    costs = [tracklet_cost(t, batch.voxel_size, L) for t in batch.tracklets]
    mine = shard_indices(costs, world_size)[rank]
    sub = type(batch)(segments=batch.segments, tracklets=[batch.tracklets[i] for i in mine],
                      voxel_size=batch.voxel_size)
    return sub, mine


def gather_results(local: List[dict], indices: List[int], total: int, dst: int = 0, group=None,
                   device: Optional[torch.device] = None) -> Optional[List[dict]]:
    """Final gather: every rank sends (global index, status, dims, labels) of its tracklets to ``dst``.

    One ``all_gather`` of the payload sizes and one ``gather`` of the padded payloads.  Returns the list of
    ``total`` results in global order on ``dst``, ``None`` elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    from .occ_annotate import STATUS_NAMES

    code = {v: k for k, v in STATUS_NAMES.items()}
    head, body = [], []
    for gi, r in zip(indices, local):
        dims = [int(v) for v in r["dims"]] if r["occ"] is not None else [0, 0, 0]
        head += [gi, code.get(r["status"], -2), *dims, int(r.get("n_unknown", 0))]
        if r["occ"] is not None:
            body.append(np.ascontiguousarray(r["occ"], np.int32).reshape(-1))
    payload = np.concatenate([np.asarray([len(indices)] + head, np.int32)] + body) if True else None
    mine = torch.from_numpy(payload).to(device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.numel()], dtype=torch.int64, device=device), group=group)
    mx = int(max(int(s) for s in sizes))
    padded = torch.zeros(mx, dtype=torch.int32, device=device)
    padded[: mine.numel()] = mine
    bufs = [torch.empty(mx, dtype=torch.int32, device=device) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out: List[Optional[dict]] = [None] * total
    for b, n in zip(bufs, sizes):
        a = b[: int(n)].cpu().numpy()
        k = int(a[0])
        hd = a[1: 1 + 6 * k].reshape(k, 6)
        pos = 1 + 6 * k
        for gi, st, X, Y, Z, nu in hd:
            if st == 0:
                occ = a[pos: pos + X * Y * Z].reshape(X, Y, Z).copy()
                pos += X * Y * Z
            else:
                occ = None
            out[int(gi)] = dict(status=STATUS_NAMES.get(int(st), str(st)), occ=occ, dims=np.array([X, Y, Z], np.int32),
                                n_unknown=int(nu))
    return out  # type: ignore[return-value]


def annotate_distributed(batch, annotate_fn: Optional[Callable] = None, dst: int = 0, group=None):
    """Shard ``batch`` by tracklet over the ranks of ``group``, annotate locally, gather on ``dst``.

    ``annotate_fn`` defaults to the CUDA path (``occ_annotate.annotate_batch``); the CPU tests inject the
    oracle to exercise the sharding / gather logic without a GPU.
    """
    if annotate_fn is None:
        from .occ_annotate import annotate_batch as annotate_fn
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sub, mine = shard_batch(batch, rank, world)
    local = annotate_fn(sub)
    return gather_results(local, mine, len(batch.tracklets), dst=dst, group=group)
