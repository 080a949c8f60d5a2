"""Sharding of an annotation job across GPUs and the final gather of its labels.

The reference shards the annotation job over worker processes with no communication at all: each worker
takes whole segments, picks a GPU with ``wid % ngpus`` and leaves its results in files
(tools/occ/occ_annotate.py:320-322, 649-671).  Here: one process per GPU, work balanced by predicted cost
(longest-processing-time greedy) either by *segment* (what a job over many segments uses: a segment's range
images then live on exactly one GPU) or by *tracklet* (independent by construction: a tracklet's grid depends only
on its own points / poses and read-only range images), and ONE collective at the very end: the labels, one byte
per voxel, straight from device memory to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests).  The payload
sizes follow from the label offsets every rank computed on the host; they are exchanged once when the job is
set up, never per gather.  Nothing is exchanged on the data path.
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
    """Longest-processing-time greedy: heaviest unit first onto the lightest rank.  Deterministic."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(v) for v in out]


def shard_batch(batch, rank: int, world_size: int, by: str = "tracklet"):
    """The sub-batch of ``rank`` and the global tracklet indices it holds.

    ``by="tracklet"``: LPT over tracklets; ``by="segment"``: LPT over whole segments (cost = sum of its
    tracklets').  Either way the sub-batch keeps ONLY the segments its tracklets reference (a rank never
    uploads range images it does not read) and ``Tracklet.segment`` is remapped to the kept list."""
    import copy

    L = len(batch.segments[0].inclinations) if batch.segments else 5
    costs = [tracklet_cost(t, batch.voxel_size, L) for t in batch.tracklets]
    if by == "segment":
        seg_cost = [0.0] * len(batch.segments)
        for t, c in zip(batch.tracklets, costs):
            seg_cost[t.segment] += c
        segs = set(shard_indices(seg_cost, world_size)[rank])
        mine = [i for i, t in enumerate(batch.tracklets) if t.segment in segs]
    else:
        mine = shard_indices(costs, world_size)[rank]
    used = sorted({batch.tracklets[i].segment for i in mine})
    remap = {s: k for k, s in enumerate(used)}
    trks = []
    for i in mine:
        t = copy.copy(batch.tracklets[i])
        t.segment = remap[t.segment]
        trks.append(t)
    sub = type(batch)(segments=[batch.segments[s] for s in used], tracklets=trks, voxel_size=batch.voxel_size)
    return sub, mine


def _device_of(group) -> torch.device:
    return (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl"
            else torch.device("cpu"))


def exchange_sizes(n_local: int, group=None) -> List[int]:
    """Job set-up: every rank learns every rank's payload size (one all_gather of one integer)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [int(n_local)]
    dev = _device_of(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(dist.get_world_size(group))]
    dist.all_gather(sizes, torch.tensor([int(n_local)], dtype=torch.int64, device=dev), group=group)
    return [int(s) for s in sizes]


def gather_labels(local: torch.Tensor, sizes: Sequence[int], dst: int = 0, group=None,
                  out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """The one collective of a job: every rank's uint8 label buffer (device memory, sizes[r] bytes) lands at
    offset sum(sizes[:r]) of ``out`` on group rank ``dst``.  Exact sizes (point-to-point send / receive batched
    into one group call), no padding, no host staging.  Returns ``out`` on ``dst``, None elsewhere."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    offs = np.concatenate([[0], np.cumsum(np.asarray(sizes, np.int64))])
    assert local.dtype == torch.uint8 and local.numel() >= sizes[rank]
    if rank == dst:
        if out is None:
            out = torch.empty(max(int(offs[-1]), 1), dtype=torch.uint8, device=local.device)
        out[int(offs[rank]): int(offs[rank + 1])].copy_(local[: sizes[rank]], non_blocking=True)
    if world > 1:
        g = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
        ops = []
        if rank == dst:
            ops = [dist.P2POp(dist.irecv, out[int(offs[r]): int(offs[r + 1])], g(r), group)
                   for r in range(world) if r != dst and sizes[r] > 0]
        elif sizes[rank] > 0:
            ops = [dist.P2POp(dist.isend, local[: sizes[rank]], g(dst), group)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
    return out if rank == dst else None


def gather_results(local: List[dict], indices: List[int], total: int, dst: int = 0, group=None,
                   device: Optional[torch.device] = None) -> Optional[List[dict]]:
    """Result dicts (host) of every rank -> the list of ``total`` results in global order on group rank ``dst``
    (None elsewhere): one exchange of sizes, one gather of a small int32 header per rank and one
    ``gather_labels`` of the uint8 labels."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if device is None:
        device = _device_of(group)
    from .occ_annotate import STATUS_NAMES

    code = {v: k for k, v in STATUS_NAMES.items()}
    head, body = [], []
    for gi, r in zip(indices, local):
        ok = r["occ"] is not None
        dims = [int(v) for v in r["dims"]] if ok else [0, 0, 0]
        size = np.asarray(r.get("size", np.zeros(3)), np.float32).view(np.int32)
        steps = int(r.get("n_steps", 0))
        head += [gi, code.get(r["status"], -2), *dims, int(r.get("n_unknown", 0)), *[int(v) for v in size],
                 steps & 0x7fffffff, steps >> 31]
        if ok:
            body.append(np.ascontiguousarray(r["occ"]).astype(np.uint8).reshape(-1))
    HW = 11
    hd = torch.from_numpy(np.asarray(head, np.int64).astype(np.int32)).to(device)
    lab = torch.from_numpy(np.concatenate(body) if body else np.zeros(0, np.uint8)).to(device)
    n_head = exchange_sizes(hd.numel() * 4, group)
    n_lab = exchange_sizes(lab.numel(), group)
    heads = gather_labels(hd.view(torch.uint8) if hd.numel() else torch.zeros(0, dtype=torch.uint8, device=device),
                          n_head, dst, group)
    labs = gather_labels(lab, n_lab, dst, group)
    if rank != dst:
        return None
    H = heads[: sum(n_head)].cpu().numpy().view(np.int32).reshape(-1, HW)
    Lb = labs[: sum(n_lab)].cpu().numpy()
    out: List[Optional[dict]] = [None] * total
    pos = 0
    for gi, st, X, Y, Z, nu, s0, s1, s2, lo, hi in H:
        occ = None
        if st == 0:
            occ = Lb[pos: pos + X * Y * Z].reshape(X, Y, Z).astype(np.int32)
            pos += X * Y * Z
        out[int(gi)] = dict(status=STATUS_NAMES.get(int(st), str(st)), occ=occ, dims=np.array([X, Y, Z], np.int32),
                            size=np.array([s0, s1, s2], np.int32).view(np.float32), n_unknown=int(nu),
                            n_steps=int(lo) | (int(hi) << 31))
    return out  # type: ignore[return-value]


def annotate_distributed(batch, annotate_fn: Optional[Callable] = None, dst: int = 0, group=None,
                         by: str = "tracklet"):
    """Shard ``batch`` over the ranks of ``group``, annotate locally, gather on ``dst``.

    ``annotate_fn`` defaults to the CUDA path (``occ_annotate.annotate_batch``); the CPU tests inject the
    oracle to exercise the sharding / gather logic without a GPU.
    """
    if annotate_fn is None:
        from .occ_annotate import annotate_batch as annotate_fn
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sub, mine = shard_batch(batch, rank, world, by=by)
    local = annotate_fn(sub)
    return gather_results(local, mine, len(batch.tracklets), dst=dst, group=group)
