"""Tracklet point extraction (first half): ``tools/ctrl/generate_track_input.py:69-117``.

For every timestamp of a segment the reference loads the frame cloud, enlarges each tracklet's box of that
timestamp by ``extra_width`` (``LiDARInstance3DBoxes.enlarged_box``, lidar_box3d.py:269-285), keeps the points
inside it (``points_in_boxes`` with ONE box at a time, so a point may go to several tracklets) and appends them to
the tracklet's ``pc_list``; at the end ``np.save``-s each list as ``<segment>--<id>.npy``.  Here all boxes of a
timestamp are tested in one ``points_in_boxes_batch`` launch (the same in-box test, A1) and the rows are gathered
per box in their original order.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .points_in_boxes import points_in_boxes_batch


def enlarged_boxes(boxes7: np.ndarray, extra_width: float) -> np.ndarray:
    """lidar_box3d.py:269-285 on f32 [K,7] rows: sizes += 2*extra, z_bottom -= extra; with a negative
    ``extra_width`` a box whose size would become <= 0 is left unchanged."""
    b = np.array(boxes7, np.float32, copy=True)
    out = b.copy()
    out[:, 3:6] += np.float32(extra_width * 2)
    out[:, 2] -= np.float32(extra_width)
    if extra_width < 0:
        bad = (out[:, 3:6] <= 0).any(1)
        out[bad] = b[bad]
    return out


def crop_frame(pc, boxes7: np.ndarray, extra_width: float, host_trig: bool = False) -> List[torch.Tensor]:
    """One timestamp (generate_track_input.py:90-101): ``pc`` f32 [M,C>=3] (numpy or CUDA tensor), ``boxes7`` f32
    [K,7] -> K CUDA tensors [n_k, C], the rows of ``pc`` inside each enlarged box, original order."""
    _lib.require_cuda()
    pc_d = pc if torch.is_tensor(pc) else torch.from_numpy(np.ascontiguousarray(pc, np.float32)).cuda()
    K = len(boxes7)
    if K == 0:
        return []
    if pc_d.shape[0] == 0:
        return [pc_d[:0] for _ in range(K)]
    bx = torch.from_numpy(enlarged_boxes(boxes7, extra_width)).to(pc_d.device)
    mask = points_in_boxes_batch(pc_d[None, :, :3].contiguous(), bx[None], host_trig=host_trig)[0]      # [M,K]
    box_id, pt_id = mask.t().nonzero(as_tuple=True)              # sorted by box, then by point index
    counts = torch.bincount(box_id, minlength=K).tolist()
    return list(torch.split(pc_d[pt_id], counts))


def extract_segment(tracklets: Sequence[dict], frames: Dict[int, np.ndarray], extra_width: float,
                    host_trig: bool = False) -> List[List[np.ndarray]]:
    """A segment (generate_track_input.py:78-101).  ``tracklets[i] = dict(ts=[...], boxes=f32 [len,7])``;
    ``frames[ts]`` = the frame cloud f32 [M,6].  Returns each tracklet's ``pc_list`` (one array per timestamp of
    the tracklet, in its own order)."""
    out: List[List[Optional[np.ndarray]]] = [[None] * len(t["ts"]) for t in tracklets]
    where: Dict[int, List] = {}
    for i, t in enumerate(tracklets):
        for k, ts in enumerate(t["ts"]):
            where.setdefault(ts, []).append((i, k))
    for ts in sorted(where):
        boxes = np.stack([tracklets[i]["boxes"][k] for i, k in where[ts]], 0).astype(np.float32)
        for (i, k), pts in zip(where[ts], crop_frame(frames[ts], boxes, extra_width, host_trig)):
            out[i][k] = pts.cpu().numpy()
    return out


def save_pc_list(save_dir: str, segment_name: str, trk_id: str, pc_list: Sequence[np.ndarray]) -> str:
    """``np.save(fw, pc)`` of generate_track_input.py:107-109: a 1-D object array, one entry per timestamp."""
    arr = np.empty(len(pc_list), dtype=object)
    for k, p in enumerate(pc_list):
        arr[k] = p
    path = os.path.join(save_dir, f"{segment_name}--{trk_id}.npy")
    os.makedirs(save_dir, exist_ok=True)
    with open(path, "wb") as fw:
        np.save(fw, arr, allow_pickle=True)
    return path


def load_pc_list(path: str) -> List[np.ndarray]:
    return list(np.load(path, allow_pickle=True))
