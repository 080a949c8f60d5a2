"""Candidate selection from whole-frame clouds on the device.

The reference reads a frame's full point cloud (~180 k points, KITTI-format ``.bin``, tools/ctrl/utils.py:60-66) for
every (tracklet, frame) and masks it with ``points_in_boxes`` (tools/occ/occ_annotate.py:96-112).  A segment's job
needs each cloud once: ``select_candidates`` uploads the clouds of a segment, tests every point against the
candidate spheres of all boxes alive in its frame in ONE pass per frame (``csrc/candidates.cu``), and returns, per
tracklet-frame, the points inside the sphere around its box -- a superset of the in-box points; the exact in-box
test stays in the annotate kernel -- in cloud order, tracklet-major (the layout ``pack_tracklets`` uploads).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def candidate_spheres(boxes: np.ndarray, margin: float) -> np.ndarray:
    """boxes f32 [n,7] (x, y, z_bottom, w, l, h, yaw) -> [n,4] sphere (cx, cy, cz, r^2) containing each box + margin."""
    b = np.asarray(boxes, np.float32).reshape(-1, 7)
    out = np.empty((len(b), 4), np.float32)
    out[:, :2] = b[:, :2]
    out[:, 2] = b[:, 2] + np.float32(0.5) * b[:, 5]
    r = np.float32(0.5) * np.linalg.norm(b[:, 3:6], axis=1).astype(np.float32) + np.float32(margin)
    out[:, 3] = r * r
    return out


def select_candidates(clouds: Sequence[np.ndarray], trk_boxes: Sequence[np.ndarray], trk_frames: Sequence[np.ndarray],
                      margin: float = 0.5, device=None, out_stride: int = None) -> Tuple[np.ndarray, np.ndarray]:
    """clouds: per segment frame f an f32 [n_f, stride] array; trk_boxes[t] f32 [B_t,7]; trk_frames[t] int [B_t] =
    the segment frame of each tracklet-frame.  Returns (points f32 [P, out_stride] host, counts int64 [sum B_t])
    with the candidates of tracklet-frame i at rows [cumsum(counts)[i-1], cumsum(counts)[i]), cloud order."""
    _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    L_ = _lib.lib()
    NF = len(clouds)
    stride = int(clouds[0].shape[1]) if NF else 3
    out_stride = stride if out_stride is None else int(out_stride)
    sizes = np.array([len(c) for c in clouds], np.int64)
    cloud_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    nB = np.array([len(b) for b in trk_boxes], np.int64)
    F = int(nB.sum())
    if F == 0 or NF == 0 or cloud_off[-1] == 0:
        return np.zeros((0, out_stride), np.float32), np.zeros(F, np.int64)
    all_boxes = np.concatenate([np.asarray(b, np.float32).reshape(-1, 7) for b in trk_boxes], 0)
    all_frames = np.concatenate([np.asarray(f, np.int64).reshape(-1) for f in trk_frames])
    sph = candidate_spheres(all_boxes, margin)
    order = np.argsort(all_frames, kind="stable")                  # spheres grouped by frame: [frame][box]
    frame_box_off = np.concatenate([[0], np.cumsum(np.bincount(all_frames, minlength=NF))]).astype(np.int64)
    boxes = np.zeros(F, _lib.CAND_BOX_DTYPE)
    boxes["cx"], boxes["cy"], boxes["cz"], boxes["r2"] = sph[order, 0], sph[order, 1], sph[order, 2], sph[order, 3]
    boxes["tf"] = order.astype(np.int32)
    CH, W = L_.occb200_candidate_chunk(), L_.occb200_candidate_warps()
    nchunk = (sizes + CH - 1) // CH
    K = np.diff(frame_box_off)
    cells_per_frame = K * nchunk * W
    cnt_off = np.concatenate([[0], np.cumsum(cells_per_frame)]).astype(np.int64)
    n_cells = int(cnt_off[-1])
    # cell -> sphere (in [frame][box] order): every sphere of frame f owns nchunk_f * W consecutive cells
    cells_per_sphere = np.repeat(nchunk * W, K)
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        d_clouds = torch.empty((int(cloud_off[-1]), stride), dtype=torch.float32, device=dev)
        for f, c in enumerate(clouds):
            if len(c):
                d_clouds[int(cloud_off[f]): int(cloud_off[f + 1])].copy_(
                    torch.from_numpy(np.ascontiguousarray(c, np.float32)), non_blocking=True)
        d_off = torch.from_numpy(cloud_off).to(dev)
        d_boxes = torch.from_numpy(boxes.view(np.uint8).reshape(-1)).to(dev)
        d_fbo = torch.from_numpy(frame_box_off).to(dev)
        d_cnt_off = torch.from_numpy(cnt_off).to(dev)
        counts = torch.zeros(max(n_cells, 1), dtype=torch.int32, device=dev)
        args = (d_clouds.data_ptr(), stride, d_off.data_ptr(), NF, int(sizes.max()), d_boxes.data_ptr(), d_fbo.data_ptr(),
                d_cnt_off.data_ptr())
        _lib.check(L_.occb200_select_candidates(*args, counts.data_ptr(), None, None, out_stride, st),
                   "occb200_select_candidates")
        # start of every cell: spheres in tracklet-frame order, inside a sphere its cells in (chunk, warp) order
        c64 = counts[:n_cells].to(torch.int64)
        sphere_of_cell = torch.repeat_interleave(torch.arange(F, device=dev),
                                                 torch.from_numpy(cells_per_sphere).to(dev))
        total = torch.zeros(F, dtype=torch.int64, device=dev).index_add_(0, sphere_of_cell, c64)   # [frame][box] order
        tf_of_sphere = torch.from_numpy(order).to(dev)
        total_tf = torch.zeros(F, dtype=torch.int64, device=dev)
        total_tf[tf_of_sphere] = total
        base_tf = torch.cumsum(total_tf, 0) - total_tf
        incl = torch.cumsum(c64, 0)
        first_cell = torch.cumsum(torch.from_numpy(cells_per_sphere).to(dev), 0) - torch.from_numpy(cells_per_sphere).to(dev)
        before_sphere = (incl - c64)[first_cell.clamp(max=max(n_cells - 1, 0))]                      # cumsum at sphere start
        prefix = (incl - c64) - before_sphere[sphere_of_cell] + base_tf[tf_of_sphere][sphere_of_cell]
        P = int(total_tf.sum())
        out = torch.empty((max(P, 1), out_stride), dtype=torch.float32, device=dev)
        _lib.check(L_.occb200_select_candidates(*args, counts.data_ptr(), prefix.contiguous().data_ptr(), out.data_ptr(),
                                                out_stride, st), "occb200_select_candidates")
        pts = out[:P].cpu().numpy()
        cnt = total_tf.cpu().numpy()
    return pts, cnt


def split_candidates(points: np.ndarray, counts: np.ndarray, frames_per_tracklet: Sequence[int]):
    """(points, counts) of ``select_candidates`` -> per tracklet (flat array, list of per-frame views)."""
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    out: List[Tuple[np.ndarray, List[np.ndarray]]] = []
    i = 0
    for n in frames_per_tracklet:
        flat = points[off[i]: off[i + n]]
        out.append((flat, [points[off[i + k]: off[i + k + 1]] for k in range(n)]))
        i += n
    return out
