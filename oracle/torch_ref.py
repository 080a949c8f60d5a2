"""Reference-driven golden generator (TEST INFRASTRUCTURE; runs only where /root/reference exists).

The annotate path of the reference is Python + torch ops, so "the reference run
here" means: execute the reference's own ``point_cloud_to_range_image_idx``
(AST-extracted from ``/root/reference/tools/occ/occ_annotate.py`` -- the module
itself cannot be imported: argparse at import time, mmcv/mmdet asserts) on CPU
tensors, glued by the same torch op sequence ``annotate_trk`` /
``get_local_point_list`` perform (occ_annotate.py:107-136, 413-471, 472-568),
with the reference-compiled ``points_in_boxes_cpu`` (oracle/_ref) as the in-box test.

This is what pins oracle/occ_oracle.c; its outputs are committed as fixtures
under tests/golden/ by oracle/make_golden.py.  Nothing here ships.
"""
from __future__ import annotations

import ast
import os

import numpy as np
import torch

from . import build as _build

REF_FILE = "/root/reference/tools/occ/occ_annotate.py"
LIDAR_NAME_LIST = ["TOP", "FRONT", "SIDE_LEFT", "SIDE_RIGHT", "REAR"]   # occ_annotate.py:235


def available() -> bool:
    return os.path.isfile(REF_FILE)


_fn = None


def reference_projection_fn():
    """The reference's point_cloud_to_range_image_idx, exec'd from its source file."""
    global _fn
    if _fn is None:
        tree = ast.parse(open(REF_FILE).read())
        node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "point_cloud_to_range_image_idx")
        ns = {"torch": torch, "np": np}
        exec(compile(ast.Module([node], []), REF_FILE, "exec"), ns)
        _fn = ns["point_cloud_to_range_image_idx"]
    return _fn


def _points_in_box(pc, box7):
    """box.points_in_boxes(pc) for a single box: index 0 inside, -1 outside (lidar_box3d.py:348-360)."""
    ext = _build.load_ref("ref_points_in_boxes")
    out = torch.zeros((1, pc.shape[0]), dtype=torch.int32)
    ext.points_in_boxes_cpu(box7.reshape(1, 7).float().contiguous(), pc.float().contiguous(), out)
    return out[0] - 1


def local_point_list(trk):
    """get_local_point_list with box_mode='max' (occ_annotate.py:91-138)."""
    pcs, sizes = [], []
    for i in range(len(trk)):
        box = torch.from_numpy(trk.boxes[i:i + 1].copy())
        pc = torch.from_numpy(np.ascontiguousarray(trk.points[i][:, :3]))
        inside = pc[_points_in_box(pc, box[0]) == 0]
        if len(inside) == 0:
            continue
        shift = -box[:, :3]
        angle = -box[0, 6]
        loc = inside + shift
        s, c = torch.sin(angle), torch.cos(angle)
        rot_t = box.new_tensor([[c, -s, 0], [s, c, 0], [0, 0, 1]])        # lidar_box3d.py:165-167
        loc = loc @ rot_t                                                  # :184
        pcs.append(loc)
        sizes.append(box[:, 3:6])
    if not pcs:
        raise AssertionError("no points in the tracklet")
    size = torch.cat(sizes, 0).max(0)[0]
    return pcs, size


def annotate_tracklet(trk, segment, voxel_size):
    """annotate_trk (occ_annotate.py:344-568) on CPU tensors -> dict(status, occ, ...)."""
    if len(trk) < 10:
        return dict(status="skip_short", occ=None)
    try:
        pcs, size = local_point_list(trk)
    except AssertionError:
        return dict(status="no_points", occ=None)
    pc = torch.cat(pcs, 0)
    dims = torch.ceil(size / voxel_size).to(torch.int32)
    occ = torch.zeros((int(dims[0]), int(dims[1]), int(dims[2])), dtype=torch.bool)
    # corners of the canonical box [0,0,0,w,l,h,0] (lidar_box3d.py:54-92) -> min bound
    norm = torch.tensor([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], dtype=size.dtype)
    corners = size.view(1, 3) * (norm - size.new_tensor([0.5, 0.5, 0]))
    min_bound = corners.min(0)[0]
    q = torch.floor((pc - min_bound) / voxel_size).to(torch.long)
    keep = (q < dims[None]).all(1)
    pc_kept, q = pc[keep], q[keep]
    if q.shape[0] == 0:
        return dict(status="empty_after_filter", occ=None)
    try:
        occ[q[:, 0], q[:, 1], q[:, 2]] = True
    except IndexError:
        return dict(status="index_error", occ=None)
    gx, gy, gz = torch.meshgrid(*[torch.arange(int(d), dtype=torch.long) for d in dims], indexing="ij")
    coors = torch.stack([gx, gy, gz], -1).view(-1, 3)
    occ = occ.view(-1)
    un = coors[~occ]
    centers = un.to(torch.float64) * voxel_size + min_bound + voxel_size / 2
    label = torch.zeros_like(occ, dtype=torch.int32)
    if un.shape[0] > 0:
        project = reference_projection_fn()
        ego = []
        for i in range(len(trk)):
            box = torch.from_numpy(trk.boxes[i:i + 1].copy())
            s, c = torch.sin(box[0, 6]), torch.cos(box[0, 6])
            rot_t = torch.tensor([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=centers.dtype)
            ego.append(centers @ rot_t + box[:, :3])
        ego = torch.stack(ego, 0)
        fids = trk.frame_ids
        vis_all = []
        for ci in range(len(LIDAR_NAME_LIST)):
            extr = torch.tensor(segment.extrinsics[fids, ci])
            incl = np.stack([segment.inclinations[ci]] * len(fids), 0)
            incl = torch.tensor(np.flip(incl, axis=1).copy())
            ri = torch.tensor(segment.range_images[ci][fids])
            idx, rng = project(ego, extr, incl, ri.shape[1:])
            vals = torch.stack([ri[i, idx[i, :, 0], idx[i, :, 1]] for i in range(len(idx))], 0)
            vis = torch.zeros_like(vals, dtype=torch.int32)
            vis[vals >= rng] = 2
            vis_all.append(vis.max(0)[0])
        vis = torch.stack(vis_all, 0).max(0)[0]
        label[~occ] = vis
    label[occ] = 1
    return dict(status="ok", occ=label.view(int(dims[0]), int(dims[1]), int(dims[2])).numpy().copy(),
                dims=dims.numpy().copy(), size=size.numpy().copy(), n_unknown=int(un.shape[0]),
                loc=pc_kept.numpy().copy(), q=q.numpy().copy(), centers=centers.numpy().copy())


def annotate_batch(batch):
    return [annotate_tracklet(t, batch.segments[t.segment], batch.voxel_size) for t in batch.tracklets]


def mean_var(res):
    """The --save-mean-var block (occ_annotate.py:627-645) on the ``loc`` / ``q`` / ``dims`` of one
    ``annotate_tracklet`` result, with scatter_v2's torch calls (sst_ops.py:150-181): torch.unique(dim=0) and a
    scatter-mean written as index_add_ / count (torch_scatter itself is not installed; its CPU kernel
    accumulates in index order, which index_add_ on CPU also does)."""
    loc, q = torch.from_numpy(res["loc"]), torch.from_numpy(res["q"])
    dims = [int(v) for v in res["dims"]]
    new_coors, inv = torch.unique(q, return_inverse=True, dim=0)

    def scatter_mean(feat):
        out = torch.zeros((new_coors.shape[0], feat.shape[1]), dtype=feat.dtype).index_add_(0, inv, feat)
        cnt = torch.zeros(new_coors.shape[0], dtype=feat.dtype).index_add_(0, inv, torch.ones(len(inv), dtype=feat.dtype))
        return out / cnt.clamp(min=1)[:, None]

    mean = scatter_mean(loc)
    var = scatter_mean((loc - mean[inv]) ** 2)
    dense = torch.zeros((dims[0], dims[1], dims[2], 6), dtype=loc.dtype)
    dense[new_coors[:, 0], new_coors[:, 1], new_coors[:, 2], :] = torch.cat([mean, var], 1)
    return dense.numpy()
