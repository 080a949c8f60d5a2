"""``quantize_points`` / ``generate_dense_voxel_centers`` -- drop-ins for ``mmdet3d/ops/occ/occ_ops.py:5-93``."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _f3(v):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(-1))
    assert a.size == 3
    return a


def quantize_points(points, rois, rois_points_idx, voxel_size, scale_wlh=[1.0, 1.0, 1.0],
                    offset_wlh=[0.0, 0.0, 0.0], to_center=False, check_index=True):
    """points [N,3] (ROI-local), rois [R,8|10], rois_points_idx [N] -> int64 [N,3] voxel coords, or the f32
    voxel centres when ``to_center`` (occ_ops.py:53-93).  ``rois_points_idx`` follows PyTorch indexing (negative
    indices count from the end); an index outside [-R, R) raises IndexError as in the reference
    (``check_index=False`` skips the device->host read of the error counter; bad rows then hold INT64_MIN / NaN)."""
    _lib.require_cuda(points, rois, rois_points_idx)
    pts = points.float().contiguous()
    r = rois.float().contiguous()
    idx = rois_points_idx.long().contiguous()
    N = pts.size(0)
    coor = None if to_center else torch.empty((N, 3), dtype=torch.long, device=pts.device)
    cen = torch.empty((N, 3), dtype=torch.float32, device=pts.device) if to_center else None
    sc, of = _f3(scale_wlh), _f3(offset_wlh)
    n_bad = torch.zeros(1, dtype=torch.int64, device=pts.device) if check_index else None
    with torch.cuda.device(pts.device):
        rc = _lib.lib().occb200_quantize_points(pts.data_ptr(), N, r.data_ptr(), r.size(0), r.size(1), idx.data_ptr(),
                                                float(voxel_size), sc.ctypes.data, of.ctypes.data, int(to_center),
                                                _lib.ptr(coor), _lib.ptr(cen), _lib.ptr(n_bad),
                                                _lib.stream_ptr(pts.device))
    _lib.check(rc, "occb200_quantize_points")
    if check_index and int(n_bad) != 0:
        raise IndexError(f"rois_points_idx: {int(n_bad)} indices out of range for {r.size(0)} rois")
    return cen if to_center else coor


def generate_dense_voxel_centers(bbox_sizes, voxel_size, scale_wlh=[1.0, 1.0, 1.0], offset_wlh=[0.0, 0.0, 0.0],
                                 as_volume=False):
    """bbox_sizes [R,3] -> list of f32 [X*Y*Z,3] (or [X,Y,Z,3]) voxel centres in the gravity-centred object
    frame (occ_ops.py:5-50).  One kernel for all R boxes instead of a python loop of meshgrids."""
    _lib.require_cuda(bbox_sizes)
    sizes = bbox_sizes.float().contiguous()
    R = sizes.size(0)
    if R == 0:
        return []
    sc, of = _f3(scale_wlh), _f3(offset_wlh)
    # grid dims on the host with the reference's f32 ops (it syncs per ROI as well: torch.arange(XS))
    s_host = sizes.cpu() * torch.from_numpy(sc) + torch.from_numpy(of)
    dims = torch.ceil(s_host / voxel_size).to(torch.int32)
    nvox = dims.long().prod(1)
    off = torch.zeros(R + 1, dtype=torch.long)
    off[1:] = torch.cumsum(nvox, 0)
    total = int(off[-1])
    dev = sizes.device
    centers = torch.empty((total, 3), dtype=torch.float32, device=dev)
    dims_d, off_d = dims.to(dev), off.to(dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().occb200_dense_voxel_centers(sizes.data_ptr(), dims_d.data_ptr(), off_d.data_ptr(), R, total,
                                                    float(voxel_size), sc.ctypes.data, of.ctypes.data,
                                                    centers.data_ptr(), _lib.stream_ptr(dev))
    _lib.check(rc, "occb200_dense_voxel_centers")
    out = []
    for r in range(R):
        c = centers[int(off[r]):int(off[r + 1])]
        if as_volume:
            c = c.view(int(dims[r, 0]), int(dims[r, 1]), int(dims[r, 2]), 3)
        out.append(c)
    return out


def jitter_voxel_center(voxel_size, voxel_centers):
    """occ_ops.py:96-100 (plain torch RNG; not on the kernel path)."""
    return voxel_centers + (torch.rand_like(voxel_centers) * voxel_size - voxel_size / 2)


def mirror_occ_label(occ_label_list):
    """``MirrorOccLabel.__call__`` (mmdet3d/datasets/pipelines/occ_pinelines.py:88-126) for a list of CUDA label
    grids int32 [X,Y,Z]: unknown voxels take the label of their mirror image across the x mid-plane; all grids of
    the list in one launch.  Returns new tensors (the inputs are not modified)."""
    if len(occ_label_list) == 0:
        return []
    _lib.require_cuda(*occ_label_list)
    dev = occ_label_list[0].device
    dims = np.array([list(g.shape) for g in occ_label_list], np.int32).reshape(-1, 3)
    sizes = dims.astype(np.int64).prod(1)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    flat = torch.cat([g.to(torch.int32).reshape(-1) for g in occ_label_list]) if off[-1] else \
        torch.zeros(1, dtype=torch.int32, device=dev)
    out = torch.empty_like(flat)
    off_d = torch.from_numpy(off).to(dev)
    dims_d = torch.from_numpy(dims).to(dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().occb200_mirror_occ_label(flat.data_ptr(), off_d.data_ptr(), dims_d.data_ptr(), None,
                                                 len(occ_label_list), int(sizes.max()), out.data_ptr(),
                                                 _lib.stream_ptr(dev))
    _lib.check(rc, "occb200_mirror_occ_label")
    return [out[int(off[i]): int(off[i + 1])].view(*[int(v) for v in dims[i]]).to(occ_label_list[i].dtype)
            for i in range(len(occ_label_list))]
