"""``dynamic_point_pool_mixed`` and ``TrackletPointRoIExtractor`` -- the step after annotation on the way to
OcCo-Net's input (second half).

Reference: mmdet3d/models/roi_heads/roi_extractors/dynamic_point_roi_extractor.py:149-300 calling
mmdet3d/ops/dynamic_point_pool_op.py:63-113.  The CUDA extension behind it (``dynamic_point_pool_ext``, from TorchEx)
is not in the reference tree: **parity unpinned**.  ``csrc/point_pool.cu`` restates it from what the reference itself
asserts about its output (the extractor's debug block, :214-230, reproduced in ``TrackletPointRoIExtractor``) and this
code base's box convention; the output order is (ROI, point index) and truncation keeps the lowest point indices
(the reference kernel's order is the arrival order of its atomics, "not strictly guaranteed").
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def dynamic_point_pool_mixed(rois, rois_batch, pts, pts_batch, extra_wlh, max_inbox_point, max_all_pts=200000):
    """rois f32 [N,7], rois_batch int [N], pts f32 [P,3], pts_batch int [P] -> (out_pts_idx int64 [K], out_roi_idx
    int64 [K], out_pts_feats f32 [K,13]): every (ROI, point) pair with equal batch index whose point lies inside the
    ROI enlarged by ``extra_wlh``; at most ``max_inbox_point`` per ROI and ``max_all_pts`` in total.  With no hit the
    reference returns one fake row (index -1, zero features: dynamic_point_pool_op.py:92-96); so does this."""
    _lib.require_cuda(rois, pts)
    assert len(rois) > 0
    dev = pts.device
    r = rois.float().contiguous()
    p = pts.float().contiguous()
    R, P = r.size(0), p.size(0)
    pb = pts_batch.to(dev).long().reshape(-1)
    rb = rois_batch.to(dev).long().reshape(-1)
    order = torch.sort(pb, stable=True)
    perm, pb_sorted = order.indices.contiguous(), order.values.contiguous()
    lo = torch.searchsorted(pb_sorted, rb, right=False).contiguous()
    hi = torch.searchsorted(pb_sorted, rb, right=True).contiguous()
    max_range = int((hi - lo).max().item()) if P else 0
    ex = np.ascontiguousarray(np.asarray(extra_wlh, np.float32).reshape(3))
    counts = torch.zeros(R, dtype=torch.int32, device=dev)
    L_ = _lib.lib()
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        args = (r.data_ptr(), lo.data_ptr(), hi.data_ptr(), R, p.data_ptr(), perm.data_ptr(), max_range, ex.ctypes.data)
        _lib.check(L_.occb200_point_pool(*args, counts.data_ptr(), None, None, None, None, None, st), "occb200_point_pool")
        c64 = counts.long()
        base = (torch.cumsum(c64, 0) - c64).contiguous()
        total = int(c64.sum().item())
        if total == 0:
            return (torch.full((1,), -1, dtype=torch.long, device=dev), torch.full((1,), -1, dtype=torch.long, device=dev),
                    torch.zeros((1, 13), dtype=torch.float32, device=dev))
        cursor = torch.zeros(R, dtype=torch.int32, device=dev)
        pidx = torch.empty(total, dtype=torch.long, device=dev)
        ridx = torch.empty(total, dtype=torch.long, device=dev)
        feats = torch.empty((total, 13), dtype=torch.float32, device=dev)
        _lib.check(L_.occb200_point_pool(*args, None, base.data_ptr(), cursor.data_ptr(), pidx.data_ptr(), ridx.data_ptr(),
                                         feats.data_ptr(), st), "occb200_point_pool")
    # (ROI, point index) order; the first max_inbox_point of each ROI, then the first max_all_pts overall
    srt = torch.sort(ridx * max(P, 1) + pidx).indices
    pidx, ridx, feats = pidx[srt], ridx[srt], feats[srt]
    rank = torch.arange(total, device=dev) - base[ridx]
    keep = rank < int(max_inbox_point)
    pidx, ridx, feats = pidx[keep], ridx[keep], feats[keep]
    if pidx.numel() > max_all_pts:
        pidx, ridx, feats = pidx[:max_all_pts], ridx[:max_all_pts], feats[:max_all_pts]
    return pidx, ridx, feats


class TrackletPointRoIExtractor(torch.nn.Module):
    """Mirror of ``TrackletPointRoIExtractor`` (dynamic_point_roi_extractor.py:149-300): same constructor, same
    ``forward`` signature and outputs ``(all_inds, all_roi_inds, ext_pts_info)``; the ``debug`` assertions are the
    reference's."""

    def __init__(self, init_cfg=None, debug=True, extra_wlh=[0, 0, 0], max_inbox_point=512, max_all_point=200000,
                 combined=False):
        super().__init__()
        self.debug = debug
        self.extra_wlh = extra_wlh
        self.max_inbox_point = max_inbox_point
        self.max_all_point = max_all_point
        self.combined = combined

    def forward(self, pts_xyz, batch_inds, pts_frame_inds, rois, roi_frame_inds, max_inbox_point=None):
        assert len(pts_xyz) > 0 and len(batch_inds) > 0 and len(rois) > 0
        if self.combined:                                           # :244-256
            pts_inds, roi_inds = batch_inds, rois[:, 0]
        else:                                                       # :176-198
            max_frames = roi_frame_inds.max().item() + 1
            pts_max_frames = pts_frame_inds.max().item() + 1
            assert pts_max_frames <= max_frames, f"{pts_max_frames} > {max_frames}"
            pts_inds = batch_inds * max_frames + pts_frame_inds
            roi_inds = rois[:, 0].int() * max_frames + roi_frame_inds
            assert len(roi_inds) == len(torch.unique(roi_inds))
        pts_inds, roi_inds = pts_inds.int(), roi_inds.int()
        if isinstance(self.max_all_point, (tuple, list)):
            max_all_point = self.max_all_point[0] if self.training else self.max_all_point[1]
        else:
            max_all_point = self.max_all_point
        all_inds, all_roi_inds, all_pts_info = dynamic_point_pool_mixed(
            rois[..., 1:], roi_inds, pts_xyz, pts_inds, self.extra_wlh, self.max_inbox_point, max_all_point)
        all_out_xyz = all_pts_info[:, :3]
        all_local_xyz = all_pts_info[:, 3:6]
        all_offset = all_pts_info[:, 6:-1]
        is_in_margin = all_pts_info[:, -1]
        if self.debug and bool((all_inds >= 0).all()):              # :214-230 / :275-288
            roi_per_pts = rois[..., 1:][all_roi_inds]
            assert torch.isclose(pts_xyz[all_inds], all_out_xyz).all()
            assert torch.isclose(all_offset[:, 0] + all_offset[:, 3], roi_per_pts[:, 4]).all()
            assert torch.isclose(all_offset[:, 1] + all_offset[:, 4], roi_per_pts[:, 3]).all()
            assert torch.isclose(all_offset[:, 2] + all_offset[:, 5], roi_per_pts[:, 5]).all()
            assert (all_local_xyz[:, 0].abs() < roi_per_pts[:, 4] + self.extra_wlh[0] + 1e-5).all()
            assert (all_local_xyz[:, 1].abs() < roi_per_pts[:, 3] + self.extra_wlh[1] + 1e-5).all()
            assert (all_local_xyz[:, 2].abs() < roi_per_pts[:, 5] + self.extra_wlh[2] + 1e-5).all()
            assert (roi_inds[all_roi_inds] == pts_inds[all_inds]).all()      # the per-ROI loop of :229-230, batched
        ext_pts_info = dict(local_xyz=all_local_xyz, boundary_offset=all_offset, is_in_margin=is_in_margin)
        return all_inds, all_roi_inds, ext_pts_info
