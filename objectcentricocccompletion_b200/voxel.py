"""``Voxelization`` / ``voxelization`` / ``DynamicScatter`` / ``dynamic_scatter`` -- drop-ins for
``mmdet3d/ops/voxel/{voxelize.py,scatter_points.py}`` (re-exported by ``mmdet3d/ops/voxel/__init__.py``).

The pybind module ``voxel_layer`` of the reference (src/voxelization.cpp:6-11) is mirrored by the
four module-level functions ``hard_voxelize``, ``dynamic_voxelize``, ``dynamic_point_to_voxel_forward``
and ``dynamic_point_to_voxel_backward`` with the same argument order and ownership rules (the caller
allocates and zero-fills the voxelize outputs; the callee allocates for scatter forward); they call
the C ABI of ``include/occ_b200.h``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from . import _lib

_REDUCE = {"sum": 0, "mean": 1, "max": 2}


def _f3(v, n):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(-1))
    assert a.size == n
    return a


# ------------------------------------------------------------------ voxel_layer mirror
def dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3):
    """voxel_layer.dynamic_voxelize (voxelization.cpp:8): fills ``coors`` int32 [N,3] (z,y,x)."""
    assert NDim == 3
    _lib.require_cuda(points, coors)
    if not points.is_contiguous():
        raise RuntimeError("points must be contiguous")
    assert coors.dtype == torch.int32 and coors.is_contiguous()
    if points.dtype == torch.float32:
        dt = 0
    elif points.dtype == torch.float64:
        dt = 1
    else:
        raise RuntimeError(f"dynamic_voxelize: unsupported point dtype {points.dtype}")
    vs, cr = _f3(voxel_size, 3), _f3(coors_range, 6)
    with torch.cuda.device(points.device):
        rc = _lib.lib().occb200_dynamic_voxelize(points.data_ptr(), dt, points.size(0), points.size(1), vs.ctypes.data,
                                                 cr.ctypes.data, coors.data_ptr(), _lib.stream_ptr(points.device))
    _lib.check(rc, "occb200_dynamic_voxelize")


def hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range, max_points, max_voxels,
                  NDim=3):
    """voxel_layer.hard_voxelize (voxelization.cpp:7): returns the number of voxels."""
    assert NDim == 3
    _lib.require_cuda(points, voxels, coors, num_points_per_voxel)
    if points.dtype != torch.float32:
        raise RuntimeError("hard_voxelize supports float32 points (as the reference's copy kernels)")
    assert points.is_contiguous() and voxels.is_contiguous() and coors.is_contiguous()
    vs, cr = _f3(voxel_size, 3), _f3(coors_range, 6)
    N = points.size(0)
    L = _lib.lib()
    wsb = L.occb200_hard_voxelize_workspace_bytes(N)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=points.device)
    n_out = C.c_int(0)
    with torch.cuda.device(points.device):
        rc = L.occb200_hard_voxelize(points.data_ptr(), N, points.size(1), vs.ctypes.data, cr.ctypes.data,
                                     int(max_points), int(max_voxels), voxels.data_ptr(), coors.data_ptr(),
                                     num_points_per_voxel.data_ptr(), ws.data_ptr(), ws.numel(),
                                     C.addressof(n_out), _lib.stream_ptr(points.device))
    _lib.check(rc, "occb200_hard_voxelize")
    return int(n_out.value)


class _Plan:
    """Sorted-unique result + reduction plan kept on the device."""
    __slots__ = ("uniq", "inverse", "counts", "order", "gstart", "M")


def _unique(coors, mode, col_max=None):
    """``col_max`` (optional, modes 1 / 2): the largest valid value of each column, from the caller's grid -- the sort
    keys are then laid out without a min/max pass over the rows (occb200_unique_rows_bounded)."""
    N, K = coors.shape
    dev = coors.device
    if coors.dtype == torch.int32:
        dt = 0
    elif coors.dtype == torch.int64:
        dt = 1
    else:
        raise RuntimeError(f"unsupported coordinate dtype {coors.dtype}")
    L = _lib.lib()
    p = _Plan()
    uniq = torch.empty((N, K), dtype=coors.dtype, device=dev)
    p.inverse = torch.empty(N, dtype=torch.int32, device=dev)
    counts = torch.empty(N, dtype=torch.int32, device=dev)
    p.order = torch.empty(N, dtype=torch.int32, device=dev)
    gstart = torch.empty(N, dtype=torch.int32, device=dev)
    ws = torch.empty(max(L.occb200_unique_workspace_bytes(N, K), 16), dtype=torch.uint8, device=dev)
    m = C.c_int64(0)
    cm = None
    if col_max is not None and mode != 0:
        cm = np.ascontiguousarray(np.asarray(col_max, np.int64).reshape(-1))
        assert cm.size == K
    with torch.cuda.device(dev):
        rc = L.occb200_unique_rows_bounded(coors.data_ptr(), dt, N, K, mode, cm.ctypes.data if cm is not None else None,
                                           uniq.data_ptr(), p.inverse.data_ptr(), counts.data_ptr(),
                                           p.order.data_ptr(), gstart.data_ptr(), ws.data_ptr(), ws.numel(),
                                           C.addressof(m), _lib.stream_ptr(dev))
    _lib.check(rc, "occb200_unique_rows")
    p.M = int(m.value)
    p.uniq, p.counts, p.gstart = uniq[:p.M], counts[:p.M], gstart[:p.M]
    return p


def _reduce(feats, plan, reduce, want_argmax=False):
    N, Cc = feats.shape
    dev = feats.device
    out = torch.empty((plan.M, Cc), dtype=feats.dtype, device=dev)
    argmax = torch.empty((plan.M, Cc), dtype=torch.int32, device=dev) if want_argmax else None
    fn = _lib.lib().occb200_segment_reduce if feats.dtype == torch.float32 else _lib.lib().occb200_segment_reduce_f64
    with torch.cuda.device(dev):
        rc = fn(feats.data_ptr(), N, Cc, plan.order.data_ptr(), plan.gstart.data_ptr(), plan.counts.data_ptr(), plan.M,
                reduce, out.data_ptr(), _lib.ptr(argmax), _lib.stream_ptr(dev))
    _lib.check(rc, "occb200_segment_reduce")
    return out, argmax


def _convert_reduce_type(reduce_type):
    if reduce_type not in _REDUCE:
        raise RuntimeError("do not support reduce type " + str(reduce_type))      # voxelization.h:92
    return _REDUCE[reduce_type]


def dynamic_point_to_voxel_forward(feats, coors, reduce_type, _mode=1, _col_max=None):
    """voxel_layer.dynamic_point_to_voxel_forward (voxelization.cpp:9, scatter_points_cuda.cu:183-234).

    Returns [voxel_feats, voxel_coors, point2voxel_map int32, voxel_points_count int32, argmax];
    the fifth entry (int32 [M,C] for 'max', else None) is an extension that lets backward skip the
    reference's arg search.
    """
    red = _convert_reduce_type(reduce_type)
    _lib.require_cuda(feats, coors)
    if not (feats.is_contiguous() and coors.is_contiguous()):
        raise RuntimeError("feats and coors must be contiguous")
    if feats.size(0) == 0:                                                          # :192-196
        return [feats.clone().detach(), coors.clone().detach(),
                coors.new_empty((0,), dtype=torch.int32), coors.new_empty((0,), dtype=torch.int32), None]
    if feats.dtype not in (torch.float32, torch.float64):       # AT_DISPATCH_FLOATING_TYPES (scatter_points_cuda.cu:215)
        raise RuntimeError("dynamic_point_to_voxel_forward supports float32 / float64 features")
    plan = _unique(coors, _mode, _col_max)
    out, argmax = _reduce(feats, plan, red, want_argmax=(red == 2))
    return [out, plan.uniq, plan.inverse, plan.counts, argmax]


def dynamic_point_to_voxel_backward(grad_feats, grad_reduced_feats, feats, reduced_feats, coors_idx, reduce_count,
                                    reduce_type, argmax=None):
    """voxel_layer.dynamic_point_to_voxel_backward (voxelization.cpp:10, scatter_points_cuda.cu:236-303)."""
    red = _convert_reduce_type(reduce_type)
    _lib.require_cuda(grad_feats, grad_reduced_feats)
    N, Cc = feats.shape
    M = reduced_feats.size(0)
    if N == 0:
        return
    fn = (_lib.lib().occb200_segment_reduce_backward if grad_feats.dtype == torch.float32
          else _lib.lib().occb200_segment_reduce_backward_f64)
    with torch.cuda.device(grad_feats.device):
        rc = fn(
            grad_feats.data_ptr(), grad_reduced_feats.data_ptr(), feats.data_ptr(), reduced_feats.data_ptr(),
            coors_idx.data_ptr(), reduce_count.data_ptr(), _lib.ptr(argmax), N, M, Cc, red,
            _lib.stream_ptr(grad_feats.device))
    _lib.check(rc, "occb200_segment_reduce_backward")


# ------------------------------------------------------------------ voxelize.py mirror
class _Voxelization(Function):

    @staticmethod
    def forward(ctx, points, voxel_size, coors_range, max_points=35, max_voxels=20000):
        """points [N, ndim] -> dynamic: coors int32 [N,3]; hard: (voxels, coors, num_points_per_voxel)
        (voxelize.py:12-58)."""
        if max_points == -1 or max_voxels == -1:
            coors = points.new_zeros(size=(points.size(0), 3), dtype=torch.int)
            dynamic_voxelize(points, coors, voxel_size, coors_range, 3)
            return coors
        voxels = points.new_zeros(size=(max_voxels, max_points, points.size(1)))
        coors = points.new_zeros(size=(max_voxels, 3), dtype=torch.int)
        num_points_per_voxel = points.new_zeros(size=(max_voxels,), dtype=torch.int)
        voxel_num = hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range, max_points,
                                  max_voxels, 3)
        return voxels[:voxel_num], coors[:voxel_num], num_points_per_voxel[:voxel_num]


voxelization = _Voxelization.apply


class Voxelization(nn.Module):
    """Same constructor and forward as voxelize.py:64-122."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else _pair(max_voxels)
        pcr = torch.tensor(point_cloud_range, dtype=torch.float32)
        vs = torch.tensor(voxel_size, dtype=torch.float32)
        grid_size = torch.round((pcr[3:] - pcr[:3]) / vs).long()
        self.grid_size = grid_size
        self.pcd_shape = [*grid_size[:2], 1][::-1]

    def forward(self, input):
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        return voxelization(input, self.voxel_size, self.point_cloud_range, self.max_num_points, max_voxels)

    def __repr__(self):
        return (f"{self.__class__.__name__}(voxel_size={self.voxel_size}, point_cloud_range={self.point_cloud_range}"
                f", max_num_points={self.max_num_points}, max_voxels={self.max_voxels})")


# ------------------------------------------------------------------ scatter_points.py mirror
class _dynamic_scatter(Function):

    @staticmethod
    def forward(ctx, feats, coors, reduce_type='max', _mode=1, _col_max=None):
        """feats [N,C], coors [N,ndim] int -> (voxel_feats [M,C], voxel_coors [M,ndim]) (scatter_points.py:11-34)."""
        voxel_feats, voxel_coors, point2voxel_map, voxel_points_count, argmax = dynamic_point_to_voxel_forward(
            feats, coors, reduce_type, _mode, _col_max)
        ctx.reduce_type = reduce_type
        ctx.has_argmax = argmax is not None
        saved = [feats, voxel_feats, point2voxel_map, voxel_points_count]
        if argmax is not None:
            saved.append(argmax)
        ctx.save_for_backward(*saved)
        ctx.mark_non_differentiable(voxel_coors)
        return voxel_feats, voxel_coors

    @staticmethod
    def backward(ctx, grad_voxel_feats, grad_voxel_coors=None):
        saved = ctx.saved_tensors
        feats, voxel_feats, point2voxel_map, voxel_points_count = saved[:4]
        argmax = saved[4] if ctx.has_argmax else None
        grad_feats = torch.zeros_like(feats)
        dynamic_point_to_voxel_backward(grad_feats, grad_voxel_feats.contiguous(), feats, voxel_feats,
                                        point2voxel_map, voxel_points_count, ctx.reduce_type, argmax)
        return grad_feats, None, None, None, None


def dynamic_scatter(feats, coors, reduce_type='max'):
    return _dynamic_scatter.apply(feats, coors, reduce_type, 1, None)


class DynamicScatter(nn.Module):
    """Same constructor and forward as scatter_points.py:53-107.

    4-column (batched) coordinates take ONE sort for the whole batch instead of the reference's
    python loop over samples (:86-99); the per-sample "drop the first unique row" behaviour
    (scatter_points_cuda.cu:207-210) is preserved by the kernel (unique mode 2).
    """

    def __init__(self, voxel_size, point_cloud_range, average_points: bool):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.average_points = average_points
        # the grid the coordinates come from (Voxelization with the same arguments, voxelize.py:93-98): bounds of the
        # (z, y, x) columns for the sort keys; coordinates beyond them are detected and handled by the library
        self._col_max = None
        try:
            pcr = np.asarray(point_cloud_range, np.float32).reshape(-1)
            vs = np.asarray(voxel_size, np.float32).reshape(-1)
            grid = np.round((pcr[3:6] - pcr[:3]) / vs[:3]).astype(np.int64)
            if grid.size == 3 and (grid >= 1).all() and (grid < 2 ** 20).all():
                self._col_max = [int(grid[2]) - 1, int(grid[1]) - 1, int(grid[0]) - 1]
        except Exception:                                            # noqa: BLE001  (odd arguments: no bounds)
            self._col_max = None

    def forward_single(self, points, coors):
        reduce = 'mean' if self.average_points else 'max'
        return _dynamic_scatter.apply(points.contiguous(), coors.contiguous(), reduce, 1, self._col_max)

    def forward(self, points, coors):
        if coors.size(-1) == 3:
            return self.forward_single(points, coors)
        reduce = 'mean' if self.average_points else 'max'
        if coors.size(0) == 0:
            raise IndexError("index -1 is out of bounds for dimension 0 with size 0")   # coors[-1, 0] in the reference
        cm = [0] + self._col_max if (self._col_max is not None and coors.size(-1) == 4) else None
        return _dynamic_scatter.apply(points.contiguous(), coors.contiguous(), reduce, 2, cm)

    def __repr__(self):
        return (f"{self.__class__.__name__}(voxel_size={self.voxel_size}, point_cloud_range={self.point_cloud_range}"
                f", average_points={self.average_points})")
