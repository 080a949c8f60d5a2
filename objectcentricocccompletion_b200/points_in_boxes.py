"""``points_in_boxes_gpu`` / ``points_in_boxes_batch`` -- drop-ins for
``mmdet3d/ops/roiaware_pool3d/points_in_boxes.py:6-50, 86-123``.

Same arguments, asserts and return layout as the reference wrappers; the kernels are in
``csrc/points_in_boxes.cu``.  Two arithmetics, both bit-exact against their reference:

* default (``host_trig=False``): what the reference's CUDA kernel computes -- ``cosf/sinf(rz + pi/2)`` on the
  device and the FMA contraction nvcc applies to ``local_x`` / ``local_y`` (points_in_boxes_cuda.cu:24-49; equal to
  the unmodified kernel built for sm_100a, tests/test_ref_cuda_gpu.py); no synchronisation;
* ``host_trig=True``: what ``points_in_boxes_cpu`` computes -- host libm trig (one D2H of the boxes) and unfused
  products (points_in_boxes_cpu.cpp:16-41); the arithmetic of the annotate path's crop and of the CPU oracle.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _check(points, boxes):
    assert boxes.shape[0] == points.shape[0], \
        f'Points and boxes should have the same batch size, got {boxes.shape[0]} and {points.shape[0]}'
    assert boxes.shape[2] == 7, f'boxes dimension should be 7, got unexpected shape {boxes.shape[2]}'
    assert points.shape[2] == 3, f'points dimension should be 3, got unexpected shape {points.shape[2]}'
    _lib.require_cuda(points, boxes)
    assert points.get_device() == boxes.get_device(), 'Points and boxes should be put on the same device'


def _trig(boxes, host_trig):
    if not host_trig:
        return None
    b = np.ascontiguousarray(boxes.detach().float().cpu().numpy().reshape(-1, 7))
    out = np.zeros((b.shape[0], 2), np.float32)
    _lib.lib().occb200_host_box_trig(b.ctypes.data, b.shape[0], out.ctypes.data)
    return torch.from_numpy(out).to(boxes.device)


def _run(fn_name, points, boxes, out, host_trig):
    pts = points.float().contiguous()
    bx = boxes.float().contiguous()
    trig = _trig(bx, host_trig)
    B, M, _ = pts.shape
    T = bx.shape[1]
    with torch.cuda.device(pts.device):
        rc = getattr(_lib.lib(), fn_name)(bx.data_ptr(), pts.data_ptr(), _lib.ptr(trig), out.data_ptr(), B, T, M,
                                          0 if host_trig else 1, _lib.stream_ptr(pts.device))
    _lib.check(rc, fn_name)
    return out


def points_in_boxes_gpu(points, boxes, host_trig: bool = False):
    """points [B,M,3], boxes [B,T,7] (x,y,z_bottom,w,l,h,ry) -> int32 [B,M], background -1."""
    _check(points, boxes)
    B, M, _ = points.shape
    out = points.new_empty((B, M), dtype=torch.int)
    if boxes.shape[1] == 0:
        return out.fill_(-1)
    return _run("occb200_points_in_boxes_gpu", points, boxes, out, host_trig)


def points_in_boxes_batch(points, boxes, host_trig: bool = False):
    """points [B,M,3], boxes [B,T,7] -> int32 [B,M,T], background 0."""
    _check(points, boxes)
    B, M, _ = points.shape
    out = points.new_empty((B, M, boxes.shape[1]), dtype=torch.int)
    if boxes.shape[1] == 0:
        return out
    return _run("occb200_points_in_boxes_batch", points, boxes, out, host_trig)
