"""``scatter_v2`` -- drop-in for ``mmdet3d/ops/sst/sst_ops.py:150-181``.

The reference runs ``torch.unique(coors, dim=0)`` and ``torch_scatter`` (an un-vendored, unpinned
dependency: mean = sum / clamp(count, 1); max returns values).  Here both come from the same
sort-based plan as DynamicScatter (csrc/scatter.cu).  The plan of a ``unq_inv`` produced here is
cached on the tensor, so the 12 scatters of a SIR stack that reuse one ``unq_inv``
(``unique_once=True``, backbones/sir.py:70) sort only once.
"""
from __future__ import annotations

import traceback

import torch
from torch.autograd import Function

from . import _lib
from .voxel import _Plan, _reduce, _unique

_MODES = {"sum": 0, "mean": 1, "max": 2}


def _plan_from_inverse(unq_inv, M):
    N = unq_inv.numel()
    dev = unq_inv.device
    L = _lib.lib()
    p = _Plan()
    p.M = int(M)
    p.inverse = unq_inv.to(torch.int32).contiguous()
    p.order = torch.empty(N, dtype=torch.int32, device=dev)
    p.gstart = torch.empty(max(p.M, 1), dtype=torch.int32, device=dev)
    p.counts = torch.empty(max(p.M, 1), dtype=torch.int32, device=dev)
    ws = torch.empty(max(L.occb200_plan_workspace_bytes(N), 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = L.occb200_plan_from_inverse(p.inverse.data_ptr(), N, p.M, p.order.data_ptr(), p.gstart.data_ptr(),
                                         p.counts.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, "occb200_plan_from_inverse")
    p.uniq = None
    return p


class _PlanScatter(Function):
    """Segmented reduce with autograd to ``feat`` (mean/sum: gather; max: to the smallest arg index)."""

    @staticmethod
    def forward(ctx, feat, plan, mode):
        red = _MODES[mode]
        out, argmax = _reduce(feat, plan, red, want_argmax=(red == 2))
        ctx.plan, ctx.red, ctx.shape = plan, red, feat.shape
        if argmax is not None:
            ctx.save_for_backward(argmax)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        plan, red = ctx.plan, ctx.red
        N, Cc = ctx.shape
        grad = torch.empty((N, Cc), dtype=torch.float32, device=grad_out.device)
        argmax = ctx.saved_tensors[0] if red == 2 else None
        g = grad_out.contiguous().float()
        with torch.cuda.device(g.device):
            rc = _lib.lib().occb200_segment_reduce_backward(grad.data_ptr(), g.data_ptr(), None, None,
                                                            plan.inverse.data_ptr(), plan.counts.data_ptr(),
                                                            _lib.ptr(argmax), N, plan.M, Cc, red,
                                                            _lib.stream_ptr(g.device))
        _lib.check(rc, "occb200_segment_reduce_backward")
        return grad, None, None


def _unique_v2(coors):
    plan = _unique(coors.contiguous(), 0)
    inv = plan.inverse.long()
    inv._occb200_plan = plan
    return plan.uniq, inv, plan


def scatter_v2(feat, coors, mode, return_inv=True, min_points=0, unq_inv=None, new_coors=None):
    assert feat.size(0) == coors.size(0)
    if mode == 'avg':
        mode = 'mean'
    _lib.require_cuda(feat, coors)

    plan = None
    if unq_inv is None:
        new_coors, unq_inv, plan = _unique_v2(coors)
    else:
        assert new_coors is not None, \
            'please pass new_coors for interface consistency, caller: {}'.format(traceback.extract_stack()[-2][2])
        plan = getattr(unq_inv, "_occb200_plan", None)
        if plan is None or plan.M != new_coors.size(0):
            plan = _plan_from_inverse(unq_inv, new_coors.size(0))

    if min_points > 0:
        cnt_per_point = plan.counts.long()[unq_inv]
        valid_mask = cnt_per_point >= min_points
        feat = feat[valid_mask]
        coors = coors[valid_mask]
        new_coors, unq_inv, plan = _unique_v2(coors)

    if mode not in _MODES:
        raise NotImplementedError
    if feat.size(0) == 0:
        new_feat = feat.new_zeros((0, feat.size(1)))
    else:
        new_feat = _PlanScatter.apply(feat.float().contiguous(), plan, mode)

    if not return_inv:
        return new_feat, new_coors
    return new_feat, new_coors, unq_inv
