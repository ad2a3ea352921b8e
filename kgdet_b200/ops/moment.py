"""Point set -> bbox moment transform as one fused op.

Fused twin of ``points2bbox(..., transform_method='moment')`` in the KGDet / RepPoints-Kp heads
(``reppoints_head_kp3rep_cas_1_assign_once.py:373-388``; same code at
``reppoints_head_kp_parallel.py:219-234``, ``reppoints_head_kp_serial.py:219-234``): mean and
unbiased std over the P points, ``exp(moment_transfer)`` scaling, bbox assembly -- forward and
backward (including the ``moment_mul``-scaled gradient of ``moment_transfer``) in one kernel each.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _capi


class _MomentFunction(Function):

    @staticmethod
    def forward(ctx, pts, moment_transfer, moment_mul, y_first):
        _capi.require_cuda(pts, 'points2bbox_moment')
        if pts.dim() < 2 or pts.shape[1] % 2 != 0:
            raise ValueError('pts must be [N, 2P, ...], got %s' % (tuple(pts.shape),))
        p = pts.detach().to(torch.float32).contiguous()
        mt = moment_transfer.detach().to(device=p.device, dtype=torch.float32).contiguous()
        N, P = p.shape[0], p.shape[1] // 2
        S = 1
        for d in p.shape[2:]:
            S *= d
        bbox = torch.empty((N, 4) + tuple(p.shape[2:]), dtype=torch.float32, device=p.device)
        _capi.check(_capi.lib().kgdet_points2bbox_moment_forward(
            p.data_ptr(), mt.data_ptr(), N, P, S, int(bool(y_first)), bbox.data_ptr(),
            _capi.stream_of(p)), 'kgdet_points2bbox_moment_forward')
        ctx.save_for_backward(p, mt)
        ctx.dims = (N, P, S, int(bool(y_first)), float(moment_mul))
        ctx.in_dtype = pts.dtype
        return bbox.to(pts.dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_bbox):
        p, mt = ctx.saved_tensors
        N, P, S, y_first, moment_mul = ctx.dims
        g = grad_bbox.detach().to(torch.float32).contiguous()
        grad_pts = torch.empty_like(p)
        grad_mt = torch.zeros(2, dtype=torch.float32, device=p.device)
        _capi.check(_capi.lib().kgdet_points2bbox_moment_backward(
            p.data_ptr(), mt.data_ptr(), g.data_ptr(), N, P, S, y_first, moment_mul,
            grad_pts.data_ptr(), grad_mt.data_ptr(), _capi.stream_of(p)),
            'kgdet_points2bbox_moment_backward')
        return grad_pts.to(ctx.in_dtype), grad_mt, None, None


def points2bbox_moment(pts, moment_transfer, moment_mul=0.01, y_first=True):
    """pts [N, 2P, H, W] (or [M, 2P]) -> bbox [N, 4, H, W] (x1, y1, x2, y2)."""
    return _MomentFunction.apply(pts, moment_transfer, moment_mul, y_first)
