"""Sigmoid focal loss: host-side mirror of
``mmdet/ops/sigmoid_focal_loss/sigmoid_focal_loss.py`` on the C-ABI library.

``SigmoidFocalLossFunction`` / ``sigmoid_focal_loss`` / ``SigmoidFocalLoss`` keep the
reference's signatures (sigmoid_focal_loss.py:8-54).  ``sigmoid_focal_loss_sum`` is the fused
form of what ``FocalLoss`` does around the op in Python (focal_loss.py:28-42,
losses/utils.py:41-52): row weight and the sum in the same kernel.
"""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _capi


def _prep(input, target):
    if not input.is_cuda:
        # sigmoid_focal_loss.cpp:25: "SigmoidFocalLoss is not implemented on the CPU"
        raise NotImplementedError('SigmoidFocalLoss is not implemented on the CPU')
    if input.dim() != 2:
        raise RuntimeError('logits should be NxClass')            # sigmoid_focal_loss_cuda.cu:107
    x = input.detach().contiguous()
    t = target.detach().to(device=x.device, dtype=torch.long).contiguous()
    if t.numel() != x.shape[0]:
        raise RuntimeError('targets should have one label per row of logits')
    return x, t


class SigmoidFocalLossFunction(Function):

    @staticmethod
    def forward(ctx, input, target, gamma=2.0, alpha=0.25):
        ctx.save_for_backward(input, target)
        num_classes = input.shape[1]
        ctx.num_classes = num_classes
        ctx.gamma = gamma
        ctx.alpha = alpha
        x, t = _prep(input, target)
        loss = torch.empty_like(x)
        _capi.check(_capi.lib().kgdet_sigmoid_focal_loss_forward(
            x.data_ptr(), t.data_ptr(), x.shape[0], num_classes, float(gamma), float(alpha),
            loss.data_ptr(), _capi.dtype_code(x), _capi.stream_of(x)), 'kgdet_sigmoid_focal_loss_forward')
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, d_loss):
        input, target = ctx.saved_tensors
        x, t = _prep(input, target)
        d_loss = d_loss.to(x.dtype).contiguous()
        d_input = torch.empty_like(x)
        _capi.check(_capi.lib().kgdet_sigmoid_focal_loss_backward(
            x.data_ptr(), t.data_ptr(), d_loss.data_ptr(), x.shape[0], ctx.num_classes,
            float(ctx.gamma), float(ctx.alpha), d_input.data_ptr(), _capi.dtype_code(x),
            _capi.stream_of(x)), 'kgdet_sigmoid_focal_loss_backward')
        return d_input, None, None, None


sigmoid_focal_loss = SigmoidFocalLossFunction.apply


class SigmoidFocalLoss(nn.Module):
    """Mirror of sigmoid_focal_loss.py:39-54."""

    def __init__(self, gamma, alpha):
        super(SigmoidFocalLoss, self).__init__()
        self.gamma = gamma
        self.alpha = alpha

    def forward(self, logits, targets):
        assert logits.is_cuda
        loss = sigmoid_focal_loss(logits, targets, self.gamma, self.alpha)
        return loss.sum()

    def __repr__(self):
        tmpstr = self.__class__.__name__ + '(gamma={}, alpha={})'.format(self.gamma, self.alpha)
        return tmpstr


class _FocalSumFunction(Function):

    @staticmethod
    def forward(ctx, input, target, weight, gamma, alpha):
        x, t = _prep(input, target)
        w = None if weight is None else weight.detach().to(device=x.device, dtype=torch.float32).contiguous().view(-1)
        ctx.save_for_backward(x, t, w) if w is not None else ctx.save_for_backward(x, t)
        ctx.has_w = w is not None
        ctx.gamma, ctx.alpha = gamma, alpha
        out = torch.zeros((), dtype=torch.float32, device=x.device)
        _capi.check(_capi.lib().kgdet_sigmoid_focal_loss_sum_forward(
            x.data_ptr(), t.data_ptr(), _capi.ptr(w), x.shape[0], x.shape[1], float(gamma), float(alpha),
            out.data_ptr(), _capi.dtype_code(x), _capi.stream_of(x)), 'kgdet_sigmoid_focal_loss_sum_forward')
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        if ctx.has_w:
            x, t, w = ctx.saved_tensors
        else:
            (x, t), w = ctx.saved_tensors, None
        gs = g.detach().to(torch.float32).contiguous()
        d_input = torch.empty_like(x)
        _capi.check(_capi.lib().kgdet_sigmoid_focal_loss_sum_backward(
            x.data_ptr(), t.data_ptr(), _capi.ptr(w), gs.data_ptr(), x.shape[0], x.shape[1],
            float(ctx.gamma), float(ctx.alpha), d_input.data_ptr(), _capi.dtype_code(x),
            _capi.stream_of(x)), 'kgdet_sigmoid_focal_loss_sum_backward')
        return d_input, None, None, None, None


def sigmoid_focal_loss_sum(pred, target, weight=None, gamma=2.0, alpha=0.25):
    """sum_{m,c} focal(pred, target)[m, c] * weight[m] as one fused kernel (fp32 scalar)."""
    return _FocalSumFunction.apply(pred, target, weight, gamma, alpha)
