"""An `mmdet.ops`-shaped module backed by the CPU oracles -- TEST INFRASTRUCTURE ONLY.

Mounted by tests/refshim.py (ops='oracle') so the UNCHANGED reference heads can run on the CPU
of this container; used to generate the golden fixtures under tests/golden/ and as the CPU
reference arm of bench.py.  DeformConv keeps the reference module's parameter set
(mmdet/ops/dcn/deform_conv.py:190-236) and calls oracle.dcn_oracle (autograd provides backward).
"""
import math
import sys
import types

import numpy as np
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from oracle import build_ref, dcn_oracle, nms_oracle


def deform_conv(input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                im2col_step=64):
    return dcn_oracle.deform_conv_forward(input, offset, weight, stride, padding, dilation, groups,
                                          deformable_groups)


def modulated_deform_conv(input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1,
                          groups=1, deformable_groups=1):
    return dcn_oracle.deform_conv_forward(input, offset, weight, stride, padding, dilation, groups,
                                          deformable_groups, mask=mask, bias=bias)


class DeformConv(nn.Module):

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=False):
        super().__init__()
        assert not bias
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride, self.padding, self.dilation = _pair(stride), _pair(padding), _pair(dilation)
        self.groups, self.deformable_groups = groups, deformable_groups
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        n = in_channels
        for k in self.kernel_size:
            n *= k
        self.weight.data.uniform_(-1. / math.sqrt(n), 1. / math.sqrt(n))

    def forward(self, x, offset):
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                           self.deformable_groups)


def nms(dets, iou_thr, device_id=None):
    """nms_wrapper.nms on the CPU: the reference's own nms_cpu.cpp when oracle/_ref has it, else the
    C restatement (same results; tests/test_oracle_cpu.py pins one to the other)."""
    ref = build_ref.load('nms_cpu')
    if isinstance(dets, np.ndarray):
        inds = nms_oracle.nms_keep(dets, iou_thr, 1)
        return dets[inds, :], inds
    if dets.shape[0] == 0:
        inds = dets.new_zeros(0, dtype=torch.long)
    elif ref is not None:
        inds = ref.nms(dets.detach().cpu().float().contiguous(), float(iou_thr))
    else:
        inds = torch.from_numpy(nms_oracle.nms_keep(dets, iou_thr, 1))
    return dets[inds, :], inds


def sigmoid_focal_loss(pred, target, gamma=2.0, alpha=0.25):
    """Elementwise [M,C] focal loss with the CUDA op's label convention (0 = background, t = c+1),
    written with torch ops so autograd works; same formula as the reference's debug twin
    py_sigmoid_focal_loss (mmdet/models/losses/focal_loss.py:10-25) on one-hot targets."""
    M, C = pred.shape
    onehot = torch.zeros_like(pred)
    pos = (target > 0) & (target <= C)
    onehot[pos, target[pos] - 1] = 1
    valid = (target >= 0).to(pred.dtype).view(-1, 1)
    p = pred.sigmoid()
    pt = (1 - p) * onehot + p * (1 - onehot)
    fw = (alpha * onehot + (1 - alpha) * (1 - onehot)) * pt.pow(gamma)
    return torch.nn.functional.binary_cross_entropy_with_logits(pred, onehot, reduction='none') * fw * valid


class SigmoidFocalLoss(nn.Module):
    def __init__(self, gamma, alpha):
        super().__init__()
        self.gamma, self.alpha = gamma, alpha

    def forward(self, logits, targets):
        return sigmoid_focal_loss(logits, targets, self.gamma, self.alpha).sum()


def _na(name):
    def fn(*a, **k):
        raise NotImplementedError(name)
    return fn


def mount():
    ops = types.ModuleType('mmdet.ops')
    names = dict(DeformConv=DeformConv, deform_conv=deform_conv, modulated_deform_conv=modulated_deform_conv,
                 nms=nms, sigmoid_focal_loss=sigmoid_focal_loss, SigmoidFocalLoss=SigmoidFocalLoss)
    for n in ('soft_nms', 'RoIAlign', 'roi_align', 'RoIPool', 'roi_pool', 'DeformConvPack', 'DeformRoIPooling',
              'DeformRoIPoolingPack', 'ModulatedDeformRoIPoolingPack', 'ModulatedDeformConv',
              'ModulatedDeformConvPack', 'deform_roi_pooling', 'MaskedConv2d', 'ContextBlock'):
        names[n] = type(n, (nn.Module,), {'__init__': _na(n)}) if n[0].isupper() else _na(n)
    ops.__dict__.update(names)
    wrapper = types.ModuleType('mmdet.ops.nms.nms_wrapper')
    wrapper.nms = nms
    wrapper.soft_nms = names['soft_nms']
    nms_pkg = types.ModuleType('mmdet.ops.nms')
    nms_pkg.nms, nms_pkg.soft_nms, nms_pkg.nms_wrapper = nms, names['soft_nms'], wrapper
    sys.modules['mmdet.ops'] = ops
    sys.modules['mmdet.ops.nms'] = nms_pkg
    sys.modules['mmdet.ops.nms.nms_wrapper'] = wrapper
    ops.nms_pkg = nms_pkg
    mm = sys.modules.get('mmdet')
    if mm is not None:
        mm.ops = ops
    return ops
