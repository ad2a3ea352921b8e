"""Target assignment + the nine training losses of the KGDet head as CUDA kernels.

Mirror of what the reference runs on the host side of PyTorch for every image and ground-truth box --
``PointAssigner.assign`` (mmdet/core/bbox/assigners/point_assigner.py:23-116), ``point_target_kp``
(mmdet/core/anchor/point_target_kp.py:7-169) and ``RepPointsHeadKp3RepCas1AssignOnce.loss / loss_single``
(reppoints_head_kp3rep_cas_1_assign_once.py:581-768) -- as three launches of ``libkgdet_b200.so``
(csrc/point_loss.cu): ``kgdet_point_assign``, ``kgdet_point_losses_forward`` and, in backward,
``kgdet_point_losses_backward``.  No host synchronisation, static shapes: CUDA-graph capturable.
Ground truth is padded (``kgdet_b200.targets.pad_ground_truth``).
"""
import ctypes

import torch
from torch.autograd import Function

from . import _capi

LOSS_NAMES = ('loss_cls_1', 'loss_cls_2', 'loss_cls_3', 'loss_bbox_1', 'loss_bbox_2', 'loss_bbox_3',
              'loss_kpt_1', 'loss_kpt_2', 'loss_kpt_3')


def point_assign(gt_bboxes, gt_valid, gt_keypoints, map_hw, stride, pos_num=25):
    """-> assigned [B, P] int32 (0 = background, g + 1 = box g), avg_factor [1] fp32 (sum over images of
    max(#positives, 1)), num_visible [B, G] fp32.  One point level of map_hw = (H, W) points at `stride`."""
    lib = _capi.lib()
    _capi.require_cuda(gt_bboxes, 'point_assign')
    B, G = gt_valid.shape
    H, W = map_hw
    boxes = gt_bboxes.detach().float().contiguous()
    valid = gt_valid.detach().to(torch.uint8).contiguous()
    kps = gt_keypoints.detach().float().contiguous()
    dev = boxes.device
    assigned = torch.empty((B, H * W), dtype=torch.int32, device=dev)
    avg = torch.empty(1, dtype=torch.float32, device=dev)
    nvis = torch.empty((B, G), dtype=torch.float32, device=dev)
    scratch = torch.empty(max(int(lib.kgdet_point_assign_scratch_bytes(B, H, W)), 16), dtype=torch.uint8, device=dev)
    _capi.check(lib.kgdet_point_assign(boxes.data_ptr(), valid.data_ptr(), kps.data_ptr(), B, G, kps.shape[2], H, W,
                                       float(stride), int(pos_num), assigned.data_ptr(), avg.data_ptr(), nvis.data_ptr(),
                                       scratch.data_ptr(), _capi.stream_of(boxes)), 'kgdet_point_assign')
    return assigned, avg, nvis


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


class _PointLosses(Function):
    @staticmethod
    def forward(ctx, cfg, assigned, avg, nvis, gt_boxes, gt_labels, gt_kps, *outs):
        lib = _capi.lib()
        assert len(outs) == 9
        outs_c = [o.detach().float().contiguous() for o in outs]
        B, NC, H, W = outs_c[0].shape
        K = outs_c[3].shape[1] // 2
        G = gt_boxes.shape[1]
        stride, pbs, weights, gamma, alpha, beta = cfg
        lw = (ctypes.c_float * 9)(*weights)
        losses = torch.empty(9, dtype=torch.float32, device=outs_c[0].device)
        _capi.check(lib.kgdet_point_losses_forward(
            ctypes.cast(_ptr_array(outs_c), ctypes.c_void_p), assigned.data_ptr(), gt_boxes.data_ptr(), gt_labels.data_ptr(),
            gt_kps.data_ptr(), avg.data_ptr(), nvis.data_ptr(), B, G, H, W, NC, K, float(stride), float(pbs),
            ctypes.cast(lw, ctypes.c_void_p), float(gamma), float(alpha), float(beta), losses.data_ptr(),
            _capi.stream_of(losses)), 'kgdet_point_losses_forward')
        ctx.cfg = cfg
        ctx.dtypes = [o.dtype for o in outs]
        ctx.save_for_backward(assigned, avg, nvis, gt_boxes, gt_labels, gt_kps, *outs_c)
        return losses

    @staticmethod
    def backward(ctx, grad_losses):
        lib = _capi.lib()
        assigned, avg, nvis, gt_boxes, gt_labels, gt_kps = ctx.saved_tensors[:6]
        outs_c = ctx.saved_tensors[6:]
        B, NC, H, W = outs_c[0].shape
        K = outs_c[3].shape[1] // 2
        G = gt_boxes.shape[1]
        stride, pbs, weights, gamma, alpha, beta = ctx.cfg
        lw = (ctypes.c_float * 9)(*weights)
        gl = grad_losses.detach().float().contiguous()
        grads = [torch.empty_like(o) if ctx.needs_input_grad[7 + i] else None for i, o in enumerate(outs_c)]
        _capi.check(lib.kgdet_point_losses_backward(
            ctypes.cast(_ptr_array(outs_c), ctypes.c_void_p), assigned.data_ptr(), gt_boxes.data_ptr(), gt_labels.data_ptr(),
            gt_kps.data_ptr(), avg.data_ptr(), nvis.data_ptr(), gl.data_ptr(), B, G, H, W, NC, K, float(stride), float(pbs),
            ctypes.cast(lw, ctypes.c_void_p), float(gamma), float(alpha), float(beta),
            ctypes.cast(_ptr_array(grads), ctypes.c_void_p), _capi.stream_of(gl)), 'kgdet_point_losses_backward')
        grads = [None if g is None else g.to(dt) for g, dt in zip(grads, ctx.dtypes)]
        return (None,) * 7 + tuple(grads)


def kgdet_point_losses(outs, gt_bboxes, gt_labels, gt_keypoints, gt_valid, stride, assigner_scale=4, pos_num=25,
                       point_base_scale=4, cls_weights=(0.5, 0.5, 1.0), bbox_weights=(0.5, 0.5, 1.0),
                       kpt_weights=(0.5, 0.5, 1.0), gamma=2.0, alpha=0.25, beta=1.0 / 9.0, return_targets=False):
    """The nine losses of KP3.loss for the single KGDet level from the 9-tuple of forward_single and padded ground
    truth: one assignment launch + one loss launch (+ one launch in backward).  Returns the dict of the nine
    scalars (views of one [9] tensor), or (dict, (assigned, avg_factor)) with `return_targets`.
    `assigner_scale` only selects the pyramid level of a box (point_assigner.py:62-64) -- with one level every box
    lands on it, so it does not enter the arithmetic."""
    _capi.require_cuda(outs[0], 'kgdet_point_losses')
    H, W = outs[0].shape[-2:]
    boxes = gt_bboxes.detach().float().contiguous()
    labels = gt_labels.detach().long().contiguous()
    kps = gt_keypoints.detach().float().contiguous()
    assigned, avg, nvis = point_assign(boxes, gt_valid, kps, (H, W), stride, pos_num)
    cfg = (float(stride), float(point_base_scale), tuple(cls_weights) + tuple(bbox_weights) + tuple(kpt_weights),
           float(gamma), float(alpha), float(beta))
    losses = _PointLosses.apply(cfg, assigned, avg, nvis, boxes, labels, kps, *outs)
    d = {n: losses[i] for i, n in enumerate(LOSS_NAMES)}
    return (d, (assigned, avg)) if return_targets else d
