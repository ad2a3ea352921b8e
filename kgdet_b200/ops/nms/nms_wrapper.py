"""NMS: host-side mirror of ``mmdet/ops/nms/nms_wrapper.py`` on the C-ABI library.

``nms(dets, iou_thr, device_id=None) -> (dets[inds], inds)`` keeps the reference's contract
(nms_wrapper.py:8-49): Tensor or ndarray in, same type out, empty in -> empty out (:39-40),
``inds`` are ascending original indices (nms_kernel.cu:127-130 / nms_cpu.cpp:58).

The reference has two back-ends that disagree at IoU == thr: ``nms_cuda`` suppresses at
``IoU > thr`` (nms_kernel.cu:60), ``nms_cpu`` at ``IoU >= thr`` (nms_cpu.cpp:55).  Both
semantics run on the GPU here: CUDA tensors use '>' like ``nms_cuda``; CPU tensors / arrays
are staged to the device and use '>=' like ``nms_cpu`` (there is no host implementation).

``batched_nms_flags`` is the one-launch replacement for the per-class loop of
``multiclass_nms_kp`` (mmdet/core/post_processing/bbox_nms_kp.py:38-52).
"""
import numpy as np
import torch

from .. import _capi


def _nms_keep_cuda(dets_th, iou_thr, cmp_mode):
    """dets_th: CUDA tensor [n, >=5]; returns LongTensor of kept indices (ascending)."""
    lib = _capi.lib()
    n = dets_th.shape[0]
    d = dets_th.detach()
    if d.dtype != torch.float32 or d.shape[1] != 5 or not d.is_contiguous():
        d = d[:, :5].to(torch.float32).contiguous()   # nms_cuda is float-only (nms_kernel.cu:71)
    keep = torch.empty(n, dtype=torch.long, device=d.device)
    num = torch.empty(1, dtype=torch.int32, device=d.device)
    ws = _capi.workspace(lib.kgdet_nms_workspace_bytes(n), d)
    _capi.check(lib.kgdet_nms(d.data_ptr(), n, float(iou_thr), cmp_mode, keep.data_ptr(),
                              num.data_ptr(), ws.data_ptr(), ws.numel(), _capi.stream_of(d)),
                'kgdet_nms')
    return keep[:int(num.item())]


def nms(dets, iou_thr, device_id=None):
    """Drop-in for mmdet.ops.nms.nms (nms_wrapper.py:8-49)."""
    if isinstance(dets, torch.Tensor):
        is_numpy = False
        dets_th = dets
    elif isinstance(dets, np.ndarray):
        is_numpy = True
        device = 'cpu' if device_id is None else 'cuda:{}'.format(device_id)
        dets_th = torch.from_numpy(dets).to(device)
    else:
        raise TypeError('dets must be either a Tensor or numpy array, but got {}'.format(type(dets)))

    if dets_th.shape[0] == 0:
        inds = dets_th.new_zeros(0, dtype=torch.long)
    elif dets_th.is_cuda:
        inds = _nms_keep_cuda(dets_th, iou_thr, _capi.NMS_GT)
    else:
        if not torch.cuda.is_available():
            raise RuntimeError('kgdet_b200.nms: no CUDA device; there is no CPU implementation')
        inds = _nms_keep_cuda(dets_th.cuda(), iou_thr, _capi.NMS_GE).cpu()

    if is_numpy:
        inds = inds.cpu().numpy()
    return dets[inds, :], inds


def batched_nms_flags(dets, seg_offsets, max_seg_len, iou_thr, cmp_mode=_capi.NMS_GT,
                      score_thr=float('-inf')):
    """One launch of greedy NMS over many independent segments (one per (image, class)).

    dets [total,5] fp32 CUDA.  seg_offsets: int32 CUDA [nseg+1], or None for dense mode (total is a
    multiple of max_seg_len and every segment has exactly max_seg_len rows).  Rows with
    score <= score_thr are treated as absent.  Returns a uint8 keep flag per row.  No host sync.
    """
    lib = _capi.lib()
    _capi.require_cuda(dets, 'batched_nms_flags')
    d = dets.detach().to(torch.float32).contiguous()
    total = d.shape[0]
    if seg_offsets is None:
        assert max_seg_len > 0 and total % max_seg_len == 0
        so_ptr, nseg = None, total // max_seg_len
    else:
        so = seg_offsets.to(device=d.device, dtype=torch.int32).contiguous()
        so_ptr, nseg = so.data_ptr(), so.numel() - 1
    flags = torch.zeros(total, dtype=torch.uint8, device=d.device)
    if total == 0 or nseg <= 0:
        return flags
    _capi.check(lib.kgdet_nms_batched(d.data_ptr(), so_ptr, nseg, total, int(max_seg_len), float(iou_thr),
                                      float(score_thr), cmp_mode, flags.data_ptr(), None, 0,
                                      _capi.stream_of(d)), 'kgdet_nms_batched')
    return flags


def soft_nms(dets, iou_thr, method='linear', sigma=0.5, min_score=1e-3):
    """Out of scope: no KGDet/RepPoints-Kp config selects soft_nms (test_cfg.nms.type == 'nms');
    the reference's version is a CPU Cython loop (nms/src/soft_nms_cpu.pyx)."""
    raise NotImplementedError('soft_nms is outside the kgdet_b200 hot path (SURVEY.md section 2, row 2)')
