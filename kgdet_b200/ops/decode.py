"""Post-head decode kernels around the batched NMS: candidate selection, box
decode + dense NMS input, and the final gather / keypoint decode of get_bboxes_single + multiclass_nms_kp
(reppoints_head_kp3rep_cas_1_assign_once.py:843-903, core/post_processing/bbox_nms_kp.py:6-75), one head
level, batched over images, static shapes, no host synchronisation."""
import torch

from . import _capi


def bbox_select(scores, apply_sigmoid, n):
    """scores [B, C, H, W] fp32 (logits if apply_sigmoid) -> order [B, n] int32: positions by descending
    max-over-classes score (topk(nms_pre) of KP3:863-874), identity when n == H*W."""
    lib = _capi.lib()
    _capi.require_cuda(scores, 'bbox_select')
    assert scores.dtype == torch.float32 and scores.is_contiguous()
    B, C, H, W = scores.shape
    order = torch.empty((B, n), dtype=torch.int32, device=scores.device)
    ws_bytes = int(lib.kgdet_bbox_select_workspace_bytes(B, H * W, n))
    if ws_bytes:              # a large level (more than 4096 positions): radix select, needs scratch
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=scores.device)
        _capi.check(lib.kgdet_bbox_select_ws(scores.data_ptr(), int(bool(apply_sigmoid)), B, C, H * W, n,
                                             order.data_ptr(), ws.data_ptr(), ws_bytes, _capi.stream_of(scores)),
                    'kgdet_bbox_select_ws')
        return order
    _capi.check(lib.kgdet_bbox_select(scores.data_ptr(), int(bool(apply_sigmoid)), B, C, H * W, n, order.data_ptr(),
                                      _capi.stream_of(scores)), 'kgdet_bbox_select')
    return order


def bbox_decode(scores, apply_sigmoid, bbox, order, img_wh, stride):
    """-> boxes [B, n, 4] (decoded, clamped) and dets [B, C, n, 5] (dense NMS input, one segment per (b, c))."""
    lib = _capi.lib()
    assert bbox.dtype == torch.float32 and bbox.is_contiguous() and img_wh.dtype == torch.float32
    B, C, H, W = scores.shape
    n = order.shape[1]
    boxes = torch.empty((B, n, 4), dtype=torch.float32, device=scores.device)
    dets = torch.empty((B, C, n, 5), dtype=torch.float32, device=scores.device)
    _capi.check(lib.kgdet_bbox_decode(scores.data_ptr(), int(bool(apply_sigmoid)), bbox.data_ptr(), order.data_ptr(),
                                      img_wh.data_ptr(), float(stride), W, B, C, H * W, n, boxes.data_ptr(),
                                      dets.data_ptr(), _capi.stream_of(scores)), 'kgdet_bbox_decode')
    return boxes, dets


def bbox_finalize(boxes, keypts, order, top_i, top_s, img_wh, stride, map_hw):
    """-> (dets [B, k, 5], labels [B, k] int64 with -1 = empty, kpts [B, k, P*3]) for the k selected
    (class, candidate) pairs top_i = class * n + candidate; keypoints are decoded only for those."""
    lib = _capi.lib()
    assert keypts.dtype == torch.float32 and keypts.is_contiguous() and top_i.dtype == torch.int64
    B, n = order.shape
    k = top_i.shape[1]
    H, W = map_hw
    P = keypts.shape[1] // 2
    top_i, top_s = top_i.contiguous(), top_s.contiguous()
    out_dets = torch.empty((B, k, 5), dtype=torch.float32, device=boxes.device)
    out_labels = torch.empty((B, k), dtype=torch.int64, device=boxes.device)
    out_kpts = torch.empty((B, k, P * 3), dtype=torch.float32, device=boxes.device)
    _capi.check(lib.kgdet_bbox_finalize(boxes.data_ptr(), keypts.data_ptr(), order.data_ptr(), top_i.data_ptr(),
                                        top_s.data_ptr(), img_wh.data_ptr(), float(stride), W, B, H * W, n, k, P,
                                        out_dets.data_ptr(), out_labels.data_ptr(), out_kpts.data_ptr(),
                                        _capi.stream_of(boxes)), 'kgdet_bbox_finalize')
    return out_dets, out_labels, out_kpts


def topk_flagged(dets, flags, k):
    """Top-k by score over the rows of `dets` [B, L, 5] whose `flags` [B, L] (uint8) are set: -> (top_s [B, k]
    descending, -1 for empty slots; top_i [B, k] int64).  L <= 16384."""
    lib = _capi.lib()
    assert dets.dtype == torch.float32 and dets.is_contiguous() and flags.dtype == torch.uint8 and flags.is_contiguous()
    B, L = dets.shape[0], dets.shape[1]
    top_s = torch.empty((B, k), dtype=torch.float32, device=dets.device)
    top_i = torch.empty((B, k), dtype=torch.int64, device=dets.device)
    _capi.check(lib.kgdet_topk_flagged(dets.data_ptr(), flags.data_ptr(), B, L, k, top_s.data_ptr(), top_i.data_ptr(),
                                       _capi.stream_of(dets)), 'kgdet_topk_flagged')
    return top_s, top_i
