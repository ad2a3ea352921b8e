"""CPU stand-ins injected into kgdet_b200.head.KGDetHead -- TEST / BENCH-BASELINE INFRASTRUCTURE.

`make_cpu_head()` builds the head mirror with every CUDA operator replaced by its oracle
(DeformConv -> oracle/dcn_oracle.py, moment -> oracle/moment_oracle.py, NMS -> the reference's
nms_cpu.cpp from oracle/_ref when built, else oracle/nms_oracle.c).  Used by the CPU tests to check
the head's host logic against the golden fixtures and by bench.py's cpu_baseline / --impl reference
legs.  Never imported by the product package.
"""
import numpy as np
import torch

from oracle import build_ref, moment_oracle, nms_oracle
from tests import oracle_ops


def cpu_batched_nms_flags(dets, seg_offsets, max_seg_len, iou_thr, cmp_mode=1, score_thr=float('-inf')):
    """Per-segment greedy NMS on the host: the reference's per-class loop
    (mmdet/core/post_processing/bbox_nms_kp.py:38-52, incl. its score filter :39) with nms_cpu
    semantics ('>=').  Same signature as kgdet_b200.ops.batched_nms_flags."""
    ref = build_ref.load('nms_cpu')
    flags = torch.zeros(dets.shape[0], dtype=torch.uint8)
    d = dets.detach().float().contiguous()
    if seg_offsets is None:
        so = list(range(0, d.shape[0] + 1, max_seg_len))
    else:
        so = seg_offsets.tolist()
    for s in range(len(so) - 1):
        a, b = so[s], so[s + 1]
        if b <= a:
            continue
        rows = torch.nonzero(d[a:b, 4] > score_thr).squeeze(1)
        if rows.numel() == 0:
            continue
        seg = d[a:b][rows].contiguous()
        if ref is not None and cmp_mode == 1:
            keep = ref.nms(seg, float(iou_thr))
        else:
            keep = torch.from_numpy(nms_oracle.nms_keep(seg, iou_thr, cmp_mode))
        flags[a + rows[keep]] = 1
    return flags


def make_cpu_head(**kwargs):
    from kgdet_b200.head import KGDetHead
    return KGDetHead(deform_conv_cls=oracle_ops.DeformConv, moment_fn=moment_oracle.points2bbox_moment,
                     nms_flags_fn=cpu_batched_nms_flags, **kwargs)


def make_cpu_reppoints_head(variant, **kwargs):
    from kgdet_b200.head import RepPointsKpHead
    return RepPointsKpHead(variant, deform_conv_cls=oracle_ops.DeformConv,
                           moment_fn=moment_oracle.points2bbox_moment, nms_flags_fn=cpu_batched_nms_flags, **kwargs)
