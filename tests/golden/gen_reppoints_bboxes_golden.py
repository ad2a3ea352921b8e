"""Golden for get_bboxes of the RepPoints-Kp baseline heads (reppoints_head_kp_parallel.py:615-752 and the same
lines of reppoints_head_kp_serial.py; multiclass_nms_kp, core/post_processing/bbox_nms_kp.py:6-75): the UNCHANGED
reference class, built by the reference's own builder through tests/refshim.py (nms = the reference's nms_cpu.cpp
compiled unmodified), on five small levels of a 256x320 image, batch 2.

    python -m tests.golden.gen_reppoints_bboxes_golden     # writes tests/golden/reppoints_bboxes.npz

Inputs are regenerated from the seed (`make_case`); the fixture stores the detections (boxes, scores, labels of all
100 per image; the 882 keypoint values of the 25 best, a float64 row sum of every one).  Levels 0 and 1 have more
positions than nms_pre (300), so the per-level top-k runs; boxes and keypoints leave the image, so the clamps run --
including this head's own keypoint clamp, which differs from the KGDet head's (PAR:721-722 index the KEYPOINT
axis with 0::3 / 1::3: keypoints 0, 3, 6... are clamped to the image WIDTH in x, y and visibility, keypoints
1, 4, 7... to the HEIGHT, keypoints 2, 5, 8... not at all).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SIZES = [(32, 40), (16, 20), (8, 10), (4, 5), (2, 3)]
IMG = (256, 320)
NMS_PRE = 300
SCALE = 1.6                     # scale_factor of the rescale=True case


def make_case(seed=99, batch=2):
    """cls logits, refined keypoint offsets [B,588,h,w] and refined point sets [B,18,h,w] per level."""
    g = torch.Generator().manual_seed(seed)
    cls, kpt, rep = [], [], []
    for h, w in SIZES:
        sc = torch.rand(batch, 13, h, w, generator=g) ** 4
        sc = sc.clamp(1e-4, 1 - 1e-4)
        cls.append(torch.log(sc / (1 - sc)))
        kpt.append(torch.randn(batch, 588, h, w, generator=g) * 3)
        rep.append(torch.randn(batch, 18, h, w, generator=g) * 4)
    return cls, kpt, rep


def main():
    from tests import refshim
    refshim.install('oracle')
    cls, kpt, rep = make_case()
    out = {}
    for variant in ('parallel', 'serial'):
        head, cfg = refshim.build_head('reppoints_moment_%s_r50_fpn_1x-deepfashion2.py' % variant)
        head.eval()
        with torch.no_grad():
            head.moment_transfer.copy_(torch.tensor([0.25, -0.15]))
        tc = refshim.AttrDict(cfg['test_cfg'])
        tc['nms_pre'] = NMS_PRE
        metas = [dict(img_shape=IMG + (3,), scale_factor=1.0)] * cls[0].shape[0]
        with torch.no_grad():
            res = head.get_bboxes([c.clone() for c in cls], [k.clone() for k in kpt], [k.clone() for k in kpt],
                                  [r.clone() for r in rep], [r.clone() for r in rep], metas, tc, rescale=False)
        # the same call with rescale=True (PAR:732-737: boxes / keypoints divided by the scale factor BEFORE the NMS)
        metas_rs = [dict(img_shape=IMG + (3,), scale_factor=SCALE)] * cls[0].shape[0]
        with torch.no_grad():
            res_rs = head.get_bboxes([c.clone() for c in cls], [k.clone() for k in kpt], [k.clone() for k in kpt],
                                     [r.clone() for r in rep], [r.clone() for r in rep], metas_rs, tc, rescale=True)
        for tag, rr in (('', res), ('rs_', res_rs)):
            for i, (d, l, k) in enumerate(rr):
                out['%s_%sdets_%d' % (variant, tag, i)] = d.numpy()
                out['%s_%slabels_%d' % (variant, tag, i)] = l.numpy()
                kk = k.reshape(d.shape[0], -1)
                out['%s_%skpts_head_%d' % (variant, tag, i)] = kk[:25].numpy()             # full rows of the 25 best
                out['%s_%skpts_rowsum_%d' % (variant, tag, i)] = kk.double().sum(1).numpy()  # every row, as a checksum
                print(variant, tag, i, tuple(d.shape), int(l.min()), int(l.max()), flush=True)
        out['%s_score_thr' % variant] = np.float32(tc['score_thr'])
        out['%s_iou_thr' % variant] = np.float32(tc['nms']['iou_thr'])
        out['%s_max_per_img' % variant] = np.int64(tc['max_per_img'])
    np.savez_compressed(os.path.join(HERE, 'reppoints_bboxes.npz'), **out)


if __name__ == '__main__':
    main()
