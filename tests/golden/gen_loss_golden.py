"""Golden fixture for the target assignment + the nine KGDet losses: runs the UNCHANGED reference
`RepPointsHeadKp3RepCas1AssignOnce.loss` (KP3:670-768 -> point_target_kp -> PointAssigner -> loss_single) on the CPU
through tests/refshim.py (DeformConv / focal loss served by the oracles, nothing else replaced).

    python -m tests.golden.gen_loss_golden        # writes tests/golden/kgdet_loss.npz

Stored: the targets `point_target_kp` produced and the nine loss values (inputs are regenerated from a seed by
make_case()).  Ground-truth boxes are placed off the grid's symmetry axes so that no two points are equidistant from a
box centre (the reference's `topk` is unspecified among ties).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def make_case(seed=21, B=3, H=13, W=21):
    g = torch.Generator().manual_seed(seed)
    outs = [torch.randn(B, 13, H, W, generator=g) for _ in range(3)] + \
           [torch.randn(B, 588, H, W, generator=g) * 2 for _ in range(3)] + \
           [torch.randn(B, 4, H, W, generator=g) * 2 for _ in range(3)]
    # image 0: two overlapping boxes competing for points; image 1: one box; image 2: three boxes, one tiny
    gt_bboxes = [torch.tensor([[31.3, 40.7, 301.9, 280.1], [150.2, 21.9, 433.4, 310.6]]),
                 torch.tensor([[63.1, 52.3, 500.7, 300.9]]),
                 torch.tensor([[10.7, 11.3, 211.9, 190.1], [333.3, 100.1, 610.9, 380.7], [401.1, 205.3, 412.9, 219.9]])]
    gt_labels = [torch.tensor([3, 7]), torch.tensor([12]), torch.tensor([1, 13, 5])]
    gt_kps = []
    for b in gt_bboxes:
        k = torch.zeros(b.shape[0], 294, 3)
        lo = int(torch.randint(0, 250, (1,), generator=g))
        k[:, lo:lo + 30, 0] = torch.rand(b.shape[0], 30, generator=g) * 600
        k[:, lo:lo + 30, 1] = torch.rand(b.shape[0], 30, generator=g) * 400
        k[:, lo:lo + 30, 2] = (torch.rand(b.shape[0], 30, generator=g) > 0.3).float() * 2
        gt_kps.append(k)
    return outs, gt_bboxes, gt_labels, gt_kps, (H * 32, W * 32)


def main():
    from tests import refshim
    from tests.golden.gen_golden import fill_state_dict
    refshim.install('oracle')
    head, cfg = refshim.build_head('kgdet_moment_r50_fpn_1x-demo.py')
    head.load_state_dict(fill_state_dict(head.state_dict()), strict=True)
    head.train()
    outs, gt_bboxes, gt_labels, gt_kps, (ih, iw) = make_case()
    B = outs[0].shape[0]
    metas = [dict(img_shape=(ih, iw, 3), pad_shape=(ih, iw, 3), scale_factor=1.0, flip=False)] * B
    tc = refshim.AttrDict(uniform=refshim.AttrDict(cfg['train_cfg']['uniform']))
    losses = head.loss(*[[o] for o in outs], [b.clone() for b in gt_bboxes], gt_labels, gt_kps, metas, tc)
    # the targets themselves, from the reference's own point_target_kp
    from mmdet.core import point_target_kp
    centers, flags = head.get_points([outs[0].shape[-2:]], metas)
    tg = point_target_kp(centers, flags, [b.clone() for b in gt_bboxes], gt_kps, metas, tc.uniform,
                         gt_bboxes_ignore_list=None, gt_labels_list=gt_labels, label_channels=13, sampling=False)
    (labels, label_w, bbox_gt, _, bbox_w, kpt_gt, kpt_w, num_pos, num_neg) = tg
    # the head outputs and the ground truth are NOT stored: make_case() regenerates them from its seed (same torch
    # build on both machines, like fill_state_dict); a checksum guards against generator drift
    out = dict(out_checksum=np.float64(sum(float(o.double().abs().sum()) for o in outs)))
    out.update(labels=labels[0].numpy(), label_weights=label_w[0].numpy(), bbox_gt=bbox_gt[0].numpy(),
               bbox_weights=bbox_w[0].numpy(), kpt_gt=kpt_gt[0].numpy(), kpt_weights=kpt_w[0].numpy(),
               num_total_pos=np.int64(num_pos), num_images=np.int64(B))
    for k, v in losses.items():
        out[k] = np.float64(float(v[0]))
    np.savez_compressed(os.path.join(HERE, 'kgdet_loss.npz'), **out)
    print({k: float(v[0]) for k, v in losses.items()}, 'num_total_pos', num_pos)


if __name__ == '__main__':
    main()
