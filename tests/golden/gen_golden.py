"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE'S OWN CODE.

Run in the build container (the reference tree is not on the GPU box):

    python -m tests.golden.gen_golden

What executes here is the unchanged reference Python (`/root/reference/mmdetection/mmdet`,
imported through tests/refshim.py) plus the reference's own `nms_cpu.cpp` compiled unmodified
into oracle/_ref.  Only DeformConv -- CUDA-only in the reference (dcn/deform_conv.py:44-45) --
is served by oracle/dcn_oracle.py, which is pinned separately against torchvision (CPU) and the
reference CUDA kernels (GPU box).  Fixtures:

  nms.npz          dets -> keep indices of the reference nms_cpu.cpp
  focal_loss.npz   logits/targets -> py_sigmoid_focal_loss (mmdet/models/losses/focal_loss.py:10-25)
  moment.npz       point sets -> RepPointsHeadKp3RepCas1AssignOnce.points2bbox (KP3:342-391)
  head_p7.npz      seeded weights recipe + x[2,256,7,11] -> the 9 outputs of forward_single (KP3:412-446)
  head_p5.npz      same at [1,256,25,42]: checksums and a strided sample of each output
  get_bboxes.npz   stage-3 maps + synthetic scores -> get_bboxes (KP3:770-914, multiclass_nms_kp)
  reppoints_{parallel,serial}.npz  the two RepPoints-Kp baseline heads on three small levels (PAR/SER:292-341)

Weights are not stored: `fill_state_dict` regenerates them from a seed (same torch build on both
machines).  Inputs are stored explicitly.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests._data import random_boxes  # noqa: E402


def fill_state_dict(sd, seed=1234, dcn_gain=4.0):
    """Deterministic, non-degenerate weights for a KGDet head state dict.

    The head's own init (std 0.01) makes every learned offset ~0, which would leave the bilinear
    sampling and the window tests unexercised; here conv weights are N(0, (g/sqrt(fan_in))^2) so
    that predicted points spread over several pixels.  Keys are visited in sorted order.
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(sd.keys()):
        v = sd[k]
        if k == 'moment_transfer':
            out[k] = torch.tensor([0.25, -0.15])
        elif k.endswith('gn.weight'):
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith('.bias'):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif v.dim() == 4:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            gain = dcn_gain if 'reppts_out' in k or 'keypts_out' in k else 1.4
            out[k] = torch.randn(v.shape, generator=g) * (gain / fan_in ** 0.5)
        else:
            out[k] = torch.randn(v.shape, generator=g)
    return out


def strided_sample(t, n=4096):
    f = t.detach().reshape(-1)
    step = max(f.numel() // n, 1)
    return f[::step][:n].numpy().copy()


def main():
    from oracle import build_ref
    from tests import refshim
    refshim.install('oracle')
    ref_nms = build_ref.load('nms_cpu')
    assert ref_nms is not None, 'run `python -m oracle.build_ref --cpu-only` first'

    # ---- NMS ------------------------------------------------------------------------------
    nms = {}
    for name, (n, clustered, seed) in dict(a=(1000, False, 0), b=(1000, True, 1), c=(3350, True, 2),
                                           d=(65, True, 3), e=(1, False, 4)).items():
        dets = random_boxes(n, seed=seed, clustered=clustered)
        nms['dets_' + name] = dets.numpy()
        nms['keep_' + name] = ref_nms.nms(dets, 0.5).numpy()
    np.savez_compressed(os.path.join(HERE, 'nms.npz'), **nms)

    # ---- focal loss -------------------------------------------------------------------------
    from mmdet.models.losses.focal_loss import py_sigmoid_focal_loss
    g = torch.Generator().manual_seed(7)
    logits = torch.randn(300, 13, generator=g) * 3
    targets = torch.randint(0, 14, (300,), generator=g)
    onehot = torch.zeros(300, 13)
    pos = targets > 0
    onehot[pos, targets[pos] - 1] = 1
    loss = py_sigmoid_focal_loss(logits, onehot, None, 2.0, 0.25, 'none')
    lg = logits.clone().requires_grad_()
    w = torch.rand(300, generator=g)
    red = py_sigmoid_focal_loss(lg, onehot, w.view(-1, 1), 2.0, 0.25, 'mean', avg_factor=17.0)
    red.backward()
    np.savez_compressed(os.path.join(HERE, 'focal_loss.npz'), logits=logits.numpy(), targets=targets.numpy(),
                        loss=loss.numpy(), weight=w.numpy(), reduced=np.float64(red.item()),
                        grad=lg.grad.numpy())

    # ---- head ---------------------------------------------------------------------------------
    head, cfg = refshim.build_head('kgdet_moment_r50_fpn_1x-demo.py')
    head.load_state_dict(fill_state_dict(head.state_dict()), strict=True)
    head.eval()

    g = torch.Generator().manual_seed(11)
    pts83 = torch.randn(2, 166, 5, 6, generator=g) * 2
    pts9 = torch.randn(50, 18, generator=g) * 2
    with torch.no_grad():
        np.savez_compressed(os.path.join(HERE, 'moment.npz'), pts83=pts83.numpy(), pts9=pts9.numpy(),
                            mt=head.moment_transfer.detach().numpy(),
                            bbox83=head.points2bbox(pts83).numpy(),
                            bbox9=head.points2bbox(pts9, y_first=False).numpy())

    names = ['cls_1', 'cls_2', 'cls_3', 'kpt_1', 'kpt_2', 'kpt_3', 'bbox_1', 'bbox_2', 'bbox_3']
    x7 = torch.randn(2, 256, 7, 11, generator=g)
    with torch.no_grad():
        out7 = head.forward_single(x7)
    np.savez_compressed(os.path.join(HERE, 'head_p7.npz'), x=x7.numpy(),
                        **{n: o.numpy() for n, o in zip(names, out7)})

    x5 = torch.randn(1, 256, 25, 42, generator=g)
    with torch.no_grad():
        out5 = head.forward_single(x5)
    d5 = dict(x=x5.numpy())
    for n, o in zip(names, out5):
        d5[n + '_sum'] = np.float64(o.double().sum().item())
        d5[n + '_abs'] = np.float64(o.double().abs().sum().item())
        d5[n + '_sample'] = strided_sample(o)
    np.savez_compressed(os.path.join(HERE, 'head_p5.npz'), **d5)

    # ---- get_bboxes ---------------------------------------------------------------------------------
    sc = torch.rand(2, 13, 7, 11, generator=g) ** 3
    logit = torch.log(sc / (1 - sc))
    kpt3 = out7[5]
    bbox3 = out7[8] * 6                      # wider boxes so that NMS suppresses something
    tc = refshim.AttrDict(cfg['test_cfg'])
    metas = [dict(img_shape=(800, 1333, 3), scale_factor=1.0)] * 2
    with torch.no_grad():
        res = head.get_bboxes([out7[0]], [out7[1]], [logit], [out7[3]], [out7[4]], [kpt3], [out7[6]], [out7[7]],
                              [bbox3], metas, tc, rescale=False)
    gb = dict(logit=logit.numpy(), kpt3=kpt3.numpy(), bbox3=bbox3.numpy())
    for i, (d, l, k) in enumerate(res):
        gb['dets_%d' % i] = d.numpy()
        gb['labels_%d' % i] = l.numpy()
        gb['kpts_%d' % i] = k.reshape(d.shape[0], -1).numpy()
    np.savez_compressed(os.path.join(HERE, 'get_bboxes.npz'), **gb)
    # ---- RepPoints-Kp parallel / serial heads (BASELINE.json configs[3]) ---------------------------------
    names5 = ['cls', 'kpt_init', 'kpt_refine', 'rep_init', 'rep_refine']
    sizes = [(13, 21), (7, 11), (4, 6)]                 # three small "levels" (P6, P7 and a ragged one)
    for variant in ('parallel', 'serial'):
        h2, _ = refshim.build_head('reppoints_moment_%s_r50_fpn_1x-deepfashion2.py' % variant)
        h2.load_state_dict(fill_state_dict(h2.state_dict(), seed=4321), strict=True)
        h2.eval()
        dd = {}
        for li, (hh, ww) in enumerate(sizes):
            xl = torch.randn(2, 256, hh, ww, generator=g)
            dd['x%d' % li] = xl.numpy()
            with torch.no_grad():
                outs = h2.forward_single(xl)
                dd['bbox_refine%d' % li] = h2.points2bbox(outs[4]).numpy()
            for n, o in zip(names5, outs):
                dd['%s%d' % (n, li)] = o.numpy()
        np.savez_compressed(os.path.join(HERE, 'reppoints_%s.npz' % variant), **dd)
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print('%-18s %8.1f KB' % (f, os.path.getsize(os.path.join(HERE, f)) / 1024))


if __name__ == '__main__':
    main()
