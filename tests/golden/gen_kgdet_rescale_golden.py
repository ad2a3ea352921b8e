"""Golden for the KGDet head's get_bboxes with rescale=True (KP3:892-898: boxes and keypoint coordinates divided by
the image's scale factor BEFORE multiclass_nms_kp), from the UNCHANGED reference class through tests/refshim.py on
the inputs already stored in tests/golden/get_bboxes.npz.

    python -m tests.golden.gen_kgdet_rescale_golden     # writes tests/golden/get_bboxes_rescale.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCALE = 1.6


def main():
    from tests import refshim
    refshim.install('oracle')
    g = np.load(os.path.join(HERE, 'get_bboxes.npz'))
    head, cfg = refshim.build_head('kgdet_moment_r50_fpn_1x-demo.py')
    head.eval()
    t = lambda n: torch.from_numpy(g[n])                                   # noqa: E731
    tc = refshim.AttrDict(cfg['test_cfg'])
    metas = [dict(img_shape=(800, 1333, 3), scale_factor=SCALE)] * 2
    dummy = [t('logit')]
    with torch.no_grad():       # only the stage-3 tensors are read (KP3:784-786)
        res = head.get_bboxes(dummy, dummy, [t('logit')], [t('kpt3')], [t('kpt3')], [t('kpt3')], [t('bbox3')],
                              [t('bbox3')], [t('bbox3')], metas, tc, rescale=True)
    out = {}
    for i, (d, l, k) in enumerate(res):
        out['dets_%d' % i] = d.numpy()
        out['labels_%d' % i] = l.numpy()
        out['kpts_%d' % i] = k.reshape(d.shape[0], -1).numpy()
        print(i, tuple(d.shape), tuple(k.shape), flush=True)
    np.savez_compressed(os.path.join(HERE, 'get_bboxes_rescale.npz'), **out)


if __name__ == '__main__':
    main()
