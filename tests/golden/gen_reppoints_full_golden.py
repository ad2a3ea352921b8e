"""Full-size golden for BASELINE.json configs[3]: the UNCHANGED reference RepPoints-Kp parallel / serial heads
(reppoints_head_kp_{parallel,serial}.py) on all five FPN levels of an 800x1333 image (P3 100x168 ... P7 7x11),
batch 1, run on the CPU through tests/refshim.py with the oracle DeformConv.

    python -m tests.golden.gen_reppoints_full_golden      # writes tests/golden/reppoints_full_{parallel,serial}.npz

The outputs are too big to store (588 channels x 22 400 positions): per output tensor the fixture keeps the sum of
absolute values, the plain sum and a strided sample of 4096 values; inputs are regenerated from the seed
(`make_inputs`), weights by `fill_state_dict(seed=4321)`.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LEVELS = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
NAMES = ['cls', 'kpt_init', 'kpt_refine', 'rep_init', 'rep_refine']


def make_inputs(seed=77):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(1, 256, h, w, generator=g) for h, w in LEVELS]


def strided_sample(t, n=4096):
    f = t.detach().reshape(-1)
    step = max(f.numel() // n, 1)
    return f[::step][:n].numpy().copy()


def main():
    from tests import refshim
    from tests.golden.gen_golden import fill_state_dict
    refshim.install('oracle')
    torch.set_num_threads(os.cpu_count() or 1)
    xs = make_inputs()
    for variant in ('parallel', 'serial'):
        head, _ = refshim.build_head('reppoints_moment_%s_r50_fpn_1x-deepfashion2.py' % variant)
        head.load_state_dict(fill_state_dict(head.state_dict(), seed=4321), strict=True)
        head.eval()
        d = {'x_checksum': np.float64(sum(float(x.double().abs().sum()) for x in xs))}
        for li, x in enumerate(xs):
            with torch.no_grad():
                outs = head.forward_single(x)
                bbox = head.points2bbox(outs[4])
            for n, o in list(zip(NAMES, outs)) + [('bbox_refine', bbox)]:
                d['%s%d_abs' % (n, li)] = np.float64(o.double().abs().sum().item())
                d['%s%d_sum' % (n, li)] = np.float64(o.double().sum().item())
                d['%s%d_sample' % (n, li)] = strided_sample(o)
            print(variant, 'level', li, tuple(x.shape[-2:]), 'done', flush=True)
        np.savez_compressed(os.path.join(HERE, 'reppoints_full_%s.npz' % variant), **d)


if __name__ == '__main__':
    main()
