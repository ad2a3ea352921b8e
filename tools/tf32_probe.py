"""Calibration probe: head-level errors against the reference goldens with cuDNN TF32 on / off (GPU box).
Prints one JSON line per (test, precision, tf32)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests._data import rel_err                      # noqa: E402
from tests.golden.gen_golden import fill_state_dict  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
NAMES = ['cls_1', 'cls_2', 'cls_3', 'kpt_1', 'kpt_2', 'kpt_3', 'bbox_1', 'bbox_2', 'bbox_3']


def head():
    from kgdet_b200.head import KGDetHead
    h = KGDetHead()
    h.load_state_dict(fill_state_dict(h.state_dict()), strict=True)
    return h.cuda().eval()


def main():
    from kgdet_b200 import ops
    g7 = np.load(os.path.join(GOLD, 'head_p7.npz'))
    g5 = np.load(os.path.join(GOLD, 'head_p5.npz'))
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        for prec in ('fp32', 'tf32x3', 'bf16'):
            ops.set_precision(prec)
            with torch.no_grad():
                o7 = head().forward_single(torch.from_numpy(g7['x']).cuda())
                o5 = head().forward_single(torch.from_numpy(g5['x']).cuda())
            e7 = {n: rel_err(o, torch.from_numpy(g7[n])) for n, o in zip(NAMES, o7)}
            e5 = {}
            for n, o in zip(NAMES, o5):
                f = o.reshape(-1)
                step = max(f.numel() // 4096, 1)
                e5[n] = rel_err(f[::step][:4096].cpu(), torch.from_numpy(g5[n + '_sample']))
            print(json.dumps({'tf32': tf32, 'prec': prec, 'p7': {k: float('%.3g' % v) for k, v in e7.items()},
                              'p5': {k: float('%.3g' % v) for k, v in e5.items()}}), flush=True)
        ops.set_precision(None)
    # full size: fused graph path vs module path, both settings
    gen = torch.Generator().manual_seed(41)
    x = torch.randn(16, 256, 25, 42, generator=gen).cuda()
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        ops.set_precision('bf16')
        h = head()
        with torch.no_grad():
            fused = h.forward_single(x)
            h._fused_inference = False
            plain = h.forward_single(x)
        ops.set_precision('fp32')
        with torch.no_grad():
            exact = h.forward_single(x)
        ops.set_precision(None)
        print(json.dumps({'tf32': tf32, 'full_size_fused_vs_module': {n: float('%.3g' % rel_err(a, b)) for n, a, b in zip(NAMES, fused, plain)},
                          'fused_vs_exact_fp32_module': {n: float('%.3g' % rel_err(a, b)) for n, a, b in zip(NAMES, fused, exact)}}), flush=True)


if __name__ == '__main__':
    main()
