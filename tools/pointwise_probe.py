"""Times the pointwise stage kernels at the KGDet shapes (CUDA events, warm L2)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kgdet_b200 import ops  # noqa: E402
from tools.tower_probe import timed  # noqa: E402


def main():
    n, h, w = 16, 25, 42
    m = n * h * w
    out = {}
    for k, nouts in ((768, (588, 166)), (768, (13,)), (256, (588, 166)), (256, (13,))):
        x = torch.randn(n, k, h, w, device='cuda')
        wt = torch.randn(sum(nouts), k, device='cuda')
        bias = torch.randn(sum(nouts), device='cuda')
        outs, c0 = [], 0
        for no in nouts:
            outs.append((torch.empty(n, no, h, w, device='cuda'), torch.randn(n, no, h, w, device='cuda'), c0, c0 + no))
            c0 += no
        for split in (False, True):
            rows = ops.nchw_to_tiled(x, split=split)
            pw = ops.pack_weight(wt, split=split)
            out['pointwise_%sK%d_N%d_us' % ('split_' if split else '', k, sum(nouts))] = timed(
                lambda: ops.pointwise_conv(rows, pw, bias, outs, h * w))
    x = torch.randn(n, 256, h, w, device='cuda')
    out['nchw_to_tiled_split_us'] = timed(lambda: ops.nchw_to_tiled(x, relu=True, split=True))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
