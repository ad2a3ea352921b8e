"""Per-CTA timeline of the tensor-core tower convolution on the KGDet shape (GPU box)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kgdet_b200.ops import _capi, conv  # noqa: E402


def main():
    lib = _capi.lib()
    g = torch.Generator().manual_seed(0)
    for n in (16, 2):
        x = torch.randn(n, 256, 25, 42, generator=g).cuda()
        w = (torch.randn(256, 256, 3, 3, generator=g) * 0.02).cuda()
        planes = conv.split_planes(x)
        for _ in range(3):
            conv.conv_planes(planes, w)
        torch.cuda.synchronize()
        ncta = 2 * ((n * 9 + 1) // 2)
        tl = torch.zeros(ncta * 8, dtype=torch.int64, device='cuda')
        flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
        flush.fill_(1)
        lib.kgdet_dcn_set_timeline(tl.data_ptr(), tl.numel())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        conv.conv_planes(planes, w)
        b.record()
        torch.cuda.synchronize()
        t = tl.view(ncta, 8).cpu()
        even = t[0::2]
        rec = {'batch': n, 'kernel_us': round(a.elapsed_time(b) * 1e3, 1),
               'setup_clk': int((even[:, 1] - even[:, 0]).float().mean()),
               'first_stage_full_clk': int((even[:, 2] - even[:, 1]).float().mean()),
               'mma_issue_span_clk': int((even[:, 3] - even[:, 2]).float().mean()),
               'clk_per_kblock': int((even[:, 3] - even[:, 2]).float().mean() / 35),
               'drain_clk': int((even[:, 4] - even[:, 3]).float().mean()),
               'epilogue_clk': int((even[:, 5] - even[:, 4]).float().mean()),
               'cta_total_clk': int((even[:, 5] - even[:, 0]).float().mean()),
               'first_to_last_cta_start_clk': int(t[:, 0].max() - t[:, 0].min())}
        print(json.dumps(rec))


if __name__ == '__main__':
    main()
