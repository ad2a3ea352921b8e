"""Per-tile timeline of the grouped persistent DCN kernel on the KGDet stage shape (GPU box).
Stamps (clock64 of the CTA's SM): 0 tile start, 1 first stage full, 2 last MMA issued, 6 last A tile handed over,
3 accumulator ready, 4 epilogue done; [5] = k-blocks of the tile."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kgdet_b200 import ops  # noqa: E402
from kgdet_b200.ops import _capi  # noqa: E402


def main():
    lib = _capi.lib()
    g = torch.Generator().manual_seed(0)
    N, C, H, W, F = 16, 256, 25, 42, 256
    xa = torch.randn(N, C, H, W, generator=g).cuda()
    xb = torch.randn(N, C, H, W, generator=g).cuda()
    pts = (torch.randn(N, 166, H, W, generator=g) * 2).cuda()
    ws = {(br, k): (torch.randn(F, C, k, k, generator=g) * 0.02).cuda() for br in 'ab' for k in (3, 5, 7)}
    ops.set_precision('bf16')
    pa, pb = ops.prepare_input(xa, F), ops.prepare_input(xb, F)
    plans, lo = {}, 0
    for k in (3, 5, 7):
        plans[k] = ops.prepare_plan_points(pts, lo, (N, C, H, W), F, k, 1, k // 2, 1)
        lo += 2 * k * k
    rows = {br: ops.TiledRows(N * H * W, 3 * F, True, 'cuda') for br in 'ab'}
    jobs = []
    for i, k in enumerate((3, 5, 7)):
        jobs.append((pa, plans[k], ws[('a', k)], rows['a'], i * F, True))
        jobs.append((pb, plans[k], ws[('b', k)], rows['b'], i * F, True))
    jobs.reverse()
    for _ in range(3):
        ops.deform_conv_prepared_group(jobs)
    torch.cuda.synchronize()
    nsm = torch.cuda.get_device_properties(0).multi_processor_count
    tl = torch.zeros(nsm * 64 * 8, dtype=torch.int64, device='cuda')
    lib.kgdet_dcn_set_timeline(tl.data_ptr(), tl.numel())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.deform_conv_prepared_group(jobs)
    b.record()
    torch.cuda.synchronize()
    t = tl.view(nsm, 64, 8).cpu()
    rec = {'kernel_us': a.elapsed_time(b) * 1e3}
    per = {}
    spans = []
    for c in range(nsm):
        tiles = [t[c, i] for i in range(64) if int(t[c, i, 0]) != 0]
        if not tiles:
            continue
        spans.append(int(tiles[-1][4] - tiles[0][0]))
        for i, s in enumerate(tiles):
            nkb = int(s[5])
            d = per.setdefault(nkb, {'n': 0, 'fill': 0, 'main': 0, 'drain': 0, 'epi': 0, 'gap': 0, 'total': 0})
            d['n'] += 1
            d['fill'] += int(s[1] - s[0])
            d['main'] += int(s[2] - s[1])
            d['drain'] += int(s[3] - s[2])
            d['epi'] += int(s[4] - s[3])
            d['total'] += int(s[4] - s[0])
            if i + 1 < len(tiles):
                d['gap'] += int(tiles[i + 1][0] - s[4])
    for nkb, d in sorted(per.items()):
        n = d.pop('n')
        rec['nkb_%d' % nkb] = {'tiles': n, **{k: round(v / n) for k, v in d.items()},
                               'clk_per_kblock_main': round(d['main'] / n / max(nkb - 1, 1))}
    rec['cta_span_clk'] = {'min': min(spans), 'max': max(spans), 'mean': round(sum(spans) / len(spans))}
    rec['tiles_per_cta'] = round(sum(v['tiles'] for k, v in rec.items() if k.startswith('nkb_')) / len(spans), 2)
    print(json.dumps(rec))


if __name__ == '__main__':
    main()
