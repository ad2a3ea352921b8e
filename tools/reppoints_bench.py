"""BASELINE.json configs[3]: the RepPoints-Kp parallel / serial baseline heads (reppoints_head_kp_parallel.py,
reppoints_head_kp_serial.py) on the five FPN levels of an 800x1333 image (P3 100x168 ... P7 7x11), forward only,
random-init weights, bf16 mode.  Prints one JSON line per (variant, batch): device ms per batch and images/s, the
number of deformable-convolution calls and their algorithmic TFLOP/s share.  `graph_nhwc_get_bboxes` adds the
multi-level get_bboxes (synthetic scores: random-init scores all sit at 0.01) inside the same captured graph.

    python tools/reppoints_bench.py > profiles/<name>.jsonl
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kgdet_b200 import ops  # noqa: E402
from kgdet_b200.head import GraphedForward, RepPointsKpDetect, RepPointsKpHead  # noqa: E402

LEVELS = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]


def main():
    torch.backends.cudnn.benchmark = os.environ.get('KGDET_CUDNN_BENCHMARK', '1') == '1'
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ops.set_precision('bf16')
    for variant in ('parallel', 'serial'):
        head = RepPointsKpHead(variant).cuda().eval()
        ndcn = sum(1 for m in head.modules() if isinstance(m, ops.DeformConv))
        for batch in (1, 8):
            g = torch.Generator().manual_seed(5)
            feats = [torch.randn(batch, 256, h, w, generator=g).cuda() for h, w in LEVELS]
            scores = [(torch.rand(batch, 13, h, w, generator=g) ** 28).cuda() for h, w in LEVELS]   # ~10 % pass 0.05, as bench.py
            detect = RepPointsKpDetect(head, [(800, 1333)] * batch, score_override=scores)
            for mode in ('eager_per_level', 'eager_grouped', 'graph_grouped', 'eager_nhwc', 'graph_nhwc',
                         'graph_nhwc_get_bboxes'):
                head.grouped_dcn = mode != 'eager_per_level'
                head.nhwc_towers = 'nhwc' in mode
                target = detect if mode.endswith('get_bboxes') else head
                fn = GraphedForward(target, feats) if mode.startswith('graph') else (lambda f: head(f))
                with torch.no_grad():
                    for _ in range(3):
                        fn(feats)
                    torch.cuda.synchronize()
                    ts = []
                    for _ in range(10):
                        flush.fill_(1)
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        fn(feats)
                        b.record()
                        torch.cuda.synchronize()
                        ts.append(a.elapsed_time(b))
                ts.sort()
                ms = ts[len(ts) // 2]
                positions = sum(h * w for h, w in LEVELS) * batch
                dcn_gflop = 2.0 * positions * 256 * 256 * 9 * ndcn / 1e9
                print(json.dumps(dict(variant=variant, batch=batch, levels=LEVELS, ms_per_batch=round(ms, 3),
                                      images_per_s=round(batch / (ms * 1e-3), 1), dcn_calls_per_level=ndcn,
                                      dcn_gflop=round(dcn_gflop, 1), launch_mode=mode,
                                      note='forward of all five levels, L2 flushed before every batch; 3x3 tower '
                                           'convolutions cuDNN (TF32 allowed); *_nhwc: position-major towers, own '
                                           'GroupNorm and 1x1 GEMMs')), flush=True)
    ops.set_precision(None)


if __name__ == '__main__':
    main()
