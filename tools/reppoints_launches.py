"""One batch-8 five-level forward of a RepPoints-Kp baseline head between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/reppoints_launches.py parallel 8 [bboxes]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kgdet_b200 import ops  # noqa: E402
from kgdet_b200.head import RepPointsKpDetect, RepPointsKpHead  # noqa: E402
from tools.reppoints_bench import LEVELS  # noqa: E402


def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else 'parallel'
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    ops.set_precision('bf16')
    head = RepPointsKpHead(variant).cuda().eval()
    g = torch.Generator().manual_seed(5)
    feats = [torch.randn(batch, 256, h, w, generator=g).cuda() for h, w in LEVELS]
    scores = [(torch.rand(batch, 13, h, w, generator=g) ** 28).cuda() for h, w in LEVELS]
    fn = RepPointsKpDetect(head, [(800, 1333)] * batch, score_override=scores) if 'bboxes' in sys.argv else head
    with torch.no_grad():
        for _ in range(2):
            fn(feats)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        fn(feats)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


if __name__ == '__main__':
    main()
