"""Tensor-core tower convolution (kgdet_conv_forward) vs cuDNN on the KGDet shapes (GPU box).  One JSON line per case."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kgdet_b200.ops import conv  # noqa: E402


def timed(fn, flush, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for (N, C, H, W, Cout, k) in [(16, 256, 25, 42, 256, 3), (2, 256, 25, 42, 256, 3), (8, 256, 100, 168, 256, 3),
                                  (8, 256, 50, 84, 256, 3), (8, 256, 13, 21, 256, 3)]:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(N, C, H, W, generator=g).cuda()
        w = (torch.randn(Cout, C, k, k, generator=g) * 0.02).cuda()
        xcl = x.contiguous(memory_format=torch.channels_last)
        wcl = w.contiguous(memory_format=torch.channels_last)
        planes = conv.split_planes(x)
        flops = 2.0 * N * H * W * C * Cout * k * k
        rec = {'shape': [N, C, H, W, Cout, k]}
        us = timed(lambda: conv.conv_planes(planes, w), flush)
        rec['kgdet_conv_us'] = round(us, 1)
        rec['kgdet_conv_tflops_useful'] = round(flops / us / 1e6, 1)
        rec['kgdet_conv_tflops_issued_bf16'] = round(3 * flops / us / 1e6, 1)
        rec['split_planes_us'] = round(timed(lambda: conv.split_planes(xcl), flush), 1)
        for name, tf32 in (('cudnn_tf32_us', True), ('cudnn_fp32_us', False)):
            torch.backends.cudnn.allow_tf32 = tf32
            rec[name] = round(timed(lambda: F.conv2d(xcl, wcl, padding=k // 2), flush, reps=5 if not tf32 else 20), 1)
        torch.backends.cudnn.allow_tf32 = True
        rec['cudnn_bf16_us'] = round(timed(lambda: F.conv2d(xcl.bfloat16(), wcl.bfloat16(), padding=k // 2), flush), 1)
        print(json.dumps(rec), flush=True)


if __name__ == '__main__':
    main()
