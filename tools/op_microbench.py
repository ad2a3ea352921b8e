"""Per-kernel timings + achieved bandwidth of NMS, focal loss and the moment transform on one B200.

    python tools/op_microbench.py > profiles/<name>.jsonl

Algorithmic bytes:
  NMS      20 n (boxes) + 2 * n * ceil(n/64) * 8 (mask write + read, large-n path only) + 8 n_keep
  focal    M*C*4 read + 8*M targets + M*C*4 write (fwd); bwd adds d_loss read + d_logit write
  moment   4*N*2P*S read + 16*N*S write
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kgdet_b200 import ops  # noqa: E402
from kgdet_b200.ops.nms import nms_wrapper  # noqa: E402
from tests._data import random_boxes  # noqa: E402


def timed(fn, flush, reps=10, warm=3, clean=False):
    """clean=False: L2 flushed by WRITING 256 MiB (bench.py's rule) -- the 126 MB of dirty lines are written back
    while the timed kernel runs and count against it; clean=True: flushed by READING the buffer (clean lines)."""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        if clean:
            flush.sum()
        else:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3          # us


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    # ---- NMS: single segment ----
    for n in (1000, 3350, 4096, 16384, 32768, 65536):
        dets = random_boxes(n, seed=n, clustered=True).cuda()
        us = timed(lambda: nms_wrapper._nms_keep_cuda(dets, 0.5, 0), flush, reps=5)
        cb = (n + 63) // 64
        mask_bytes = 2 * n * cb * 8 if n > 4096 else 0
        row = dict(op='nms', n=n, us=round(us, 1), path='mask+sweep' if n > 4096 else 'single-CTA',
                   note='includes the .item() sync of the wrapper')
        if mask_bytes:
            row['alg_MB'] = round((20 * n + mask_bytes) / 1e6, 1)
            row['GBps'] = round((20 * n + mask_bytes) / us / 1e3, 1)
        print(json.dumps(row), flush=True)
    # ---- NMS: batched (16 images x 13 classes), KGDet post-processing shape ----
    for per_seg in (100, 400, 1000):
        nseg = 208
        dets = torch.cat([random_boxes(per_seg, seed=s, clustered=True) for s in range(nseg)]).cuda()
        offs = torch.arange(0, (nseg + 1) * per_seg, per_seg, dtype=torch.int32).cuda()
        us = timed(lambda: ops.batched_nms_flags(dets, offs, per_seg, 0.5), flush, reps=5)
        print(json.dumps(dict(op='nms_batched', segments=nseg, boxes_per_segment=per_seg, us=round(us, 1),
                              pairs_per_us=round(nseg * per_seg * (per_seg - 1) / 2 / us, 1))), flush=True)
    # ---- focal loss ----
    for M, dense in ((2100, True), (22400 * 2, True), (22400 * 8, True), (22400 * 64, True), (22400 * 8, False),
                     (22400 * 64, False)):
        C = 13
        x = torch.randn(M, C, device='cuda')
        t = torch.randint(0, C + 1, (M,), device='cuda')
        if not dense:           # a detector's labels: ~1 % of the points are positives, the rest background (0)
            t = t * (torch.rand(M, device='cuda') < 0.01).long()
        d = torch.rand(M, C, device='cuda')
        f = lambda: ops.sigmoid_focal_loss(x, t, 2.0, 0.25)
        us_f = timed(f, flush)
        us_f_clean = timed(f, flush, clean=True)
        xg = x.clone().requires_grad_()

        def fb():
            xg.grad = None
            ops.sigmoid_focal_loss(xg, t, 2.0, 0.25).backward(d)
        us_fb = timed(fb, flush)
        w = torch.rand(M, device='cuda')
        us_sum = timed(lambda: ops.sigmoid_focal_loss_sum(x, t, w, 2.0, 0.25), flush)
        fwd_bytes = M * C * 4 * 2 + 8 * M
        print(json.dumps(dict(op='focal', M=M, C=C, labels='uniform 0..13' if dense else '1 % positive rows', fwd_us=round(us_f, 1), fwd_GBps=round(fwd_bytes / us_f / 1e3, 1),
                              fwd_us_read_flushed_l2=round(us_f_clean, 1),
                              fwd_GBps_read_flushed_l2=round(fwd_bytes / us_f_clean / 1e3, 1),
                              fwd_bwd_us=round(us_fb, 1), fused_sum_us=round(us_sum, 1),
                              fused_sum_GBps=round((M * C * 4 + 12 * M) / us_sum / 1e3, 1))), flush=True)
    # ---- moment transform ----
    for shape in ((16, 166, 25, 42), (8, 18, 100, 168), (64, 166, 25, 42), (64, 18, 100, 168)):
        p = torch.randn(*shape, device='cuda')
        mt = torch.zeros(2, device='cuda')
        us = timed(lambda: ops.points2bbox_moment(p, mt), flush)
        N, P2, H, W = shape
        by = 4 * N * P2 * H * W + 16 * N * H * W
        pg = p.clone().requires_grad_()
        mg = mt.clone().requires_grad_()
        gb = torch.randn(N, 4, H, W, device='cuda')

        def fb():
            pg.grad = None; mg.grad = None
            ops.points2bbox_moment(pg, mg).backward(gb)
        us_fb = timed(fb, flush)
        print(json.dumps(dict(op='moment', shape=list(shape), fwd_us=round(us, 1), fwd_GBps=round(by / us / 1e3, 1),
                              fwd_bwd_us=round(us_fb, 1))), flush=True)

    # ---- streaming GroupNorm (+ ReLU) on FPN-sized maps (round 2): 3 x 4 B per value (two reads, one write) ----
    from kgdet_b200.ops.pointwise import groupnorm_relu_nhwc
    from kgdet_b200.ops.decode import bbox_select
    gn = torch.nn.GroupNorm(32, 256).cuda()
    for shape in ((8, 256, 100, 168), (8, 256, 50, 84), (1, 256, 100, 168)):
        x = torch.randn(*shape, device='cuda').contiguous(memory_format=torch.channels_last)
        us = timed(lambda: groupnorm_relu_nhwc(x, gn), flush)
        us_p = timed(lambda: groupnorm_relu_nhwc(x, gn, dense=False, prepared_for=256), flush)
        by = x.numel() * 4 * 3
        print(json.dumps(dict(op='groupnorm_stream', shape=list(shape), dense_us=round(us, 1),
                              dense_GBps=round(by / us / 1e3, 1), planes_only_us=round(us_p, 1),
                              note='stats + finalize + apply; bytes = 2 reads + 1 write of fp32')), flush=True)
    # ---- candidate selection on large levels (round 2): keys + radix select + ranks ----
    for B, hw, n in ((8, (100, 168), 1000), (8, (50, 84), 1000), (1, (100, 168), 1000), (16, (25, 42), 1000)):
        sc = torch.randn(B, 13, *hw, device='cuda')
        us = timed(lambda: bbox_select(sc, True, min(n, hw[0] * hw[1])), flush)
        print(json.dumps(dict(op='bbox_select', batch=B, map=list(hw), n=n, us=round(us, 1),
                              score_GBps=round(sc.numel() * 4 / us / 1e3, 1))), flush=True)


if __name__ == '__main__':
    main()
