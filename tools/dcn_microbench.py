"""Per-kernel timings of the deformable convolution on one B200 (CUDA events, L2 flushed).

    python tools/dcn_microbench.py [--ref]   # --ref also times the reference CUDA op (oracle/_ref)

Writes one JSON object per shape to stdout.  Shapes: the three KGDet calls ([16,256,25,42],
K = 9/25/49) and the RepPoints sweep P3..P7 at batch 8 (BASELINE.json configs[1]).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kgdet_b200 import ops  # noqa: E402
from kgdet_b200.ops import _capi  # noqa: E402
from tests._data import dcn_case  # noqa: E402


def timed(fn, flush, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def kernel_only(fn, flush, reps=10, warm=3):
    lib = _capi.lib()
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); b.record()
        lib.kgdet_dcn_set_profile_events(a.cuda_event, b.cuda_event)
        fn()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def quick():
    """bf16 fused-kernel time of the three KGDet calls only (A/B runs of kernel variants via environment)."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ops.set_precision('bf16')
    row = {k: v for k, v in os.environ.items() if k.startswith('KGDET_')}
    for name, k in (('k3', 3), ('k5', 5), ('k7', 7)):
        d = dcn_case(N=16, C=256, H=25, W=42, Cout=256, k=k)
        x, off, w = (d[q].cuda() for q in ('x', 'offset', 'weight'))
        ms = kernel_only(lambda: ops.deform_conv(x, off, w, 1, k // 2), flush)
        row[name + '_us'] = round(ms * 1e3, 1)
        row[name + '_tflops'] = round(2.0 * 16 * 25 * 42 * 256 * 256 * k * k / ms / 1e9, 1)
        if '--bwd' in sys.argv:
            go = d['grad_out'].cuda()
            xg, og, wg = x.clone().requires_grad_(), off.clone().requires_grad_(), w.clone().requires_grad_()

            def fb():
                xg.grad = og.grad = wg.grad = None
                ops.deform_conv(xg, og, wg, 1, k // 2).backward(go)
            row[name + '_fwd_bwd_us'] = round(timed(fb, flush, reps=5, warm=2) * 1e3, 1)
    print(json.dumps(row), flush=True)


def main():
    use_ref = '--ref' in sys.argv
    if '--quick' in sys.argv:
        return quick()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    shapes = [('kgdet_k3', 16, 25, 42, 3), ('kgdet_k5', 16, 25, 42, 5), ('kgdet_k7', 16, 25, 42, 7),
              ('P3', 8, 100, 168, 3), ('P4', 8, 50, 84, 3), ('P5', 8, 25, 42, 3), ('P6', 8, 13, 21, 3),
              ('P7', 8, 7, 11, 3)]
    ref = None
    if use_ref:
        from oracle import build_ref
        ref = build_ref.load('deform_conv_cuda')
    for name, N, H, W, k in shapes:
        d = dcn_case(N=N, C=256, H=H, W=W, Cout=256, k=k)
        x, off, w, go = (d[q].cuda() for q in ('x', 'offset', 'weight', 'grad_out'))
        flops = 2.0 * N * H * W * 256 * 256 * k * k
        row = dict(shape=name, N=N, H=H, W=W, k=k, gflop=round(flops / 1e9, 2))
        for prec in ('bf16', 'tf32x3', 'tf32'):
            ops.set_precision(prec)
            f = lambda: ops.deform_conv(x, off, w, 1, k // 2)
            ms_call = timed(f, flush)
            ms_k = kernel_only(f, flush)
            row[prec] = dict(call_us=round(ms_call * 1e3, 1), kernel_us=round(ms_k * 1e3, 1),
                             kernel_tflops=round(flops / ms_k / 1e9, 1))
        ops.set_precision('bf16')
        xg, og, wg = x.clone().requires_grad_(), off.clone().requires_grad_(), w.clone().requires_grad_()

        def fb16():
            xg.grad = og.grad = wg.grad = None
            ops.deform_conv(xg, og, wg, 1, k // 2).backward(go)
        row['bf16_fwd_bwd_us'] = round(timed(fb16, flush, reps=5, warm=2) * 1e3, 1)
        if N * H * W <= 20000:
            ops.set_precision('fp32')
            f = lambda: ops.deform_conv(x, off, w, 1, k // 2)
            ms_k = kernel_only(f, flush, reps=3, warm=1)
            row['fp32_simt'] = dict(kernel_us=round(ms_k * 1e3, 1), kernel_tflops=round(flops / ms_k / 1e9, 1))
            xg, og, wg = x.clone().requires_grad_(), off.clone().requires_grad_(), w.clone().requires_grad_()

            def fb():
                xg.grad = og.grad = wg.grad = None
                ops.deform_conv(xg, og, wg, 1, k // 2).backward(go)
            row['fp32_simt_fwd_bwd_us'] = round(timed(fb, flush, reps=3, warm=1) * 1e3, 1)
        ops.set_precision(None)
        if ref is not None:
            out = x.new_empty(N, 256, H, W)
            bufs = [x.new_empty(0), x.new_empty(0)]
            step = min(64, N)
            fr = lambda: ref.deform_conv_forward_cuda(x, w, off, out, bufs[0], bufs[1], k, k, 1, 1, k // 2, k // 2,
                                                      1, 1, 1, 1, step)
            ms = timed(fr, flush, reps=5, warm=2)
            row['reference_cuda_fwd'] = dict(call_us=round(ms * 1e3, 1), tflops=round(flops / ms / 1e9, 1))
        print(json.dumps(row), flush=True)


if __name__ == '__main__':
    main()
