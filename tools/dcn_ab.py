"""A/B timing of the fused bf16 DCN forward kernel under the current environment knobs
(KGDET_UMMA_PAIR, KGDET_UMMA_STAGES): kernel-only CUDA-event times, L2 flushed, plus a max-error check
against the exact fp32 path so that a fast-but-wrong variant is visible in the same line.

    KGDET_UMMA_PAIR=0 python tools/dcn_ab.py          # one JSON line
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kgdet_b200 import ops  # noqa: E402
from tests._data import dcn_case, rel_err  # noqa: E402
from tools.dcn_microbench import kernel_only  # noqa: E402


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    shapes = [('k3', 16, 25, 42, 3), ('k5', 16, 25, 42, 5), ('k7', 16, 25, 42, 7), ('P3', 8, 100, 168, 3),
              ('P4', 8, 50, 84, 3), ('P7', 8, 7, 11, 3)]
    row = {'pair': os.environ.get('KGDET_UMMA_PAIR', 'default'), 'stages': os.environ.get('KGDET_UMMA_STAGES', 'default')}
    for name, N, H, W, k in shapes:
        d = dcn_case(N=N, C=256, H=H, W=W, Cout=256, k=k)
        x, off, w = (d[q].cuda() for q in ('x', 'offset', 'weight'))
        flops = 2.0 * N * H * W * 256 * 256 * k * k
        ops.set_precision('bf16')
        f = lambda: ops.deform_conv(x, off, w, 1, k // 2)
        out = f()
        ms = kernel_only(f, flush)
        err = None
        if N * H * W <= 20000:
            ops.set_precision('fp32')
            err = rel_err(out, f())
        row[name] = dict(us=round(ms * 1e3, 1), tflops=round(flops / ms / 1e9, 1), err=err)
    ops.set_precision(None)
    print(json.dumps(row), flush=True)


if __name__ == '__main__':
    main()
