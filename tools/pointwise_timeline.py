"""Pipeline timeline of the pointwise tcgen05 GEMM (development tool): per shape the medians over CTAs of set-up,
first full stage, per-channel-block interval, accumulator-ready -> epilogue-done and CTA total, in SM clocks, plus
the spread of CTA start times (waves)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kgdet_b200 import ops  # noqa: E402
from kgdet_b200.ops import _capi  # noqa: E402


def main():
    lib = _capi.lib()
    n, h, w = 16, 25, 42
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for k, nouts, resid in ((768, (588, 166), True), (768, (588, 166), False), (256, (588, 166), False), (768, (13,), False)):
        x = torch.randn(n, k, h, w, device='cuda')
        wt = torch.randn(sum(nouts), k, device='cuda')
        bias = torch.randn(sum(nouts), device='cuda')
        outs, c0 = [], 0
        for no in nouts:
            outs.append((torch.empty(n, no, h, w, device='cuda'),
                         torch.randn(n, no, h, w, device='cuda') if resid else None, c0, c0 + no))
            c0 += no
        rows = ops.nchw_to_tiled(x, split=True)
        pw = ops.pack_weight(wt, split=True)
        for _ in range(3):
            ops.pointwise_conv(rows, pw, bias, outs, h * w)
        kg = k // 64
        bn = 64 if sum(nouts) <= 64 else 256
        ctas = ((n * h * w + 127) // 128) * ((sum(nouts) + bn - 1) // bn)
        per = kg + 8
        for cold in (True, False):
            buf = torch.zeros(ctas * per, dtype=torch.int64, device='cuda')
            if cold:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            lib.kgdet_dcn_set_timeline(buf.data_ptr(), buf.numel())
            a.record()
            ops.pointwise_conv(rows, pw, bias, outs, h * w)
            b.record()
            torch.cuda.synchronize()
            grid = min(ctas, torch.cuda.get_device_properties(0).multi_processor_count)   # persistent CTAs: first tile of each
            t = buf.view(ctas, per)[:grid].cpu().double()
            med = lambda v: float(v.median())
            full = t[:, 2:2 + kg]
            iv = (full[:, 1:] - full[:, :-1]).median(0).values
            print(json.dumps(dict(K=k, N=sum(nouts), residual=resid, cold_l2=cold, ctas=ctas, kernel_us=round(a.elapsed_time(b) * 1e3, 1),
                                  setup=med(t[:, 1] - t[:, 0]), first_full=med(full[:, 0] - t[:, 0]),
                                  intervals=[round(float(v)) for v in iv],
                                  mainloop=med(full[:, -1] - full[:, 0]),
                                  acc_ready_after_last_full=med(t[:, 2 + kg] - full[:, -1]),
                                  epilogue=med(t[:, 3 + kg] - t[:, 2 + kg]), total=med(t[:, 3 + kg] - t[:, 0]))), flush=True)


if __name__ == '__main__':
    main()
