"""Pipeline timeline of the fused tcgen05 DCN forward (development tool, one B200).

    python tools/dcn_timeline.py > profiles/<name>.jsonl

Arms the library's timeline hook (kgdet_dcn_set_timeline) around one forward call per shape and prints, per
shape: set-up time, time to the first full stage, the per-k-block interval statistics seen by the MMA issuer
(median / mean / the slowest ones and where they are), accumulator-ready -> epilogue-done, and the CTA's total,
all in SM clock cycles (clock64) as medians over the CTAs of the launch.
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


def main():
    lib = _capi.lib()
    ops.set_precision('bf16')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    shapes = [('kgdet_k3', 16, 25, 42, 3), ('kgdet_k5', 16, 25, 42, 5), ('kgdet_k7', 16, 25, 42, 7),
              ('P7', 8, 7, 11, 3)]
    for name, N, H, W, k in shapes:
        d = dcn_case(N=N, C=256, H=H, W=W, Cout=256, k=k)
        x, off, w = (d[q].cuda() for q in ('x', 'offset', 'weight'))
        nkb = 4 * k * k
        ctas = (N * H * W + 127) // 128
        per = 12 * nkb + 8
        for _ in range(3):
            ops.deform_conv(x, off, w, 1, k // 2)
        buf = torch.zeros(ctas * per, dtype=torch.int64, device='cuda')
        flush.fill_(1)
        lib.kgdet_dcn_set_timeline(buf.data_ptr(), buf.numel())
        ops.deform_conv(x, off, w, 1, k // 2)
        torch.cuda.synchronize()
        t = buf.view(ctas, per).cpu().double()
        t0 = t[:, 0:1]
        full = t[:, 2:2 + nkb] - t0                      # control lane saw stage j full
        prod = t[:, 4 + nkb:4 + 2 * nkb] - t0            # producer thread 0 finished k-block j
        acq = t[:, 4 + 2 * nkb:4 + 3 * nkb] - t0         # producer thread 0 acquired the stage
        sto = t[:, 4 + 3 * nkb:4 + 4 * nkb] - t0         # ... its combine / stores / re-arm are issued
        d_full = full[:, 1:] - full[:, :-1]
        med = lambda v: float(v.median())
        iv = d_full.median(0).values                     # per-k-block interval, median over CTAs
        slow = torch.topk(iv, min(8, iv.numel()))
        row = dict(shape=name, ctas=ctas, nkb=nkb,
                   setup=med(t[:, 1] - t[:, 0]), first_full=med(full[:, 0]),
                   first_prod=med(prod[:, 0]),
                   interval_median=med(iv), interval_mean=float(iv.mean()),
                   slowest_intervals=[(int(i) + 1, float(v)) for v, i in zip(slow.values, slow.indices)],
                   mainloop=med(full[:, -1] - full[:, 0]),
                   last_full_to_acc_ready=med(t[:, 2 + nkb] - t[:, 1 + nkb]),
                   epilogue=med(t[:, 3 + nkb] - t[:, 2 + nkb]),
                   total=med(t[:, 3 + nkb] - t[:, 0]),
                   # steady state (k-blocks 8 .. nkb-4), medians over k-blocks and CTAs, in clocks
                   prod_wait_empty=med((acq[:, 8:-4] - prod[:, 7:-5])),         # arrive(kb-1) -> stage of kb acquired
                   prod_work=med((sto[:, 8:-4] - acq[:, 8:-4])),                # combine + stores + re-arm
                   prod_fence_arrive=med((prod[:, 8:-4] - sto[:, 8:-4])),       # fence.proxy.async + syncwarp + arrive
                   full_lag_after_thread0=med((full[:, 8:-4] - prod[:, 8:-4])), # slowest warp + control poll
                   recycle=med((acq[:, 11:-1] - full[:, 8:-4])),                # full(kb) seen -> stage re-acquired for kb+3
                   # arrival of each producer warp relative to the barrier completing (control lane's view), clocks
                   warp_arrive_before_full=[round(med(full[:, 8:-4] - (t[:, 4 + (4 + wq) * nkb:4 + (5 + wq) * nkb] - t0)[:, 8:-4]))
                                            for wq in range(8)],
                   intervals_first_60=[round(float(v)) for v in iv[:60]])
        print(json.dumps(row), flush=True)


if __name__ == '__main__':
    main()
