"""serve() with the two graph instances on one compute stream vs on two (concurrent=True): images/s with HOST
buffers (H2D + D2H inside the timed region), with and without the L2 flush before every step, and a check that both
modes return the same detections.

    python tools/e2e_concurrent_probe.py [steps]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from kgdet_b200 import head as head_mod  # noqa: E402
from kgdet_b200 import ops  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ops.set_precision('bf16')
dev = torch.device('cuda', 0)
head = bench.make_weights(head_mod.KGDetHead()).to(dev).eval()
x, sc = bench.make_inputs(16, 0)
x_dev, sc_dev = x.to(dev), sc.to(dev)
shapes = [bench.IMG_SHAPE] * 16
g = head_mod.GraphedInference(head, x_dev, shapes, 0.05, 0.5, 1000, 100, score_override=sc_dev)
hosts = [(x + 0.01 * i).clone().pin_memory() for i in range(3)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
outs = [tuple(torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in g.static_out) for _ in range(steps)]
ref = None
for conc in (False, 2, 3, 4, False, 2, 3, 4):
    for fl in (True, False):
        before = (lambda i: flush.fill_(1)) if fl else None
        g.serve([hosts[i % 3] for i in range(6)], None, before_step=before, concurrent=bool(conc), instances=conc or 2)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        res = g.serve([hosts[i % 3] for i in range(steps)], outs, before_step=before, concurrent=bool(conc), instances=conc or 2)
        b.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = a.elapsed_time(b) / steps
        if ref is None:
            ref = [tuple(t.clone() for t in r) for r in res[:3]]
        same = all(torch.equal(p, q) for r, rr in zip(res[:3], ref) for p, q in zip(r, rr))
        if not same:
            for bi, (r, rr) in enumerate(zip(res[:3], ref)):
                for ti, (p, q) in enumerate(zip(r, rr)):
                    if not torch.equal(p, q):
                        dd = (p.double() - q.double()).abs()
                        print('  differs: batch', bi, 'tensor', ti, tuple(p.shape), 'n_diff', int((dd > 0).sum()),
                              'max', float(dd.max()), flush=True)
        print(json.dumps(dict(concurrent=conc, flush=fl, ms_per_step=round(ms, 4), images_per_s=round(16e3 / ms, 1),
                              wall_ms_per_step=round(wall * 1e3 / steps, 4), same_results=same)), flush=True)
