"""Stress of serve(concurrent=True): repeat the run and compare every batch's results with the sequential mode."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from kgdet_b200 import head as head_mod  # noqa: E402
from kgdet_b200 import ops  # noqa: E402

steps, reps = 24, int(sys.argv[1]) if len(sys.argv) > 1 else 30
ops.set_precision('bf16')
dev = torch.device('cuda', 0)
head = bench.make_weights(head_mod.KGDetHead()).to(dev).eval()
x, sc = bench.make_inputs(16, 0)
x_dev, sc_dev = x.to(dev), sc.to(dev)
g = head_mod.GraphedInference(head, x_dev, [bench.IMG_SHAPE] * 16, 0.05, 0.5, 1000, 100, score_override=sc_dev)
hosts = [(x + 0.01 * i).clone().pin_memory() for i in range(3)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
before = lambda i: flush.fill_(1)
ref = g.serve([hosts[i % 3] for i in range(3)], None, before_step=before, concurrent=False)
ref = [tuple(t.clone() for t in r) for r in ref]
bad = 0
for rep in range(reps):
    res = g.serve([hosts[i % 3] for i in range(steps)], None, before_step=before if rep % 2 == 0 else None,
                  concurrent=True)
    for i, r in enumerate(res):
        for ti, (p, q) in enumerate(zip(r, ref[i % 3])):
            if not torch.equal(p, q):
                dd = (p.double() - q.double()).abs()
                bad += 1
                if bad <= 12:
                    idx = (dd > 0).nonzero()
                    print('rep', rep, 'flush', rep % 2 == 0, 'batch', i, 'tensor', ti, tuple(p.shape), 'n_diff',
                          int((dd > 0).sum()), 'max', float(dd.max()), 'first idx', idx[0].tolist(), 'last idx',
                          idx[-1].tolist(), flush=True)
print('mismatching (batch, tensor) pairs:', bad, 'of', reps * steps * 3)
