"""Replays of ONE captured inference step while an unrelated stream keeps the GPU busy (fills and matmuls of varying
size): the outputs must not depend on what else runs.  A mismatch means a missing dependency between parallel
branches INSIDE the captured graph (it cannot come from sharing: nothing else touches the graph's memory)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from kgdet_b200 import head as head_mod  # noqa: E402
from kgdet_b200 import ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ops.set_precision('bf16')
dev = torch.device('cuda', 0)
head = bench.make_weights(head_mod.KGDetHead()).to(dev).eval()
x, sc = bench.make_inputs(16, 0)
x_dev, sc_dev = x.to(dev), sc.to(dev)
g = head_mod.GraphedInference(head, x_dev, [bench.IMG_SHAPE] * 16, 0.05, 0.5, 1000, 100, score_override=sc_dev)
ref = [t.clone() for t in g(x_dev)]
torch.cuda.synchronize()
noise = torch.cuda.Stream()
buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
a = torch.randn(2048, 2048, device=dev)
gen = torch.Generator().manual_seed(0)
bad = 0
for rep in range(reps):
    kind = rep % 4
    with torch.cuda.stream(noise):
        for _ in range(int(torch.randint(1, 6, (1,), generator=gen))):
            if kind == 0:
                buf[: int(torch.randint(1 << 20, 256 << 20, (1,), generator=gen))].fill_(1)
            elif kind == 1:
                a @ a
            elif kind == 2:
                buf.fill_(2); a @ a
            # kind 3: quiet
    out = g(x_dev)
    torch.cuda.synchronize()
    for ti, (p, q) in enumerate(zip(out, ref)):
        if not torch.equal(p, q):
            bad += 1
            if bad <= 10:
                dd = (p.double() - q.double()).abs()
                print('rep', rep, 'noise kind', kind, 'tensor', ti, 'n_diff', int((dd > 0).sum()), 'max', float(dd.max()),
                      flush=True)
print('mismatching (replay, tensor) pairs:', bad, 'of', reps * 3)
