"""Two captured instances of the inference step replayed AT THE SAME TIME on two streams (both wait for one event),
results compared with a quiet replay; repeated for head variants with one fast path switched off each, to find which
kernel is not safe next to a twin of itself.

    python tools/lockstep_stress.py [trials]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from kgdet_b200 import head as head_mod  # noqa: E402
from kgdet_b200 import ops  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
only = sys.argv[2] if len(sys.argv) > 2 else None
ops.set_precision('bf16')
dev = torch.device('cuda', 0)
x, sc = bench.make_inputs(16, 0)
x_dev, sc_dev = x.to(dev), sc.to(dev)
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()


def blocks(head):
    return [m for m in head.modules() if hasattr(m, 'grouped_dcn')]


def variant(name):
    head = bench.make_weights(head_mod.KGDetHead()).to(dev).eval()
    if name == 'cudnn_towers':
        head._own_convs = False
    elif name == 'torch_decode':
        head._fused_decode = False
    elif name == 'six_stream_dcn':
        for b in blocks(head):
            b.grouped_dcn = False
    elif name == 'serial_dcn':
        for b in blocks(head):
            b.grouped_dcn = False
            b.concurrent_dcn = False
    return head


for name in ('default', 'cudnn_towers', 'torch_decode', 'six_stream_dcn', 'serial_dcn'):
    if only and name != only:
        continue
    head = variant(name)
    g = head_mod.GraphedInference(head, x_dev, [bench.IMG_SHAPE] * 16, 0.05, 0.5, 1000, 100, score_override=sc_dev)
    t = g._second_instance()
    ref = [o.clone() for o in g()]
    torch.cuda.synchronize()
    bad = [0, 0]
    first = None
    for trial in range(trials):
        ev = torch.cuda.current_stream().record_event()
        for st, inst in ((s0, g), (s1, t)):
            st.wait_event(ev)
            with torch.cuda.stream(st):
                torch.cuda._sleep(200000 + 1000 * (trial % 7))       # both replays are queued before either starts
                inst.graph.replay()
        torch.cuda.synchronize()
        for k, inst in enumerate((g, t)):
            diff = [ti for ti, (p, q) in enumerate(zip(inst.static_out, ref)) if not torch.equal(p, q)]
            if diff:
                bad[k] += 1
                if first is None:
                    p, q = inst.static_out[diff[0]], ref[diff[0]]
                    dd = (p.double() - q.double()).abs()
                    first = 'trial %d instance %d tensors %s n_diff %d max %.3f' % (trial, k, diff, int((dd > 0).sum()),
                                                                                   float(dd.max()))
    print(name, 'mismatching replays (first instance, twin):', bad, 'of', trials, '|', first, flush=True)
    del g, t, head
    torch.cuda.empty_cache()
