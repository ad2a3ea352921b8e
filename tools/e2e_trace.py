import os, sys, json, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from kgdet_b200 import ops, head as head_mod
ops.set_precision('bf16')
dev = torch.device('cuda', 0)
head = bench.make_weights(head_mod.KGDetHead()).to(dev).eval()
x, sc = bench.make_inputs(16, 0)
x_dev, sc_dev = x.to(dev), sc.to(dev)
shapes = [bench.IMG_SHAPE] * 16
g = head_mod.GraphedInference(head, x_dev, shapes, 0.05, 0.5, 1000, 100, score_override=sc_dev)
hosts = [x.clone().pin_memory() for _ in range(3)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
outs = None
for _ in range(2):
    g.serve([hosts[i % 3] for i in range(8)], None, before_step=lambda i: flush.fill_(1))
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g.serve([hosts[i % 3] for i in range(10)], None, before_step=lambda i: flush.fill_(1))
    torch.cuda.synchronize()
prof.export_chrome_trace('gpurun_out/e2e_trace.json')
print('ok')
