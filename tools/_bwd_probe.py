import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kgdet_b200 import ops
from tests._data import dcn_case
k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
d = dcn_case(N=16, C=256, H=25, W=42, Cout=256, k=k)
x, off, w, go = (d[q].cuda() for q in ('x', 'offset', 'weight', 'grad_out'))
ops.set_precision('bf16')
for _ in range(3):
    xg, og, wg = x.clone().requires_grad_(), off.clone().requires_grad_(), w.clone().requires_grad_()
    ops.deform_conv(xg, og, wg, 1, k // 2).backward(go)
torch.cuda.synchronize()
