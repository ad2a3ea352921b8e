"""One forward + backward of a KGDet deformable convolution in bf16 mode (for ncu launch lists).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/dcn_bwd_probe.py [k] [N]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kgdet_b200 import ops  # noqa: E402
from tests._data import dcn_case  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 7
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
d = dcn_case(N=N, C=256, H=25, W=42, Cout=256, k=k)
x, off, w, go = (d[q].cuda() for q in ('x', 'offset', 'weight', 'grad_out'))
ops.set_precision('bf16')
for it in range(3):
    xg, og, wg = x.clone().requires_grad_(), off.clone().requires_grad_(), w.clone().requires_grad_()
    torch.cuda.synchronize()
    if it == 2:
        torch.zeros(1, device='cuda').fill_(7.0)          # marker launch in the list
    ops.deform_conv(xg, og, wg, 1, k // 2).backward(go)
torch.cuda.synchronize()
