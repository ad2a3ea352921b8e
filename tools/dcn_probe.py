"""Minimal launch target for ncu: a few bf16 fused-DCN forward calls at one KGDet shape.

    ncu --set full -k regex:dcn_umma_stream -s 2 -c 1 -o gpurun_out/prof python tools/dcn_probe.py 5
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kgdet_b200 import ops  # noqa: E402
from tests._data import dcn_case  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
d = dcn_case(N=16, C=256, H=25, W=42, Cout=256, k=k)
x, off, w = (d[q].cuda() for q in ('x', 'offset', 'weight'))
ops.set_precision('bf16')
with torch.no_grad():
    for _ in range(4):
        ops.deform_conv(x, off, w, 1, k // 2)
torch.cuda.synchronize()
