"""Normalised error of each tensor-core mode against the exact fp32 path at the KGDet shapes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kgdet_b200 import ops
from tests._data import dcn_case, rel_err

for k in (3, 5, 7):
    d = dcn_case(N=16, C=256, H=25, W=42, Cout=256, k=k, seed=k)
    x, off, w = (d[q].cuda() for q in ('x', 'offset', 'weight'))
    ops.set_precision('fp32')
    ref = ops.deform_conv(x, off, w, 1, k // 2)
    row = dict(k=k)
    for prec in ('bf16', 'tf32x3', 'tf32'):
        ops.set_precision(prec)
        row[prec] = float('%.3g' % rel_err(ops.deform_conv(x, off, w, 1, k // 2), ref))
    print(json.dumps(row), flush=True)
