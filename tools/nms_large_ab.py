import json, os, sys
sys.path.insert(0, os.getcwd())
import torch
from kgdet_b200.ops.nms import nms_wrapper
from tests._data import random_boxes
from tools.op_microbench import timed
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for clustered in (True, False):
    for n in (16384, 32768, 65536):
        dets = random_boxes(n, seed=n, clustered=clustered).cuda()
        us = timed(lambda: nms_wrapper._nms_keep_cuda(dets, 0.5, 0), flush, reps=5)
        cb = (n + 63) // 64
        b = 20 * n + 2 * n * cb * 8
        keep = nms_wrapper._nms_keep_cuda(dets, 0.5, 0).numel()
        print(json.dumps(dict(pipelined=os.environ.get('KGDET_NMS_SWEEP_PIPELINED', '1'), n=n, clustered=clustered, kept=keep, us=round(us, 1), alg_MB=round(b / 1e6, 1), GBps=round(b / us / 1e3, 1))), flush=True)
