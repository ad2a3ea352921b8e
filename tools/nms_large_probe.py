import os, sys
sys.path.insert(0, os.getcwd())
import torch
from kgdet_b200.ops.nms import nms_wrapper
from tests._data import random_boxes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dets = random_boxes(n, seed=n, clustered=True).cuda()
for _ in range(2):
    nms_wrapper._nms_keep_cuda(dets, 0.5, 0)
torch.cuda.synchronize()
