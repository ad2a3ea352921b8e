"""ctypes front-end of ``oracle/nms_oracle.c`` — TEST INFRASTRUCTURE ONLY.

Restates ``mmdet/ops/nms/nms_wrapper.py:8-49`` on top of the C restatement:
``nms(dets, iou_thr) -> (dets[inds], inds)`` for a CPU tensor or ndarray.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, '_build')
_SO = os.path.join(_BUILD, 'libnms_oracle.so')
_lib = None


def build(force=False):
    """gcc-compile the C restatement (outputs under oracle/_build/, git-ignored)."""
    src = os.path.join(_HERE, 'nms_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-shared', '-fPIC',
                               src, '-o', _SO])
    return _SO


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.kgdet_oracle_nms.restype = ctypes.c_int64
        lib.kgdet_oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_float,
                                         ctypes.c_int, ctypes.c_void_p]
        _lib = lib
    return _lib


def nms_keep(dets, iou_thr, cmp_mode=1):
    """dets [n,5] float32 (tensor/ndarray, CPU) -> int64 ndarray of kept indices, ascending.

    cmp_mode 1 = '>=' (nms_cpu.cpp:55), 0 = '>' (nms_kernel.cu:60).
    """
    a = dets.detach().cpu().numpy() if isinstance(dets, torch.Tensor) else np.asarray(dets)
    a = np.ascontiguousarray(a, dtype=np.float32)
    n = a.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    m = _load().kgdet_oracle_nms(a.ctypes.data, n, float(iou_thr), int(cmp_mode),
                                 keep.ctypes.data)
    return keep[:m].copy()


def nms(dets, iou_thr, cmp_mode=1):
    """Mirror of nms_wrapper.nms for CPU inputs (nms_wrapper.py:26-49)."""
    inds = nms_keep(dets, iou_thr, cmp_mode)
    if isinstance(dets, torch.Tensor):
        inds_t = torch.from_numpy(inds)
        return dets[inds_t, :], inds_t
    return dets[inds, :], inds
