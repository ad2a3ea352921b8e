"""Build kgdet_b200/_lib/libkgdet_b200.so with nvcc for sm_100a (in-tree, no torch headers).

    python -m kgdet_b200.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/kgdet_b200.h); the Python mirror loads
it with ctypes (kgdet_b200/ops/_capi.py).  Objects are cached per source under
kgdet_b200/_lib/obj and rebuilt when the source or a header is newer.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, '_lib')
LIB_PATH = os.path.join(LIB_DIR, 'libkgdet_b200.so')

SOURCES = ['runtime.cu', 'nms.cu', 'focal_loss.cu', 'point_loss.cu', 'moment.cu', 'dcn_common.cu', 'dcn_simt.cu',
           'dcn_umma.cu', 'dcn_umma_stream.cu', 'dcn_umma_group.cu', 'gemm_umma.cu', 'pointwise_umma.cu', 'conv_umma.cu', 'tower_nhwc.cu', 'decode.cu', 'dcn_bwd_tc.cu', 'dcn_col2im_own.cu', 'dcn_wgrad_umma.cu', 'dcn_api.cu']
HEADERS = [os.path.join(CSRC, 'common.cuh'), os.path.join(CSRC, 'focal.cuh'), os.path.join(CSRC, 'dcn.cuh'), os.path.join(CSRC, 'dcn_umma.cuh'),
           os.path.join(ROOT, 'include', 'kgdet_b200.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
              '--expt-relaxed-constexpr', '-Xcudafe', '--diag_suppress=177']


def _nvcc():
    home = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    return os.path.join(home, 'bin', 'nvcc')


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(os.path.join(LIB_DIR, 'obj'), exist_ok=True)
    jobs = []
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(LIB_DIR, 'obj', s + '.o')
        objs.append(obj)
        if force or _stale(obj, [src] + HEADERS):
            cmd = [_nvcc(), *NVCC_FLAGS, '-c', src, '-o', obj]
            if verbose:
                cmd.insert(1, '-Xptxas')
                cmd.insert(2, '-v')
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return cmd, r.returncode, r.stdout

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for cmd, rc, out in ex.map(run, jobs):
                if verbose or rc != 0:
                    print(' '.join(cmd))
                    print(out)
                if rc != 0:
                    raise RuntimeError('nvcc failed for %s' % cmd[-3])
    if jobs or force or _stale(LIB_PATH, objs):
        cmd = [_nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', *objs, '-o', LIB_PATH,
               '-Xlinker', '--no-undefined', '-lcudart']
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            print(r.stdout)
            raise RuntimeError('link failed')
    return LIB_PATH


if __name__ == '__main__':
    p = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(p)
