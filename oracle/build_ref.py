"""Compile the reference's own native sources, UNMODIFIED and where they lie under
/root/reference, into ``oracle/_ref/`` — TEST INFRASTRUCTURE ONLY.

Outputs (git-ignored, but they travel to the GPU box with the gpurun snapshot):

* ``oracle/_ref/nms_cpu.so``                 <- mmdet/ops/nms/src/nms_cpu.cpp
      the bit-exactness pin for NMS and the CPU baseline named in BASELINE.json
* ``oracle/_ref/deform_conv_cuda.so``        <- mmdet/ops/dcn/src/deform_conv_cuda.cpp + _kernel.cu
      (sm_100a)  GPU-side pin of oracle/dcn_oracle.py and the incumbent to beat
* ``oracle/_ref/sigmoid_focal_loss_cuda.so`` <- mmdet/ops/sigmoid_focal_loss/src/*.{cpp,cu}
      (sm_100a)  GPU-side pin of oracle/focal_oracle.py

No reference source is copied into the repo: the compilers read the files in
place.  The only additions are command-line macros (``-DAT_CHECK=TORCH_CHECK``:
torch >= 1.5 dropped AT_CHECK) and, for the focal loss, a 3-line ``THC/THC.h``
shim header written to ``oracle/_ref/shim/`` (torch >= 1.11 removed that header;
the reference only uses ``THCudaCheck`` and ``THCCeilDiv`` from it).
``nms_cuda`` does not compile on torch 2.x (``nms_kernel.cu:83`` uses the
removed THCState API) and is not built; its comparator ('>') is restated in
``oracle/nms_oracle.c``.

Usage: ``python -m oracle.build_ref [--cpu-only]``; ``load(name)`` imports a built
module (or returns None when it has not been built).
"""
import importlib.util
import os
import subprocess
import sys
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, '_ref')
REF_ROOT = os.environ.get('KGDET_REFERENCE_ROOT', '/root/reference')
OPS = os.path.join(REF_ROOT, 'mmdetection', 'mmdet', 'ops')

_THC_SHIM = """#pragma once
#include <c10/cuda/CUDAException.h>
#define THCudaCheck(x) C10_CUDA_CHECK(x)
template <class T> __host__ __device__ inline T THCCeilDiv(T a, T b) { return (a + b - 1) / b; }
"""


def _torch_flags():
    import torch
    from torch.utils import cpp_extension as ce
    inc = []
    for p in ce.include_paths():
        inc += ['-isystem', p]
    inc += ['-isystem', sysconfig.get_paths()['include']]
    libdir = os.path.join(os.path.dirname(torch.__file__), 'lib')
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    return inc, libdir, abi


def _run(cmd):
    print('[build_ref]', ' '.join(cmd[:6]), '...', flush=True)
    subprocess.check_call(cmd)


def _stale(out, srcs):
    return (not os.path.exists(out)) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs)


def build_nms_cpu():
    src = os.path.join(OPS, 'nms', 'src', 'nms_cpu.cpp')
    out = os.path.join(REF_DIR, 'nms_cpu.so')
    if not _stale(out, [src]):
        return out
    inc, libdir, abi = _torch_flags()
    os.makedirs(REF_DIR, exist_ok=True)
    _run(['g++', '-O3', '-std=c++17', '-fPIC', '-shared', '-w',
          '-DTORCH_EXTENSION_NAME=nms_cpu', '-DTORCH_API_INCLUDE_EXTENSION_H',
          f'-D_GLIBCXX_USE_CXX11_ABI={abi}', *inc, src, '-o', out,
          f'-L{libdir}', f'-Wl,-rpath,{libdir}', '-lc10', '-ltorch_cpu', '-ltorch', '-ltorch_python'])
    return out


def _build_cuda_ext(name, cpp_srcs, cu_srcs, extra_inc=(), nvcc_extra=()):
    out = os.path.join(REF_DIR, name + '.so')
    if not _stale(out, cpp_srcs + cu_srcs):
        return out
    inc, libdir, abi = _torch_flags()
    cuda_home = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    inc = inc + ['-isystem', os.path.join(cuda_home, 'include')]
    for p in extra_inc:
        inc = ['-I', p] + inc
    os.makedirs(os.path.join(REF_DIR, 'obj'), exist_ok=True)
    defs = [f'-DTORCH_EXTENSION_NAME={name}', '-DTORCH_API_INCLUDE_EXTENSION_H',
            f'-D_GLIBCXX_USE_CXX11_ABI={abi}', '-DAT_CHECK=TORCH_CHECK']
    objs = []
    for s in cpp_srcs:
        o = os.path.join(REF_DIR, 'obj', name + '_' + os.path.basename(s) + '.o')
        _run(['g++', '-O3', '-std=c++17', '-fPIC', '-w', '-c', *defs, *inc, s, '-o', o])
        objs.append(o)
    for s in cu_srcs:
        o = os.path.join(REF_DIR, 'obj', name + '_' + os.path.basename(s) + '.o')
        _run([os.path.join(cuda_home, 'bin', 'nvcc'), '-O3', '-std=c++17', '-w',
              '-gencode', 'arch=compute_100a,code=sm_100a',
              '-D__CUDA_NO_HALF_OPERATORS__', '-D__CUDA_NO_HALF_CONVERSIONS__',
              '-D__CUDA_NO_HALF2_OPERATORS__', '--expt-relaxed-constexpr',
              '-Xcompiler', '-fPIC', *nvcc_extra, '-c', *defs, *inc, s, '-o', o])
        objs.append(o)
    _run(['g++', '-shared', *objs, '-o', out, f'-L{libdir}', f'-Wl,-rpath,{libdir}',
          f'-L{cuda_home}/lib64', '-lc10', '-lc10_cuda', '-ltorch_cpu', '-ltorch_cuda',
          '-ltorch', '-ltorch_python', '-lcudart'])
    return out


def build_deform_conv_cuda():
    d = os.path.join(OPS, 'dcn', 'src')
    return _build_cuda_ext('deform_conv_cuda', [os.path.join(d, 'deform_conv_cuda.cpp')],
                           [os.path.join(d, 'deform_conv_cuda_kernel.cu')])


def build_deform_conv_cuda_r64():
    """The same unmodified sources with `-maxrregcount=64`: the reference launches 1024 threads per block
    (deform_conv_cuda_kernel.cu:71-81), which its float64 instantiations cannot do at ptxas' default register
    allocation on sm_100a ("too many resources requested for launch" -- the kernel silently does not run).  This
    build is the float64 TRUTH of the full-size parity tests only; the float32 incumbent is the default build."""
    d = os.path.join(OPS, 'dcn', 'src')
    return _build_cuda_ext('deform_conv_cuda_r64', [os.path.join(d, 'deform_conv_cuda.cpp')],
                           [os.path.join(d, 'deform_conv_cuda_kernel.cu')], nvcc_extra=['-maxrregcount=64'])


def build_sigmoid_focal_loss_cuda():
    d = os.path.join(OPS, 'sigmoid_focal_loss', 'src')
    shim = os.path.join(REF_DIR, 'shim')
    os.makedirs(os.path.join(shim, 'THC'), exist_ok=True)
    with open(os.path.join(shim, 'THC', 'THC.h'), 'w') as f:
        f.write(_THC_SHIM)
    return _build_cuda_ext('sigmoid_focal_loss_cuda', [os.path.join(d, 'sigmoid_focal_loss.cpp')],
                           [os.path.join(d, 'sigmoid_focal_loss_cuda.cu')], extra_inc=[shim])


def reference_available():
    return os.path.isdir(OPS)


PYTREE_ZIP = os.path.join(REF_DIR, 'pytree.zip')


def stage_python_tree():
    """Archive the reference's PYTHON package (mmdetection/mmdet/**/*.py) and its four configs, untouched, into
    the git-ignored ``oracle/_ref/pytree.zip`` so that the UNCHANGED reference heads can be imported (zipimport)
    and executed on the GPU box, where /root/reference does not exist -- the built .so files above travel the
    same way.  Nothing is edited; nothing enters the repository (``oracle/_ref/`` is in .gitignore)."""
    import zipfile
    src_pkg = os.path.join(REF_ROOT, 'mmdetection', 'mmdet')
    cfg_src = os.path.join(REF_ROOT, 'configs')
    files = []
    for root, dirs, names in os.walk(src_pkg):
        dirs[:] = sorted(d for d in dirs if d != '__pycache__')
        for f in sorted(names):
            if f.endswith('.py'):
                s = os.path.join(root, f)
                files.append((s, os.path.join('mmdetection', 'mmdet', os.path.relpath(s, src_pkg))))
    for f in sorted(os.listdir(cfg_src)):
        if f.endswith('.py'):
            files.append((os.path.join(cfg_src, f), os.path.join('configs', f)))
    if not _stale(PYTREE_ZIP, [s for s, _ in files]):
        return PYTREE_ZIP
    os.makedirs(REF_DIR, exist_ok=True)
    with zipfile.ZipFile(PYTREE_ZIP, 'w', zipfile.ZIP_DEFLATED) as z:
        for s, arc in files:
            z.write(s, arc)
    return '%s (%d files)' % (PYTREE_ZIP, len(files))


def build_all(cpu_only=False):
    """Build whatever the local reference tree allows.  Returns {name: path|error}."""
    res = {}
    if not reference_available():
        return res
    steps = [('nms_cpu', build_nms_cpu), ('pytree', stage_python_tree)]
    if not cpu_only:
        steps += [('deform_conv_cuda', build_deform_conv_cuda), ('deform_conv_cuda_r64', build_deform_conv_cuda_r64),
                  ('sigmoid_focal_loss_cuda', build_sigmoid_focal_loss_cuda)]
    for name, fn in steps:
        try:
            res[name] = fn()
        except Exception as e:  # recorded, not fatal: the C/py restatement is always there
            res[name] = 'FAILED: %r' % (e,)
    return res


def load(name):
    """Import oracle/_ref/<name>.so as a Python module; None if it was not built."""
    path = os.path.join(REF_DIR, name + '.so')
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    if name in sys.modules and getattr(sys.modules[name], '__file__', None) == path:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


if __name__ == '__main__':
    r = build_all(cpu_only='--cpu-only' in sys.argv)
    for k, v in r.items():
        print(k, '->', v)
