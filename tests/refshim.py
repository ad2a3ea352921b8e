"""Import shim that lets the UNCHANGED reference package (`/root/reference/mmdetection/mmdet`)
be imported in this container -- TEST INFRASTRUCTURE ONLY.

The reference depends on third-party modules that are absent here (mmcv==0.2.13, pycocotools,
terminaltables, imagecorruptions, matplotlib) and on a generated `mmdet/version.py`.  This module
registers minimal stand-ins in `sys.modules`, aliases `collections.Sequence` (removed in Python
3.10), mounts an `mmdet.ops` implementation of the caller's choice, and puts the reference tree
on `sys.path`.  Nothing is copied from the reference; its files are imported where they lie.

    install(ops='kgdet')   -> mmdet.ops = kgdet_b200.ops      (the drop-in under test)
    install(ops='oracle')  -> mmdet.ops = tests/oracle_ops.py  (CPU oracle, for golden vectors)
"""
import collections
import collections.abc
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF_ROOT = os.environ.get('KGDET_REFERENCE_ROOT', '/root/reference')
MMDET_PARENT = os.path.join(REF_ROOT, 'mmdetection')
_ZIP = None
if not os.path.isdir(os.path.join(MMDET_PARENT, 'mmdet')):
    # GPU box: the reference tree is absent; oracle/build_ref.py archived its Python package + configs, untouched,
    # into the git-ignored oracle/_ref/pytree.zip (it travels with the gpurun snapshot); imported by zipimport
    _z = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref', 'pytree.zip')
    if os.path.isfile(_z):
        _ZIP = _z
        MMDET_PARENT = os.path.join(_z, 'mmdetection')


def available():
    return _ZIP is not None or os.path.isdir(os.path.join(MMDET_PARENT, 'mmdet'))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


# ---- mmcv.cnn initialisers (mmcv 0.2.13 semantics; tolerate bias-free modules) -----------------
def normal_init(module, mean=0, std=1, bias=0):
    nn.init.normal_(module.weight, mean, std)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    if distribution == 'uniform':
        nn.init.xavier_uniform_(module.weight, gain=gain)
    else:
        nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def kaiming_init(module, mode='fan_out', nonlinearity='relu', bias=0, distribution='normal'):
    if distribution == 'uniform':
        nn.init.kaiming_uniform_(module.weight, mode=mode, nonlinearity=nonlinearity)
    else:
        nn.init.kaiming_normal_(module.weight, mode=mode, nonlinearity=nonlinearity)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def caffe2_xavier_init(module, bias=0):
    kaiming_init(module, mode='fan_in', nonlinearity='leaky_relu', distribution='uniform')


def obj_from_dict(info, parent=None, default_args=None):
    args = info.copy()
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_type = getattr(parent, obj_type) if parent is not None else sys.modules[obj_type]
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    return obj_type(**args)


class _Dummy(object):
    def __init__(self, *a, **k):
        pass


class AttrDict(dict):
    """cfg.uniform.assigner / cfg.score_thr / cfg.get('nms_pre') access (KP3:711,864,909)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v

    def copy(self):
        return AttrDict(dict.copy(self))


def load_config(name):
    """exec a reference config file (they are plain Python) -> dict of its globals."""
    path = os.path.join(REF_ROOT, 'configs', name)
    ns = {}
    if _ZIP is not None:
        import zipfile
        with zipfile.ZipFile(_ZIP) as z:
            src = z.read('configs/' + name).decode()
    else:
        with open(path) as f:
            src = f.read()
    exec(compile(src, path, 'exec'), ns)
    return {k: v for k, v in ns.items() if not k.startswith('__')}


_installed = None


def install(ops='kgdet'):
    """Register the stand-ins and mount `mmdet.ops`.  Returns the mounted ops module."""
    global _installed
    if not available():
        raise RuntimeError('reference tree not found at %s' % REF_ROOT)
    for n in ('Sequence', 'Mapping', 'Iterable', 'MutableMapping'):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))
    if _installed is None:
        _mod('mmcv', is_str=lambda x: isinstance(x, str), is_list_of=lambda s, t: all(isinstance(i, t) for i in s),
             imread=None, imwrite=None)
        _mod('mmcv.cnn', normal_init=normal_init, constant_init=constant_init, kaiming_init=kaiming_init,
             xavier_init=xavier_init, VGG=type('VGG', (nn.Module,), {}))
        _mod('mmcv.cnn.weight_init', normal_init=normal_init, xavier_init=xavier_init,
             caffe2_xavier_init=caffe2_xavier_init, constant_init=constant_init, kaiming_init=kaiming_init)
        _mod('mmcv.runner', obj_from_dict=obj_from_dict, Hook=_Dummy, OptimizerHook=_Dummy,
             get_dist_info=lambda: (0, 1), load_checkpoint=lambda *a, **k: None, Runner=_Dummy,
             DistSamplerSeedHook=_Dummy)
        _mod('mmcv.runner.utils', get_dist_info=lambda: (0, 1))
        _mod('mmcv.parallel', DataContainer=_Dummy, collate=lambda *a, **k: None, scatter=lambda *a, **k: None,
             MMDataParallel=_Dummy, MMDistributedDataParallel=_Dummy)
        sys.modules['mmcv'].cnn = sys.modules['mmcv.cnn']
        sys.modules['mmcv'].runner = sys.modules['mmcv.runner']
        sys.modules['mmcv'].parallel = sys.modules['mmcv.parallel']
        _mod('pycocotools')
        _mod('pycocotools.mask')
        _mod('pycocotools.coco', COCO=_Dummy)
        _mod('pycocotools.cocoeval', COCOeval=_Dummy)
        _mod('terminaltables', AsciiTable=_Dummy)
        _mod('imagecorruptions', corrupt=lambda *a, **k: None)
        if 'matplotlib' not in sys.modules:
            try:
                import matplotlib  # noqa: F401
            except Exception:
                _mod('matplotlib')
                _mod('matplotlib.pyplot')
        _mod('mmdet.version', __version__='1.0rc0+kgdet', short_version='1.0rc0')
        if MMDET_PARENT not in sys.path:
            sys.path.insert(0, MMDET_PARENT)
    if ops == 'kgdet':
        import kgdet_b200
        mounted = kgdet_b200.mount_as_mmdet_ops()
    elif ops == 'oracle':
        from tests import oracle_ops
        mounted = oracle_ops.mount()
    else:
        raise ValueError(ops)
    if _installed is not None and _installed != ops:
        # re-point already-imported reference modules at the newly mounted ops
        for name, m in list(sys.modules.items()):
            if name.startswith('mmdet.') and not name.startswith('mmdet.ops') and m is not None:
                for attr in ('DeformConv', 'ModulatedDeformConv', 'nms', 'nms_wrapper', '_sigmoid_focal_loss'):
                    if hasattr(m, attr):
                        if attr == 'nms_wrapper':
                            setattr(m, attr, sys.modules['mmdet.ops.nms.nms_wrapper'])
                        elif attr == '_sigmoid_focal_loss':
                            setattr(m, attr, mounted.sigmoid_focal_loss)
                        else:
                            setattr(m, attr, getattr(mounted, attr))
    _installed = ops
    return mounted


def patch_point_generator_device(device):
    """PointGenerator.grid_points / valid_flags default to device='cuda'
    (mmdet/core/anchor/point_generator.py:14,24); off-GPU the default is re-pointed (defaults only,
    the method bodies are the reference's)."""
    from mmdet.core.anchor.point_generator import PointGenerator
    for name in ('grid_points', 'valid_flags'):
        fn = getattr(PointGenerator, name)
        d = list(fn.__defaults__)
        d[-1] = device
        fn.__defaults__ = tuple(d)


def patch_point_assigner_device():
    """`PointAssigner.assign` builds `torch.arange(num_points)` without a device and indexes it with masks that live
    on the points' device (mmdet/core/bbox/assigners/point_assigner.py:70-76) -- fine on torch 1.x, an error on
    torch 2.x when the points are CUDA tensors.  The method body stays the reference's; it merely runs with the
    points' device as the default device for factory calls."""
    from mmdet.core.bbox.assigners.point_assigner import PointAssigner
    if getattr(PointAssigner.assign, '_kgdet_device_patch', False):
        return
    orig = PointAssigner.assign

    def assign(self, points, *args, **kwargs):
        with torch.device(points.device):
            return orig(self, points, *args, **kwargs)
    assign._kgdet_device_patch = True
    PointAssigner.assign = assign


def build_head(config_name, device='cpu'):
    """Build the bbox_head of a reference config with the reference's own builder."""
    cfg = load_config(config_name)
    from mmdet.models import build_head as _build_head
    patch_point_generator_device(device)
    head = _build_head(cfg['model']['bbox_head'])
    head.init_weights()
    return head.to(device), cfg
