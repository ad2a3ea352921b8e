"""Names of ``mmdet.ops`` that are outside the KGDet point-set head hot path (SURVEY.md section 2,
rows 15-16: RoI ops of two-stage detectors, MaskedConv2d of GA-RetinaNet, GCNet's ContextBlock).
They are exported so that ``from mmdet.ops import ...`` lines elsewhere in the reference keep
importing (mmdet/ops/__init__.py:12-19), and raise on use."""
import torch.nn as nn


def _unavailable(name):
    def fn(*args, **kwargs):
        raise NotImplementedError(
            '%s is outside the kgdet_b200 hot path (KGDet / RepPoints-Kp heads do not use it)' % name)
    fn.__name__ = name
    return fn


def _unavailable_module(name):
    class _M(nn.Module):
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(
                '%s is outside the kgdet_b200 hot path (KGDet / RepPoints-Kp heads do not use it)' % name)
    _M.__name__ = name
    return _M


RoIAlign = _unavailable_module('RoIAlign')
RoIPool = _unavailable_module('RoIPool')
DeformRoIPooling = _unavailable_module('DeformRoIPooling')
DeformRoIPoolingPack = _unavailable_module('DeformRoIPoolingPack')
ModulatedDeformRoIPoolingPack = _unavailable_module('ModulatedDeformRoIPoolingPack')
MaskedConv2d = _unavailable_module('MaskedConv2d')
ContextBlock = _unavailable_module('ContextBlock')
roi_align = _unavailable('roi_align')
roi_pool = _unavailable('roi_pool')
deform_roi_pooling = _unavailable('deform_roi_pooling')
