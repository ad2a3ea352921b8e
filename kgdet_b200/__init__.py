"""kgdet_b200 -- B200-native operators of the KGDet point-set detection head.

``kgdet_b200.ops`` mirrors ``mmdet.ops`` for the hot path; ``kgdet_b200.mount_as_mmdet_ops()``
registers it (and its dcn / nms / sigmoid_focal_loss sub-modules) in ``sys.modules`` under the
reference's names so that the unchanged heads (`from mmdet.ops import DeformConv`), FocalLoss
(`from mmdet.ops import sigmoid_focal_loss`) and multiclass_nms_kp
(`from mmdet.ops.nms import nms_wrapper`) resolve to this package.
"""
import sys

from . import ops  # noqa: F401

__version__ = '0.1.0'


def accelerate(ref_head, **inject):
    """Rebind an UNCHANGED reference head instance to the fused paths of this package (kgdet_b200/adopt.py)."""
    from .adopt import accelerate as _accelerate
    return _accelerate(ref_head, **inject)


def mount_as_mmdet_ops():
    """Make `import mmdet.ops` (and the sub-modules the reference imports) resolve to kgdet_b200.ops.

    Call before `import mmdet.models`.  Only the ops namespace is replaced; the rest of mmdet is
    the user's own (unchanged) installation.
    """
    import importlib
    # (the functions `nms` / `sigmoid_focal_loss` shadow their modules as attributes of the
    # package, exactly as in mmdet/ops/__init__.py -- fetch the modules through importlib)
    dcn = importlib.import_module(__name__ + '.ops.dcn')
    nms = importlib.import_module(__name__ + '.ops.nms')
    nms_wrapper = importlib.import_module(__name__ + '.ops.nms.nms_wrapper')
    sigmoid_focal_loss = importlib.import_module(__name__ + '.ops.sigmoid_focal_loss')
    sys.modules['mmdet.ops'] = ops
    sys.modules['mmdet.ops.dcn'] = dcn
    sys.modules['mmdet.ops.dcn.deform_conv'] = dcn
    sys.modules['mmdet.ops.nms'] = nms
    sys.modules['mmdet.ops.nms.nms_wrapper'] = nms_wrapper
    sys.modules['mmdet.ops.sigmoid_focal_loss'] = sigmoid_focal_loss
    mm = sys.modules.get('mmdet')
    if mm is not None:
        mm.ops = ops
    return ops
