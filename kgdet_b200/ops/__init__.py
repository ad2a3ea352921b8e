"""Drop-in for ``mmdet.ops`` (mmdet/ops/__init__.py:1-19) restricted to the KGDet hot path.

In scope (B200-native, C-ABI library): DeformConv*/ModulatedDeformConv*/deform_conv/
modulated_deform_conv, nms, SigmoidFocalLoss/sigmoid_focal_loss, plus the fused extras
points2bbox_moment, sigmoid_focal_loss_sum, batched_nms_flags.  Everything else imports and
raises NotImplementedError on use.
"""
from ._out_of_scope import (ContextBlock, DeformRoIPooling, DeformRoIPoolingPack, MaskedConv2d,
                            ModulatedDeformRoIPoolingPack, RoIAlign, RoIPool, deform_roi_pooling,
                            roi_align, roi_pool)
from .dcn import (DeformConv, DeformConvFunction, DeformConvPack, ModulatedDeformConv,
                  ModulatedDeformConvFunction, ModulatedDeformConvPack, deform_conv,
                  deform_conv_prepared, deform_conv_prepared_group, get_precision, modulated_deform_conv, prepare_input,
                  prepare_plan, prepare_plan_points, set_precision)
from .pointwise import (TiledRows, groupnorm_relu_nhwc, invalidate_weight_caches, nchw_to_tiled, pack_weight,
                        pointwise_conv)
from .decode import bbox_decode, bbox_finalize, bbox_select
from .moment import points2bbox_moment
from .nms import batched_nms_flags, nms, soft_nms
from .sigmoid_focal_loss import SigmoidFocalLoss, sigmoid_focal_loss, sigmoid_focal_loss_sum

__all__ = [
    'nms', 'soft_nms', 'RoIAlign', 'roi_align', 'RoIPool', 'roi_pool',
    'DeformConv', 'DeformConvPack', 'DeformRoIPooling', 'DeformRoIPoolingPack',
    'ModulatedDeformRoIPoolingPack', 'ModulatedDeformConv',
    'ModulatedDeformConvPack', 'deform_conv', 'modulated_deform_conv',
    'deform_roi_pooling', 'SigmoidFocalLoss', 'sigmoid_focal_loss',
    'MaskedConv2d', 'ContextBlock',
    # extras beyond the reference surface
    'DeformConvFunction', 'ModulatedDeformConvFunction', 'points2bbox_moment',
    'sigmoid_focal_loss_sum', 'batched_nms_flags', 'set_precision', 'get_precision',
    'prepare_input', 'prepare_plan', 'prepare_plan_points', 'deform_conv_prepared', 'deform_conv_prepared_group',
    'pointwise_conv',
    'nchw_to_tiled', 'pack_weight', 'TiledRows', 'groupnorm_relu_nhwc', 'bbox_select', 'bbox_decode',
    'bbox_finalize', 'invalidate_weight_caches',
]
