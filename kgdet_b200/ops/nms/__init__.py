from .nms_wrapper import batched_nms_flags, nms, soft_nms

__all__ = ['nms', 'soft_nms', 'batched_nms_flags']
