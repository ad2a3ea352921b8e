"""ctypes binding of include/kgdet_b200.h (the C-ABI boundary).

PyTorch is only plumbing here: it owns device memory and streams; every compute call goes
through ``libkgdet_b200.so``.  There is no CPU fallback and no alternative backend: if the
library is missing or a call fails, a ``RuntimeError`` is raised.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, '_lib', 'libkgdet_b200.so')

# enums of include/kgdet_b200.h
ABI_VERSION = 2
F32, BF16 = 0, 1
PREC_FP32, PREC_TF32X3, PREC_BF16, PREC_TF32 = 0, 1, 2, 3
NMS_GT, NMS_GE = 0, 1
PRECISIONS = {'fp32': PREC_FP32, 'tf32x3': PREC_TF32X3, 'bf16': PREC_BF16, 'tf32': PREC_TF32}

c_i32, c_f32, c_sz, c_ptr = ctypes.c_int32, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p


class DcnShape(ctypes.Structure):
    """struct kgdet_dcn_shape"""
    _fields_ = [(n, c_i32) for n in (
        'N', 'C', 'H', 'W', 'Cout', 'kh', 'kw', 'stride_h', 'stride_w', 'pad_h', 'pad_w',
        'dil_h', 'dil_w', 'groups', 'deformable_groups')]


_SHAPE_P = ctypes.POINTER(DcnShape)


class PointwiseSegment(ctypes.Structure):
    """struct kgdet_pointwise_segment"""
    _fields_ = [('out', ctypes.c_void_p), ('residual', ctypes.c_void_p), ('col_begin', c_i32), ('col_end', c_i32),
                ('channels_total', c_i32), ('channel_offset', c_i32)]


LAYOUT_NCHW, LAYOUT_TILED, LAYOUT_TILED_SPLIT = 0, 1, 2
DCN_GROUP_MAX = 6


class DcnGroupItem(ctypes.Structure):
    """struct kgdet_dcn_group_item"""
    _fields_ = [('prepared_input', ctypes.c_void_p), ('plan', ctypes.c_void_p), ('weight_packed', ctypes.c_void_p),
                ('bias', ctypes.c_void_p), ('output', ctypes.c_void_p), ('out_channel_offset', c_i32),
                ('out_channels_total', c_i32), ('fuse_relu', c_i32), ('out_layout', c_i32), ('dtype', c_i32),
                ('shape', DcnShape)]


# name -> (restype, argtypes); must list every KGDET_API symbol of the header
SIGNATURES = {
    'kgdet_last_error': (ctypes.c_char_p, []),
    'kgdet_abi_version': (ctypes.c_int, []),
    'kgdet_launch_count': (ctypes.c_uint64, []),
    'kgdet_dcn_fast_path_supported': (ctypes.c_int, [_SHAPE_P, ctypes.c_int]),
    'kgdet_dcn_packed_weight_bytes': (c_sz, [_SHAPE_P, ctypes.c_int]),
    'kgdet_dcn_pack_weight': (ctypes.c_int, [c_ptr, c_ptr, _SHAPE_P, ctypes.c_int, c_ptr]),
    'kgdet_dcn_forward_workspace_bytes': (c_sz, [_SHAPE_P, ctypes.c_int, ctypes.c_int]),
    'kgdet_dcn_forward': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, _SHAPE_P,
                                         ctypes.c_int, ctypes.c_int, c_ptr, c_sz, c_ptr]),
    'kgdet_dcn_prepared_input_bytes': (c_sz, [_SHAPE_P, ctypes.c_int]),
    'kgdet_dcn_prepare_input': (ctypes.c_int, [c_ptr, c_ptr, _SHAPE_P, ctypes.c_int, ctypes.c_int, c_ptr]),
    'kgdet_dcn_prepare_input_rows': (ctypes.c_int, [c_ptr, c_ptr, _SHAPE_P, ctypes.c_int, c_ptr]),
    'kgdet_dcn_plan_bytes': (c_sz, [_SHAPE_P, ctypes.c_int]),
    'kgdet_dcn_prepare_plan': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, _SHAPE_P, ctypes.c_int, c_ptr]),
    'kgdet_dcn_prepare_plan_points': (ctypes.c_int, [c_ptr, c_i32, c_i32, c_f32, c_f32, c_ptr, _SHAPE_P, ctypes.c_int,
                                                     c_ptr]),
    'kgdet_dcn_forward_prepared': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32, ctypes.c_int,
                                                  ctypes.c_int, _SHAPE_P, ctypes.c_int, ctypes.c_int, c_ptr, c_sz,
                                                  c_ptr]),
    'kgdet_dcn_forward_prepared_workspace_bytes': (c_sz, [_SHAPE_P, ctypes.c_int]),
    'kgdet_dcn_group_supported': (ctypes.c_int, [_SHAPE_P, ctypes.c_int]),
    'kgdet_dcn_forward_prepared_group': (ctypes.c_int, [c_ptr, c_i32, ctypes.c_int, c_ptr, c_sz, c_ptr]),
    'kgdet_dcn_group_set_profile_events': (None, [c_ptr, c_ptr]),
    'kgdet_bbox_select': (ctypes.c_int, [c_ptr, ctypes.c_int, c_i32, c_i32, c_i32, c_i32, c_ptr, c_ptr]),
    'kgdet_bbox_select_workspace_bytes': (ctypes.c_size_t, [c_i32, c_i32, c_i32]),
    'kgdet_bbox_select_ws': (ctypes.c_int, [c_ptr, ctypes.c_int, c_i32, c_i32, c_i32, c_i32, c_ptr, c_ptr, ctypes.c_size_t,
                                            c_ptr]),
    'kgdet_bbox_decode': (ctypes.c_int, [c_ptr, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_f32, c_i32, c_i32, c_i32, c_i32,
                                         c_i32, c_ptr, c_ptr, c_ptr]),
    'kgdet_bbox_finalize': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_f32, c_i32, c_i32, c_i32, c_i32,
                                           c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr]),
    'kgdet_pointwise_tiled_bytes': (c_sz, [c_i32, c_i32, ctypes.c_int]),
    'kgdet_pointwise_packed_weight_bytes': (c_sz, [c_i32, c_i32, ctypes.c_int]),
    'kgdet_pointwise_pack_weight': (ctypes.c_int, [c_ptr, c_ptr, c_i32, c_i32, ctypes.c_int, c_ptr]),
    'kgdet_nchw_to_tiled_bf16': (ctypes.c_int, [c_ptr, c_ptr, c_i32, c_i32, c_i32, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, c_ptr]),
    'kgdet_rows_to_tiled_bf16': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, ctypes.c_int64, c_i32, ctypes.c_int, ctypes.c_int,
                                                c_ptr]),
    'kgdet_groupnorm_relu_nhwc': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_f32, c_i32, ctypes.c_int, c_ptr, c_i32,
                                                 c_i32, c_i32, c_ptr]),
    'kgdet_pointwise_conv_tiled': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, ctypes.c_int,
                                                  c_ptr, c_i32, c_ptr]),
    'kgdet_dcn_backward_input_workspace_bytes': (c_sz, [_SHAPE_P, ctypes.c_int, ctypes.c_int]),
    'kgdet_dcn_backward_input': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                                c_ptr, _SHAPE_P, ctypes.c_int, ctypes.c_int,
                                                c_ptr, c_sz, c_ptr]),
    'kgdet_dcn_backward_weight_workspace_bytes': (c_sz, [_SHAPE_P, ctypes.c_int, ctypes.c_int]),
    'kgdet_dcn_backward_weight': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_f32,
                                                 _SHAPE_P, ctypes.c_int, ctypes.c_int, c_ptr, c_sz,
                                                 c_ptr]),
    'kgdet_conv_supported': (ctypes.c_int, [c_i32, c_i32, c_i32]),
    'kgdet_conv_split_planes_bytes': (c_sz, [c_i32, c_i32, c_i32, c_i32]),
    'kgdet_conv_split_planes_from_nchw': (ctypes.c_int, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    'kgdet_conv_split_planes_from_rows': (ctypes.c_int, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    'kgdet_conv_packed_weight_bytes': (c_sz, [c_i32, c_i32, c_i32]),
    'kgdet_conv_pack_weight': (ctypes.c_int, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_ptr]),
    'kgdet_conv_forward': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                                          ctypes.c_int, c_ptr]),
    'kgdet_conv_forward_pair': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32,
                                               c_i32, c_i32, c_i32, ctypes.c_int, c_ptr]),
    'kgdet_groupnorm_relu_nhwc_planes': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_f32, c_i32, ctypes.c_int, c_ptr, c_ptr,
                                                        c_i32, c_i32, c_i32, c_i32, c_ptr]),
    'kgdet_groupnorm_relu_nhwc_backward': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_f32, c_i32, ctypes.c_int, c_ptr, c_ptr,
                                                          c_ptr, c_i32, c_i32, c_i32, c_ptr]),
    'kgdet_groupnorm_stream_workspace_bytes': (ctypes.c_size_t, [c_i32, c_i32, c_i32, c_i32]),
    'kgdet_groupnorm_relu_nhwc_stream': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_f32, c_i32, ctypes.c_int, c_ptr, c_ptr,
                                                        ctypes.c_int, c_i32, c_i32, c_i32, c_i32, c_ptr, ctypes.c_size_t,
                                                        c_ptr]),
    'kgdet_point_assign_scratch_bytes': (ctypes.c_size_t, [c_i32, c_i32, c_i32]),
    'kgdet_point_assign': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_f32, c_i32, c_ptr,
                                          c_ptr, c_ptr, c_ptr, c_ptr]),
    'kgdet_point_losses_forward': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32,
                                                  c_i32, c_i32, c_i32, c_f32, c_f32, c_ptr, c_f32, c_f32, c_f32, c_ptr,
                                                  c_ptr]),
    'kgdet_point_losses_backward': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32,
                                                   c_i32, c_i32, c_i32, c_i32, c_f32, c_f32, c_ptr, c_f32, c_f32, c_f32,
                                                   c_ptr, c_ptr]),
    'kgdet_nms_workspace_bytes': (c_sz, [c_i32]),
    'kgdet_nms': (ctypes.c_int, [c_ptr, c_i32, c_f32, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_sz,
                                 c_ptr]),
    'kgdet_nms_batched_workspace_bytes': (c_sz, [c_i32, c_i32, c_i32]),
    'kgdet_nms_batched': (ctypes.c_int, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_f32, c_f32, ctypes.c_int,
                                         c_ptr, c_ptr, c_sz, c_ptr]),
    'kgdet_sigmoid_focal_loss_forward': (ctypes.c_int, [c_ptr, c_ptr, c_i32, c_i32, c_f32, c_f32,
                                                        c_ptr, ctypes.c_int, c_ptr]),
    'kgdet_sigmoid_focal_loss_backward': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_f32,
                                                         c_f32, c_ptr, ctypes.c_int, c_ptr]),
    'kgdet_sigmoid_focal_loss_sum_forward': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_i32, c_i32,
                                                            c_f32, c_f32, c_ptr, ctypes.c_int,
                                                            c_ptr]),
    'kgdet_sigmoid_focal_loss_sum_backward': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i32,
                                                             c_i32, c_f32, c_f32, c_ptr,
                                                             ctypes.c_int, c_ptr]),
    'kgdet_points2bbox_moment_forward': (ctypes.c_int, [c_ptr, c_ptr, c_i32, c_i32, c_i32,
                                                        ctypes.c_int, c_ptr, c_ptr]),
    'kgdet_points2bbox_moment_backward': (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32,
                                                         ctypes.c_int, c_f32, c_ptr, c_ptr, c_ptr]),
    'kgdet_topk_flagged': (ctypes.c_int, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    'kgdet_dcn_set_profile_events': (None, [c_ptr, c_ptr]),
    'kgdet_dcn_set_timeline': (None, [c_ptr, ctypes.c_longlong]),
    'kgdet_nchw_to_nhwc': (ctypes.c_int, [c_ptr, c_ptr, c_i32, c_i32, c_i32, ctypes.c_int,
                                          ctypes.c_int, c_ptr]),
}

_lib = None


def lib():
    """Load libkgdet_b200.so once.  Missing library is a hard error (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'kgdet_b200: %s not found. Build it with `python -m kgdet_b200.build` '
                '(nvcc, sm_100a). There is no CPU or PyTorch fallback.' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.kgdet_abi_version() != ABI_VERSION:
            raise RuntimeError('kgdet_b200: ABI version mismatch')
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().kgdet_last_error()
        raise RuntimeError('%s failed (%d): %s' % (what, rc, msg.decode() if msg else '?'))


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError('kgdet_b200 ops support float32 and bfloat16 tensors, got %s' % t.dtype)


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_of(t):
    """Current stream of the tensor's device.  The library launches on the process's CURRENT device (as the
    reference extension does: no device guard in deform_conv_cuda.cpp), so a tensor on another device is a loud
    error here instead of a launch with a foreign stream."""
    idx = t.device.index
    if idx is not None and idx != torch.cuda.current_device():
        raise RuntimeError('kgdet_b200: tensor lives on cuda:%d but the current device is cuda:%d; call '
                           'torch.cuda.set_device(%d) (or use `with torch.cuda.device(...)`) first'
                           % (idx, torch.cuda.current_device(), idx))
    return torch.cuda.current_stream(t.device).cuda_stream


def workspace(nbytes, like):
    """Scratch from torch's caching allocator (stream-ordered reuse, no cudaMalloc in steady state)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=like.device)


def require_cuda(t, name):
    if not t.is_cuda:
        # the reference raises NotImplementedError for CPU tensors (dcn/deform_conv.py:44-45)
        raise NotImplementedError('%s: kgdet_b200 ops are CUDA-only (got a %s tensor)' % (name, t.device))
