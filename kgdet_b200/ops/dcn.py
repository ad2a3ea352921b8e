"""Deformable convolution: host-side mirror of ``mmdet/ops/dcn/deform_conv.py``.

Same names, argument order, parameter set and error behaviour as the reference
(``DeformConvFunction`` DC.py:12-110, ``ModulatedDeformConvFunction`` :113-183,
``DeformConv`` :190-236, ``DeformConvPack`` :239-261, ``ModulatedDeformConv`` :264-308,
``ModulatedDeformConvPack`` :311-337), with the native work done by
``libkgdet_b200.so`` (include/kgdet_b200.h) instead of the ``deform_conv_cuda`` pybind
module.  Differences, all documented in DESIGN.md:

* ``backward`` returns one gradient slot per forward argument (9 / 10) -- the reference
  returns 8 for 9 inputs, which modern autograd rejects when ``im2col_step`` is passed.
* ``im2col_step`` is validated like the reference (DC.py:47-49) but not used: the fused
  kernels never materialise the column buffer it chunks.
* the arithmetic of the contraction is selectable (``set_precision``): ``'tf32x3'`` (default for
  fp32 tensors: fused tcgen05 kernel, hi/lo split operands + accumulator promotion, rel 1e-5 parity
  with the reference's fp32 SGEMM; shapes the tensor-core path does not cover and the whole backward
  run the exact FFMA kernels), ``'fp32'`` (exact FFMA everywhere), ``'bf16'`` (default for bf16
  tensors; fused tcgen05 kernel forward and backward, rel 1e-2) or ``'tf32'`` (single pass, 5e-3).
"""
import math
import os
import weakref

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from . import _capi

_precision_override = os.environ.get('KGDET_DCN_PRECISION') or None

# Grad mode of the CALLER of deform_conv / modulated_deform_conv: inside autograd.Function.forward grad mode is
# always off and ctx.needs_input_grad ignores torch.no_grad(), so the wrappers below record it for the
# derived-weight cache ("is autograd recording this call?").
import threading
_caller = threading.local()


def _recording(ctx):
    outer = getattr(_caller, 'grad_enabled', None)
    return any(ctx.needs_input_grad) and (outer is None or outer)


def set_precision(name):
    """'fp32' | 'tf32x3' | 'bf16' | 'tf32' | None (None = pick by tensor dtype)."""
    global _precision_override
    if name is not None and name not in _capi.PRECISIONS:
        raise ValueError('unknown precision %r' % (name,))
    _precision_override = name


def get_precision(dtype=torch.float32):
    if _precision_override is not None:
        return _precision_override
    return 'bf16' if dtype == torch.bfloat16 else 'tf32x3'


def _shape(input, weight, stride, padding, dilation, groups, deformable_groups):
    s = _capi.DcnShape()
    s.N, s.C, s.H, s.W = input.shape
    s.Cout, s.kh, s.kw = weight.shape[0], weight.shape[2], weight.shape[3]
    s.stride_h, s.stride_w = stride
    s.pad_h, s.pad_w = padding
    s.dil_h, s.dil_w = dilation
    s.groups, s.deformable_groups = groups, deformable_groups
    return s


def _output_size(input, weight, padding, dilation, stride):
    # DC.py:96-110
    channels = weight.size(0)
    output_size = (input.size(0), channels)
    for d in range(input.dim() - 2):
        in_size = input.size(d + 2)
        pad = padding[d]
        kernel = dilation[d] * (weight.size(d + 2) - 1) + 1
        stride_ = stride[d]
        output_size += ((in_size + (2 * pad) - kernel) // stride_ + 1, )
    if not all(map(lambda s: s > 0, output_size)):
        raise ValueError('convolution input is too small (output would be {})'.format(
            'x'.join(map(str, output_size))))
    return output_size


# ---- packed-weight cache -------------------------------------------------------------------
# The tensor-core path reads weights in a pre-swizzled UMMA layout; repacking is a few
# microseconds but is skipped while the parameter is unchanged (same object, same _version).
_pack_cache = {}


def _packed_weight(weight, shape, precision, training=False):
    """`training`: the call is recorded by autograd (inside an autograd.Function.forward grad mode is off, so the
    Function passes `any(ctx.needs_input_grad)` down)."""
    from .pointwise import _generation, cache_allowed
    lib = _capi.lib()
    key = (id(weight), precision, weight.device.index)
    use_cache = cache_allowed((weight,)) and not training
    sig = (_generation[0], weight._version, weight.data_ptr(), tuple(weight.shape))
    if use_cache:
        ent = _pack_cache.get(key)
        if ent is not None and ent[0]() is weight and ent[1] == sig:
            return ent[2]
    else:
        # training call: the pack is rebuilt every time (microseconds) and any cached copy is dropped -- the
        # weights are about to change, possibly by a replayed CUDA graph that never bumps `_version`
        _pack_cache.pop(key, None)
    w32 = weight.detach()
    if w32.dtype != torch.float32:
        w32 = w32.float()
    w32 = w32.contiguous()
    nbytes = lib.kgdet_dcn_packed_weight_bytes(ctypes_ref(shape), precision)
    packed = torch.empty(int(nbytes), dtype=torch.uint8, device=weight.device)
    _capi.check(lib.kgdet_dcn_pack_weight(w32.data_ptr(), packed.data_ptr(), ctypes_ref(shape),
                                          precision, _capi.stream_of(weight)),
                'kgdet_dcn_pack_weight')
    if not use_cache:
        return packed
    if len(_pack_cache) > 256:
        for k in [k for k, v in _pack_cache.items() if v[0]() is None]:
            del _pack_cache[k]
    try:
        _pack_cache[key] = (weakref.ref(weight), sig, packed)
    except TypeError:
        pass
    return packed


def ctypes_ref(shape):
    import ctypes
    return ctypes.byref(shape)


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def _check_offset(input, offset, weight, shape_out, deformable_groups, what='offset', per_tap=2):
    kh, kw = weight.shape[2], weight.shape[3]
    # DC.cpp:128-135 (shape_check) / :194 (batch)
    if offset.shape[0] != input.shape[0]:
        raise RuntimeError('invalid batch size of %s' % what)
    if offset.shape[2] != shape_out[2] or offset.shape[3] != shape_out[3]:
        raise RuntimeError('invalid spatial size of %s, expected height: %d width: %d, but got height: '
                           '%d width: %d' % (what, shape_out[2], shape_out[3], offset.shape[2],
                                             offset.shape[3]))
    if offset.shape[1] != deformable_groups * per_tap * kh * kw:
        raise RuntimeError('invalid number of channels of %s' % what)


def _dcn_forward(input, offset, mask, weight, bias, stride, padding, dilation, groups,
                 deformable_groups, training=False):
    lib = _capi.lib()
    _capi.require_cuda(input, 'deform_conv')
    if weight.dim() != 4:
        raise RuntimeError('4D weight tensor (nOutputPlane,nInputPlane,kH,kW) expected, but got: %d'
                           % weight.dim())
    if input.shape[1] != weight.shape[1] * groups:
        raise RuntimeError('invalid number of input planes, expected: %d, but got: %d'
                           % (weight.shape[1] * groups, input.shape[1]))
    out_size = _output_size(input, weight, padding, dilation, stride)
    _check_offset(input, offset, weight, out_size, deformable_groups)
    if mask is not None:
        _check_offset(input, mask, weight, out_size, deformable_groups, 'mask', 1)
    x = input.detach().contiguous()
    dt = _capi.dtype_code(x)
    prec = _capi.PRECISIONS[get_precision(x.dtype)]
    shape = _shape(x, weight, stride, padding, dilation, groups, deformable_groups)
    packed = _packed_weight(weight, shape, prec, training)
    off = _f32c(offset)
    msk = None if mask is None else _f32c(mask)
    b = None if bias is None else _f32c(bias)
    output = x.new_empty(out_size)
    nws = lib.kgdet_dcn_forward_workspace_bytes(ctypes_ref(shape), dt, prec)
    ws = _capi.workspace(nws, x)
    _capi.check(lib.kgdet_dcn_forward(x.data_ptr(), off.data_ptr(), _capi.ptr(msk), packed.data_ptr(),
                                      _capi.ptr(b), output.data_ptr(), ctypes_ref(shape), dt, prec,
                                      ws.data_ptr(), ws.numel(), _capi.stream_of(x)),
                'kgdet_dcn_forward')
    return output


def _dcn_backward_input(input, offset, mask, weight, grad_output, stride, padding, dilation, groups,
                        deformable_groups):
    lib = _capi.lib()
    x = input.detach().contiguous()
    go = grad_output.detach().to(x.dtype).contiguous()
    dt = _capi.dtype_code(x)
    prec = _capi.PRECISIONS[get_precision(x.dtype)]
    shape = _shape(x, weight, stride, padding, dilation, groups, deformable_groups)
    off = _f32c(offset)
    msk = None if mask is None else _f32c(mask)
    w32 = _f32c(weight)
    grad_input = torch.empty_like(x)
    grad_offset = torch.empty_like(off)
    grad_mask = None if msk is None else torch.empty_like(msk)
    nws = lib.kgdet_dcn_backward_input_workspace_bytes(ctypes_ref(shape), dt, prec)
    ws = _capi.workspace(nws, x)
    _capi.check(lib.kgdet_dcn_backward_input(
        x.data_ptr(), off.data_ptr(), _capi.ptr(msk), w32.data_ptr(), go.data_ptr(),
        grad_input.data_ptr(), grad_offset.data_ptr(), _capi.ptr(grad_mask), ctypes_ref(shape), dt, prec,
        ws.data_ptr(), ws.numel(), _capi.stream_of(x)), 'kgdet_dcn_backward_input')
    grad_offset = grad_offset.to(offset.dtype)
    if grad_mask is not None:
        grad_mask = grad_mask.to(mask.dtype)
    return grad_input, grad_offset, grad_mask


def _dcn_backward_weight(input, offset, mask, weight, grad_output, with_bias, stride, padding,
                         dilation, groups, deformable_groups):
    lib = _capi.lib()
    x = input.detach().contiguous()
    go = grad_output.detach().to(x.dtype).contiguous()
    dt = _capi.dtype_code(x)
    prec = _capi.PRECISIONS[get_precision(x.dtype)]
    shape = _shape(x, weight, stride, padding, dilation, groups, deformable_groups)
    off = _f32c(offset)
    msk = None if mask is None else _f32c(mask)
    grad_weight = torch.empty(weight.shape, dtype=torch.float32, device=x.device)
    grad_bias = torch.empty(weight.shape[0], dtype=torch.float32, device=x.device) if with_bias else None
    nws = lib.kgdet_dcn_backward_weight_workspace_bytes(ctypes_ref(shape), dt, prec)
    ws = _capi.workspace(nws, x)
    _capi.check(lib.kgdet_dcn_backward_weight(
        x.data_ptr(), off.data_ptr(), _capi.ptr(msk), go.data_ptr(), grad_weight.data_ptr(),
        _capi.ptr(grad_bias), 1.0, ctypes_ref(shape), dt, prec, ws.data_ptr(), ws.numel(),
        _capi.stream_of(x)), 'kgdet_dcn_backward_weight')
    grad_weight = grad_weight.to(weight.dtype)
    return grad_weight, grad_bias


# ---- prepared (shared) forward: inference fast path used by kgdet_b200.head ------------------------
class PreparedInput(object):
    """NHWC copy (+ guard band) of an activation, shareable by every deformable convolution that reads
    it (kgdet_dcn_prepare_input)."""
    __slots__ = ('buf', 'shape4', 'dtype_code', 'precision', 'fast', 'torch_dtype')


class SamplePlan(object):
    """Per-(position, tap) sampling records of one offset tensor (kgdet_dcn_prepare_plan), shareable by
    the convolutions that use the same offsets (the cls and keypoint branches of a Kp3RepBlock)."""
    __slots__ = ('buf', 'geom', 'precision', 'fast')


def _geom_shape(n, c, h, w, cout, k, stride, padding, dilation):
    s = _capi.DcnShape()
    s.N, s.C, s.H, s.W = n, c, h, w
    s.Cout, s.kh, s.kw = cout, k[0], k[1]
    s.stride_h, s.stride_w = stride
    s.pad_h, s.pad_w = padding
    s.dil_h, s.dil_w = dilation
    s.groups, s.deformable_groups = 1, 1
    return s


def prepare_input(x, out_channels, kernel_size=3, stride=1, padding=1, dilation=1, precision=None):
    lib = _capi.lib()
    _capi.require_cuda(x, 'prepare_input')
    x = x.detach()
    # a channels_last fp32 activation is already position-major rows: no transpose kernel
    rows_src = (x.dim() == 4 and x.dtype == torch.float32 and not x.is_contiguous()
                and x.is_contiguous(memory_format=torch.channels_last))
    if not rows_src:
        x = x.contiguous()
    prec_name = precision or get_precision(x.dtype)
    prec = _capi.PRECISIONS[prec_name]
    shape = _geom_shape(*x.shape, out_channels, _pair(kernel_size), _pair(stride), _pair(padding), _pair(dilation))
    p = PreparedInput()
    p.shape4 = tuple(x.shape)
    p.dtype_code = _capi.dtype_code(x)
    p.torch_dtype = x.dtype
    p.precision = prec
    p.fast = bool(lib.kgdet_dcn_fast_path_supported(ctypes_ref(shape), prec))
    p.buf = torch.empty(int(lib.kgdet_dcn_prepared_input_bytes(ctypes_ref(shape), prec)), dtype=torch.uint8,
                        device=x.device)
    if rows_src:
        _capi.check(lib.kgdet_dcn_prepare_input_rows(x.data_ptr(), p.buf.data_ptr(), ctypes_ref(shape), prec,
                                                     _capi.stream_of(x)), 'kgdet_dcn_prepare_input_rows')
    else:
        _capi.check(lib.kgdet_dcn_prepare_input(x.data_ptr(), p.buf.data_ptr(), ctypes_ref(shape), p.dtype_code, prec,
                                                _capi.stream_of(x)), 'kgdet_dcn_prepare_input')
    return p


def prepare_plan(offset, input_shape, out_channels, kernel_size, stride=1, padding=0, dilation=1, mask=None,
                 precision=None, like_dtype=torch.float32):
    lib = _capi.lib()
    _capi.require_cuda(offset, 'prepare_plan')
    prec = _capi.PRECISIONS[precision or get_precision(like_dtype)]
    k, st, pd, dl = _pair(kernel_size), _pair(stride), _pair(padding), _pair(dilation)
    shape = _geom_shape(*input_shape, out_channels, k, st, pd, dl)
    off = _f32c(offset)
    msk = None if mask is None else _f32c(mask)
    pl = SamplePlan()
    pl.geom = (tuple(input_shape), k, st, pd, dl)
    pl.precision = prec
    pl.fast = bool(lib.kgdet_dcn_fast_path_supported(ctypes_ref(shape), prec))
    pl.buf = torch.empty(int(lib.kgdet_dcn_plan_bytes(ctypes_ref(shape), prec)), dtype=torch.uint8,
                         device=offset.device)
    _capi.check(lib.kgdet_dcn_prepare_plan(off.data_ptr(), _capi.ptr(msk), pl.buf.data_ptr(), ctypes_ref(shape),
                                           prec, _capi.stream_of(off)), 'kgdet_dcn_prepare_plan')
    return pl


def prepare_plan_points(points, channel_offset, input_shape, out_channels, kernel_size, stride=1, padding=0,
                        dilation=1, precision=None, like_dtype=torch.float32, gradient_mul=0.0):
    """Sample plan of ``points[:, channel_offset:channel_offset + 2K] - base_grid`` without materialising the
    slice or the subtraction (KP3:131-143: the head's ``dcn_offset = pts - dcn_base_offset``).  With
    ``gradient_mul`` the head's ``gradient_mul * pts + (1 - gradient_mul) * pts.detach()`` (KP3:135-143) is
    evaluated first in fp32, as the reference does even in inference (bit-identical sample locations)."""
    lib = _capi.lib()
    _capi.require_cuda(points, 'prepare_plan_points')
    assert points.dtype == torch.float32 and points.is_contiguous() and points.dim() == 4
    prec = _capi.PRECISIONS[precision or get_precision(like_dtype)]
    k, st, pd, dl = _pair(kernel_size), _pair(stride), _pair(padding), _pair(dilation)
    shape = _geom_shape(*input_shape, out_channels, k, st, pd, dl)
    pl = SamplePlan()
    pl.geom = (tuple(input_shape), k, st, pd, dl)
    pl.precision = prec
    pl.fast = bool(lib.kgdet_dcn_fast_path_supported(ctypes_ref(shape), prec))
    pl.buf = torch.empty(int(lib.kgdet_dcn_plan_bytes(ctypes_ref(shape), prec)), dtype=torch.uint8,
                         device=points.device)
    _capi.check(lib.kgdet_dcn_prepare_plan_points(points.data_ptr(), int(channel_offset), points.shape[1],
                                                  float(gradient_mul), float(1 - gradient_mul) if gradient_mul else 0.0,
                                                  pl.buf.data_ptr(), ctypes_ref(shape), prec,
                                                  _capi.stream_of(points)), 'kgdet_dcn_prepare_plan_points')
    return pl


def deform_conv_prepared(pin, plan, weight, out=None, channel_offset=0, relu=False, bias=None):
    """out[:, channel_offset:channel_offset+Cout] = (relu)(deform_conv(input, offset, weight)); no autograd.

    `out` is an NCHW tensor, or a ``pointwise.TiledRows`` (tensor-core path only): the kernel then writes
    position-major bf16 rows straight into the tiled A operand of the 1x1-convolution GEMM that follows."""
    from .pointwise import TiledRows
    lib = _capi.lib()
    (n, c, h, w), k, st, pd, dl = plan.geom
    assert pin.shape4 == (n, c, h, w) and pin.precision == plan.precision
    assert tuple(weight.shape[1:]) == (c, k[0], k[1]), 'weight does not match the prepared geometry'
    shape = _geom_shape(n, c, h, w, weight.shape[0], k, st, pd, dl)
    fast = bool(lib.kgdet_dcn_fast_path_supported(ctypes_ref(shape), plan.precision))
    assert fast == pin.fast == plan.fast, 'prepared buffers were built for a different kernel path'
    packed = _packed_weight(weight, shape, plan.precision)
    ho = (h + 2 * pd[0] - (dl[0] * (k[0] - 1) + 1)) // st[0] + 1
    wo = (w + 2 * pd[1] - (dl[1] * (k[1] - 1) + 1)) // st[1] + 1
    b = None if bias is None else _f32c(bias)
    if isinstance(out, TiledRows):
        assert out.M == n * ho * wo
        layout = _capi.LAYOUT_TILED_SPLIT if out.split else _capi.LAYOUT_TILED
        ptr, ctot, dt, ref = out.buf.data_ptr(), out.K, _capi.BF16, out.buf
    else:
        if out is None:
            out = torch.empty((n, weight.shape[0], ho, wo), dtype=pin.torch_dtype, device=weight.device)
            channel_offset = 0
        assert out.is_contiguous() and out.shape[0] == n and tuple(out.shape[2:]) == (ho, wo)
        layout = _capi.LAYOUT_NCHW
        ptr, ctot, dt, ref = out.data_ptr(), out.shape[1], _capi.dtype_code(out), out
    ws, nws = None, 0
    if layout == _capi.LAYOUT_NCHW:
        nws = int(lib.kgdet_dcn_forward_prepared_workspace_bytes(ctypes_ref(shape), plan.precision))
        if nws:
            ws = _capi.workspace(nws, ref)
    _capi.check(lib.kgdet_dcn_forward_prepared(pin.buf.data_ptr(), plan.buf.data_ptr(), packed.data_ptr(),
                                               _capi.ptr(b), ptr, int(channel_offset), ctot, int(bool(relu)), layout,
                                               ctypes_ref(shape), dt, plan.precision, _capi.ptr(ws), nws,
                                               _capi.stream_of(ref)),
                'kgdet_dcn_forward_prepared')
    return out




def deform_conv_prepared_group(jobs):
    """ONE persistent launch for several prepared deformable convolutions (kgdet_dcn_forward_prepared_group): the
    six DCNs of a Kp3RepBlock stage.  jobs: list of (pin, plan, weight, out, channel_offset, relu) as for
    `deform_conv_prepared`; bf16 mode, equal Cout.  Falls back to single launches when the grouped kernel does not
    cover a job.  Results are bit-identical either way."""
    import ctypes
    from .pointwise import TiledRows
    lib = _capi.lib()
    items = (_capi.DcnGroupItem * len(jobs))()
    keep = []
    ok = 1 <= len(jobs) <= _capi.DCN_GROUP_MAX
    ref = None
    for i, (pin, plan, weight, out, channel_offset, relu) in enumerate(jobs):
        (n, c, h, w), k, st, pd, dl = plan.geom
        assert pin.shape4 == (n, c, h, w) and pin.precision == plan.precision
        shape = _geom_shape(n, c, h, w, weight.shape[0], k, st, pd, dl)
        ok = ok and plan.precision == _capi.PREC_BF16 and bool(lib.kgdet_dcn_group_supported(ctypes_ref(shape), plan.precision))
        ok = ok and pin.fast and plan.fast and weight.shape[0] == jobs[0][2].shape[0]
        if not ok:
            break
        packed = _packed_weight(weight, shape, plan.precision)
        keep.append(packed)
        it = items[i]
        it.prepared_input, it.plan, it.weight_packed, it.bias = pin.buf.data_ptr(), plan.buf.data_ptr(), packed.data_ptr(), None
        if isinstance(out, TiledRows):
            it.out_layout = _capi.LAYOUT_TILED_SPLIT if out.split else _capi.LAYOUT_TILED
            it.output, it.out_channels_total, it.dtype = out.buf.data_ptr(), out.K, _capi.BF16
            ok = ok and channel_offset % 64 == 0
            ref = out.buf
        else:
            assert out.is_contiguous() and out.shape[0] == n
            it.out_layout = _capi.LAYOUT_NCHW
            it.output, it.out_channels_total, it.dtype = out.data_ptr(), out.shape[1], _capi.dtype_code(out)
            ref = out
        it.out_channel_offset, it.fuse_relu = int(channel_offset), int(bool(relu))
        it.shape = shape
    if not ok:
        for job in jobs:
            deform_conv_prepared(*job)
        return
    # Tile counter of the persistent launch (zeroed by the library on the launch stream).  A fresh stream-ordered
    # allocation per call: a counter cached per (device, stream) was shared by every CUDA graph captured on torch's
    # capture stream, and two such graphs replayed at the same time on two streams then handed out each other's
    # tiles (tools/lockstep_stress.py: 20 of 40 simultaneous replays of two instances returned wrong detections).
    ctr = _capi.workspace(16, ref)
    _capi.check(lib.kgdet_dcn_forward_prepared_group(ctypes.cast(items, ctypes.c_void_p), len(jobs), _capi.PREC_BF16,
                                                     ctr.data_ptr(), 16, _capi.stream_of(ref)),
                'kgdet_dcn_forward_prepared_group')


class DeformConvFunction(Function):
    """Mirror of DC.py:12-110."""

    @staticmethod
    def forward(ctx, input, offset, weight, stride=1, padding=0, dilation=1, groups=1,
                deformable_groups=1, im2col_step=64):
        if input is not None and input.dim() != 4:
            raise ValueError('Expected 4D tensor as input, got {}D tensor instead.'.format(input.dim()))
        ctx.stride = _pair(stride)
        ctx.padding = _pair(padding)
        ctx.dilation = _pair(dilation)
        ctx.groups = groups
        ctx.deformable_groups = deformable_groups
        ctx.im2col_step = im2col_step
        ctx.save_for_backward(input, offset, weight)
        if not input.is_cuda:
            raise NotImplementedError
        cur_im2col_step = min(ctx.im2col_step, input.shape[0])
        assert (input.shape[0] % cur_im2col_step) == 0, 'im2col step must divide batchsize'
        return _dcn_forward(input, offset, None, weight, None, ctx.stride, ctx.padding, ctx.dilation,
                            ctx.groups, ctx.deformable_groups, training=_recording(ctx))

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input, offset, weight = ctx.saved_tensors
        grad_input = grad_offset = grad_weight = None
        if not grad_output.is_cuda:
            raise NotImplementedError
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:        # DC.py:72
            grad_input, grad_offset, _ = _dcn_backward_input(
                input, offset, None, weight, grad_output, ctx.stride, ctx.padding, ctx.dilation,
                ctx.groups, ctx.deformable_groups)
        if ctx.needs_input_grad[2]:                                   # DC.py:83
            grad_weight, _ = _dcn_backward_weight(
                input, offset, None, weight, grad_output, False, ctx.stride, ctx.padding,
                ctx.dilation, ctx.groups, ctx.deformable_groups)
        return (grad_input, grad_offset, grad_weight, None, None, None, None, None, None)

    _output_size = staticmethod(_output_size)


class ModulatedDeformConvFunction(Function):
    """Mirror of DC.py:113-183."""

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1,
                groups=1, deformable_groups=1):
        ctx.stride = stride
        ctx.padding = padding
        ctx.dilation = dilation
        ctx.groups = groups
        ctx.deformable_groups = deformable_groups
        ctx.with_bias = bias is not None
        if not input.is_cuda:
            raise NotImplementedError
        ctx.save_for_backward(input, offset, mask, weight)
        return _dcn_forward(input, offset, mask, weight, bias, _pair(stride), _pair(padding),
                            _pair(dilation), groups, deformable_groups, training=_recording(ctx))

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        input, offset, mask, weight = ctx.saved_tensors
        s, p, d = _pair(ctx.stride), _pair(ctx.padding), _pair(ctx.dilation)
        grad_input, grad_offset, grad_mask = _dcn_backward_input(
            input, offset, mask, weight, grad_output, s, p, d, ctx.groups, ctx.deformable_groups)
        grad_weight, grad_bias = _dcn_backward_weight(
            input, offset, mask, weight, grad_output, ctx.with_bias, s, p, d, ctx.groups,
            ctx.deformable_groups)
        return (grad_input, grad_offset, grad_mask, grad_weight, grad_bias, None, None, None, None, None)

    @staticmethod
    def _infer_shape(ctx, input, weight):
        n = input.size(0)
        channels_out = weight.size(0)
        height, width = input.shape[2:4]
        kernel_h, kernel_w = weight.shape[2:4]
        height_out = (height + 2 * ctx.padding - (ctx.dilation * (kernel_h - 1) + 1)) // ctx.stride + 1
        width_out = (width + 2 * ctx.padding - (ctx.dilation * (kernel_w - 1) + 1)) // ctx.stride + 1
        return n, channels_out, height_out, width_out


def deform_conv(*args):
    """DeformConvFunction.apply (DC.py:186) -- records the caller's grad mode for the weight cache."""
    _caller.grad_enabled = torch.is_grad_enabled()
    try:
        return DeformConvFunction.apply(*args)
    finally:
        _caller.grad_enabled = None


def modulated_deform_conv(*args):
    """ModulatedDeformConvFunction.apply (DC.py:187)."""
    _caller.grad_enabled = torch.is_grad_enabled()
    try:
        return ModulatedDeformConvFunction.apply(*args)
    finally:
        _caller.grad_enabled = None


class DeformConv(nn.Module):
    """Mirror of DC.py:190-236: one parameter ``weight``, no ``bias`` attribute."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, deformable_groups=1, bias=False):
        super(DeformConv, self).__init__()
        assert not bias
        assert in_channels % groups == 0, \
            'in_channels {} cannot be divisible by groups {}'.format(in_channels, groups)
        assert out_channels % groups == 0, \
            'out_channels {} cannot be divisible by groups {}'.format(out_channels, groups)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.weight = nn.Parameter(
            torch.Tensor(out_channels, in_channels // self.groups, *self.kernel_size))
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)

    def forward(self, x, offset):
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation,
                           self.groups, self.deformable_groups)


class DeformConvPack(DeformConv):
    """Mirror of DC.py:239-261."""

    def __init__(self, *args, **kwargs):
        super(DeformConvPack, self).__init__(*args, **kwargs)
        self.conv_offset = nn.Conv2d(
            self.in_channels,
            self.deformable_groups * 2 * self.kernel_size[0] * self.kernel_size[1],
            kernel_size=self.kernel_size, stride=_pair(self.stride), padding=_pair(self.padding),
            bias=True)
        self.init_offset()

    def init_offset(self):
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()

    def forward(self, x):
        offset = self.conv_offset(x)
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation,
                           self.groups, self.deformable_groups)


class ModulatedDeformConv(nn.Module):
    """Mirror of DC.py:264-308."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, deformable_groups=1, bias=True):
        super(ModulatedDeformConv, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.with_bias = bias
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, offset, mask):
        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride,
                                     self.padding, self.dilation, self.groups, self.deformable_groups)


class ModulatedDeformConvPack(ModulatedDeformConv):
    """Mirror of DC.py:311-337."""

    def __init__(self, *args, **kwargs):
        super(ModulatedDeformConvPack, self).__init__(*args, **kwargs)
        self.conv_offset_mask = nn.Conv2d(
            self.in_channels,
            self.deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1],
            kernel_size=self.kernel_size, stride=_pair(self.stride), padding=_pair(self.padding),
            bias=True)
        self.init_offset()

    def init_offset(self):
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()

    def forward(self, x):
        out = self.conv_offset_mask(x)
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        offset = torch.cat((o1, o2), dim=1)
        mask = torch.sigmoid(mask)
        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride,
                                     self.padding, self.dilation, self.groups, self.deformable_groups)
