"""Pointwise (1x1) convolutions of the Kp3RepBlock on the tensor cores.

``cls_out`` / ``keypts_out`` / ``reppts_out`` of the reference block
(reppoints_head_kp3rep_cas_1_assign_once.py:79-96,152-171) are 1x1 ``nn.Conv2d`` on the concatenated,
ReLU-ed deformable-convolution outputs.  Here the fused DCN kernel writes those activations as position-major
bf16 rows in the tiled layout the tensor core reads (``TiledRows``) and ONE tcgen05 GEMM
(kgdet_pointwise_conv_tiled) per branch produces its outputs; bias, the cascade's residual adds
(KP3:431-432,440-441) and the NCHW fp32 layout of the results are the GEMM's epilogue.  With ``split=True``
both operands carry a bf16 hi and a bf16 lo part and the GEMM issues three MMAs per k-step ("bf16x3"): the
1x1 convolutions stay fp32-grade although they run on the bf16 tensor cores.
"""
import ctypes
import weakref

import torch

from . import _capi

_cache = {}


class TiledRows(object):
    """Position-major activations [M, K] in the UMMA-tiled bf16 layout (opaque buffer)."""
    __slots__ = ('buf', 'M', 'K', 'split')

    def __init__(self, M, K, split, device):
        lib = _capi.lib()
        assert K % 64 == 0, 'K must be a multiple of 64'
        nbytes = int(lib.kgdet_pointwise_tiled_bytes(M, K, int(bool(split))))
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.M, self.K, self.split = M, K, bool(split)

    def to_dense(self):
        """fp32 [M, K] view of the stored values (hi + lo when split) -- for tests."""
        kb = (2 if self.split else 1) * (self.K // 64)
        t = self.buf.view(torch.bfloat16).view(-1, kb, 128, 8, 8)                   # tile, k-block, row, chunk, elem
        r = torch.arange(128, device=t.device).view(128, 1)
        c = torch.arange(8, device=t.device).view(1, 8)
        src = (c ^ (r & 7)).view(1, 1, 128, 8, 1).expand(t.shape[0], kb, 128, 8, 8)  # logical chunk c lives at c^(r&7)
        logical = torch.gather(t, 3, src).float()
        rows = logical.permute(0, 2, 1, 3, 4).reshape(t.shape[0] * 128, kb * 64)[:self.M]
        return rows[:, :self.K] + rows[:, self.K:] if self.split else rows


_generation = [0]


def invalidate_weight_caches():
    """Drop every derived-weight buffer (packed DCN / 1x1 weights, channels_last copies).

    The caches are keyed on (parameter identity, ``_version``, ``data_ptr``).  Two update paths change a
    parameter WITHOUT bumping ``_version``: in-place writes through ``.data`` (``p.data.copy_``, legacy
    optimisers, mmcv's Fp16OptimizerHook / EMAHook) and replays of a CUDA graph that contains the optimiser
    step.  Call this after such an update and before the next inference call.  Training-mode calls (grad enabled
    on a parameter that requires grad) never use the caches, and a ``GraphedInference`` is a snapshot of the
    weights at capture time by contract (``GraphedInference.refresh_weights()`` re-captures)."""
    _generation[0] += 1
    _cache.clear()
    from . import dcn
    dcn._pack_cache.clear()


def cache_allowed(params):
    """Derived-weight caches serve inference only: while autograd is recording for a parameter the weights are
    about to change (optimizer step, possibly inside a replayed CUDA graph that Python never sees)."""
    return not (torch.is_grad_enabled() and any(p.requires_grad for p in params))


def cached(params, build, tag=None):
    """Cache of host-prepared device buffers keyed by `tag` (what is derived) + the identity + version of the
    source parameters."""
    key = (tag,) + tuple(id(p) for p in params)
    if not cache_allowed(params):
        _cache.pop(key, None)               # a training call: whatever was derived from these weights is stale soon
        return build()
    sig = (_generation[0],) + tuple((p._version, p.data_ptr()) for p in params)
    ent = _cache.get(key)
    if ent is not None and ent[0] == sig and all(r() is p for r, p in zip(ent[1], params)):
        return ent[2]
    val = build()
    if len(_cache) > 256:
        # prune entries whose parameters are gone; live entries are never dropped (a captured CUDA graph may
        # hold raw pointers into them -- GraphedInference additionally keeps its own references)
        for k in [k for k, v in _cache.items() if any(r() is None for r in v[1])]:
            del _cache[k]
    _cache[key] = (sig, [weakref.ref(p) for p in params], val)
    return val


def cache_values():
    """Every buffer currently held by the derived-weight caches (GraphedInference pins them)."""
    from . import dcn
    return [v[2] for v in _cache.values()] + [v[2] for v in dcn._pack_cache.values()]


def pack_weight(w, split=True):
    """fp32 [Nout, K] (or [Nout, K, 1, 1]) -> packed GEMM operand (kgdet_pointwise_pack_weight)."""
    lib = _capi.lib()
    w = w.detach().float().flatten(1).contiguous()
    nout, k = w.shape
    nbytes = int(lib.kgdet_pointwise_packed_weight_bytes(nout, k, int(bool(split))))
    if nbytes == 0:
        raise ValueError('pointwise weights need K % 64 == 0, got %r' % (tuple(w.shape),))
    packed = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    _capi.check(lib.kgdet_pointwise_pack_weight(w.data_ptr(), packed.data_ptr(), nout, k, int(bool(split)),
                                                _capi.stream_of(w)), 'kgdet_pointwise_pack_weight')
    return packed, nout, k, bool(split)


def nchw_to_tiled(x, relu=False, split=True, bias=None):
    """[N, C, H, W] fp32/bf16 -> TiledRows [N*H*W, C] (optional per-channel bias, then optional ReLU)."""
    lib = _capi.lib()
    _capi.require_cuda(x, 'nchw_to_tiled')
    x = x.detach()
    n, c, h, w = x.shape
    rows = TiledRows(n * h * w, c, split, x.device)
    if x.dtype == torch.float32 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last):
        # channels_last source: already position-major rows, no transpose; the bias add rides along
        b = None if bias is None else bias.detach().float().contiguous()
        _capi.check(lib.kgdet_rows_to_tiled_bf16(x.data_ptr(), _capi.ptr(b), rows.buf.data_ptr(), n * h * w, c,
                                                 int(bool(relu)), int(bool(split)), _capi.stream_of(x)),
                    'kgdet_rows_to_tiled_bf16')
        return rows
    if bias is not None:
        x = x + bias.detach().to(x.dtype).view(1, -1, 1, 1)
    x = x.contiguous()
    _capi.check(lib.kgdet_nchw_to_tiled_bf16(x.data_ptr(), rows.buf.data_ptr(), n, c, h * w, _capi.dtype_code(x),
                                             int(bool(relu)), int(bool(split)), _capi.stream_of(x)),
                'kgdet_nchw_to_tiled_bf16')
    return rows


def to_channels_last(x):
    """NCHW fp32 -> the same values as a channels_last tensor, through the library's tiled transpose
    (kgdet_nchw_to_nhwc; about half the time of Tensor.contiguous(memory_format=channels_last) at [16,256,25,42])."""
    lib = _capi.lib()
    _capi.require_cuda(x, 'to_channels_last')
    if x.dim() != 4 or x.dtype != torch.float32 or not x.is_contiguous():
        return x.contiguous(memory_format=torch.channels_last)
    n, c, h, w = x.shape
    y = torch.empty_like(x, memory_format=torch.channels_last)
    code = _capi.dtype_code(x)
    _capi.check(lib.kgdet_nchw_to_nhwc(x.detach().data_ptr(), y.data_ptr(), n, c, h * w, code, code, _capi.stream_of(x)),
                'kgdet_nchw_to_nhwc')
    return y


def groupnorm_relu_nhwc(x, gn, relu=True, dense=True, prepared_for=None):
    """GroupNorm (+ ReLU) of a channels_last fp32 activation; returns a channels_last tensor.
    `gn` is the torch.nn.GroupNorm module (same parameters and semantics).  Maps of up to 1600 positions are
    normalised by one resident kernel; larger ones (FPN levels P3 / P4) by the two streaming passes.

    prepared_for: out_channels of a following 3x3 deformable convolution -- the streaming kernel then also writes
    the fused bf16 DCN's PreparedInput (returned second; with dense=False only that is produced and the first result
    is None)."""
    lib = _capi.lib()
    _capi.require_cuda(x, 'groupnorm_relu_nhwc')
    assert x.dim() == 4 and x.dtype == torch.float32 and x.is_contiguous(memory_format=torch.channels_last)
    n, c, h, w = x.shape
    gamma, beta = gn.weight.detach().float().contiguous(), gn.bias.detach().float().contiguous()
    resident = h * w <= 1600
    if resident and prepared_for is None:
        y = torch.empty_like(x, memory_format=torch.channels_last)
        _capi.check(lib.kgdet_groupnorm_relu_nhwc(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(gn.eps),
                                                  int(gn.num_groups), int(bool(relu)), y.data_ptr(), n, h * w, c,
                                                  _capi.stream_of(x)), 'kgdet_groupnorm_relu_nhwc')
        return y
    ws_bytes = int(lib.kgdet_groupnorm_stream_workspace_bytes(n, h * w, c, int(gn.num_groups)))
    if ws_bytes == 0:
        raise ValueError('groupnorm_relu_nhwc: unsupported shape [%d, %d, %d, %d] with %d groups for the streaming '
                         'kernel' % (n, c, h, w, gn.num_groups))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    y = torch.empty_like(x, memory_format=torch.channels_last) if dense else None
    prep = None
    if prepared_for is not None:
        from .dcn import PreparedInput, _geom_shape, ctypes_ref
        shape = _geom_shape(n, c, h, w, int(prepared_for), (3, 3), (1, 1), (1, 1), (1, 1))
        prep = PreparedInput()
        prep.shape4 = (n, c, h, w)
        prep.dtype_code = _capi.F32
        prep.torch_dtype = torch.float32
        prep.precision = _capi.PREC_BF16
        prep.fast = bool(lib.kgdet_dcn_fast_path_supported(ctypes_ref(shape), _capi.PREC_BF16))
        if not prep.fast:
            raise ValueError('groupnorm_relu_nhwc: the fused deformable convolution does not support this shape')
        prep.buf = torch.empty(int(lib.kgdet_dcn_prepared_input_bytes(ctypes_ref(shape), _capi.PREC_BF16)),
                               dtype=torch.uint8, device=x.device)
        assert prep.buf.numel() * 2 == int(lib.kgdet_conv_split_planes_bytes(n, c, h, w)), 'plane layouts differ'
    else:
        assert dense
    _capi.check(lib.kgdet_groupnorm_relu_nhwc_stream(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(gn.eps),
                                                     int(gn.num_groups), int(bool(relu)), _capi.ptr(y),
                                                     None if prep is None else prep.buf.data_ptr(), 1, n, h, w, c,
                                                     ws.data_ptr(), ws_bytes, _capi.stream_of(x)),
                'kgdet_groupnorm_relu_nhwc_stream')
    return y if prepared_for is None else (y, prep)


class _GroupNormReLUNHWC(torch.autograd.Function):
    """GroupNorm (+ ReLU) on a channels_last fp32 activation with both directions on this library's kernels
    (kgdet_groupnorm_relu_nhwc / _backward): the training step of the towers (ConvModule's norm + activation,
    mmdet/models/utils/conv_module.py:96-110,156-164) without ATen's moments / normalise / clamp kernels forward
    and its four kernels backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, num_groups, eps, relu):
        lib = _capi.lib()
        n, c, h, w = x.shape
        xd = x.detach()
        gamma, beta = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        y = torch.empty_like(xd, memory_format=torch.channels_last)
        _capi.check(lib.kgdet_groupnorm_relu_nhwc(xd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(eps),
                                                  int(num_groups), int(bool(relu)), y.data_ptr(), n, h * w, c,
                                                  _capi.stream_of(xd)), 'kgdet_groupnorm_relu_nhwc')
        ctx.save_for_backward(xd, gamma, beta)
        ctx.cfg = (int(num_groups), float(eps), bool(relu))
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        lib = _capi.lib()
        x, gamma, beta = ctx.saved_tensors
        groups, eps, relu = ctx.cfg
        n, c, h, w = x.shape
        dy = dy.detach().float().contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        parts = torch.empty((2, n, c), dtype=torch.float32, device=x.device)
        _capi.check(lib.kgdet_groupnorm_relu_nhwc_backward(x.data_ptr(), dy.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                                           eps, groups, int(relu), dx.data_ptr(), parts[0].data_ptr(),
                                                           parts[1].data_ptr(), n, h * w, c, _capi.stream_of(x)),
                    'kgdet_groupnorm_relu_nhwc_backward')
        sums = parts.sum(1)
        return dx, sums[0], sums[1], None, None, None


def groupnorm_relu_nhwc_autograd(x, gn, relu=True):
    """Differentiable GroupNorm (+ ReLU) of a channels_last fp32 activation (maps of at most 1600 positions, whole
    groups inside 32-channel blocks); anything else goes through torch."""
    n, c, h, w = x.shape
    cpg = c // gn.num_groups
    ok = (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous(memory_format=torch.channels_last)
          and h * w <= 1600 and c % 32 == 0 and cpg in (4, 8, 16, 32) and gn.weight is not None)
    if not ok:
        y = torch.nn.functional.group_norm(x, gn.num_groups, gn.weight, gn.bias, gn.eps)
        return torch.relu(y) if relu else y
    return _GroupNormReLUNHWC.apply(x, gn.weight, gn.bias, gn.num_groups, gn.eps, relu)


def pointwise_conv(rows, packed_weight, bias, outputs, hw):
    """One GEMM over ``rows`` (TiledRows [M, K]) with ``packed_weight`` (from ``pack_weight``, same ``split``).

    outputs: list of (out NCHW fp32 tensor, residual or None, col_begin, col_end); the column ranges must tile
    [0, Nout) in order.  Returns the list of output tensors."""
    lib = _capi.lib()
    packed, nout, k, split = packed_weight
    assert isinstance(rows, TiledRows) and rows.K == k and rows.split == split, 'operands do not match'
    segs = (_capi.PointwiseSegment * len(outputs))()
    for i, (out, res, c0, c1) in enumerate(outputs):
        assert out.dtype == torch.float32 and out.is_contiguous() and out.shape[1] == c1 - c0
        assert res is None or (res.dtype == torch.float32 and res.is_contiguous() and res.shape == out.shape)
        segs[i].out = out.data_ptr()
        segs[i].residual = None if res is None else res.data_ptr()
        segs[i].col_begin, segs[i].col_end = c0, c1
        segs[i].channels_total, segs[i].channel_offset = out.shape[1], 0
    b = None if bias is None else bias.detach().float().contiguous()
    _capi.check(lib.kgdet_pointwise_conv_tiled(rows.buf.data_ptr(), packed.data_ptr(), _capi.ptr(b), rows.M, k, nout,
                                               hw, int(split), ctypes.cast(segs, ctypes.c_void_p), len(outputs),
                                               _capi.stream_of(rows.buf)), 'kgdet_pointwise_conv_tiled')
    return [o[0] for o in outputs]
