"""Plain k x k convolutions of the head towers on the tensor cores.

The reference's ``ConvModule`` (mmdet/models/utils/conv_module.py:96-110,156-164: conv -> GroupNorm -> ReLU;
KP3:292-313 the 3 + 3 tower modules, KP3:98-106 the two stage-1 3x3 convolutions) runs ``nn.Conv2d`` through
cuDNN.  On this GPU cuDNN's default (TF32) rounds both operands to 11 bits -- measured 8e-4 on the stage-1
outputs, amplified by the two deformable stages to 1e-1 at stage 3 against the reference golden -- and its fp32
path takes 1.5 ms per convolution.  ``conv_planes`` is the library's own implicit-GEMM kernel
(kgdet_conv_forward: TMA tile loads + tcgen05, "bf16x3" split precision, fp32-grade) on SPLIT PLANES:

    SplitPlanes          channel-blocked bf16 planes of an activation's hi parts + of its lo parts; the hi half is
                         bit-identical to the fused DCN kernel's prepared input (``as_prepared_input``)
    split_planes(x)      from an NCHW or channels_last fp32 tensor
    groupnorm_relu_planes(y, gn)   GroupNorm + ReLU of a convolution's NHWC output, written as SplitPlanes
    conv_planes(planes, weight, bias, relu)   -> channels_last fp32 tensor [N, Cout, H, W]
"""
import torch

from . import _capi
from .pointwise import cached


class SplitPlanes(object):
    """[hi planes | lo planes] of an activation [N, C, H, W] (opaque buffer, kgdet_conv_split_planes_bytes)."""
    __slots__ = ('buf', 'shape4')

    def __init__(self, shape4, device):
        lib = _capi.lib()
        n, c, h, w = shape4
        nbytes = int(lib.kgdet_conv_split_planes_bytes(n, c, h, w))
        if nbytes == 0:
            raise ValueError('split planes need C %% 64 == 0, got %r' % (tuple(shape4),))
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.shape4 = tuple(shape4)

    def as_prepared_input(self, out_channels, kernel_size=3, padding=1):
        """The hi half as the PreparedInput of the fused bf16 deformable convolution (no copy)."""
        from .dcn import PreparedInput, _geom_shape, ctypes_ref
        from torch.nn.modules.utils import _pair
        lib = _capi.lib()
        n, c, h, w = self.shape4
        shape = _geom_shape(n, c, h, w, out_channels, _pair(kernel_size), (1, 1), _pair(padding), (1, 1))
        half = int(lib.kgdet_dcn_prepared_input_bytes(ctypes_ref(shape), _capi.PREC_BF16))
        assert 2 * half == self.buf.numel(), 'split-plane layout does not match the DCN prepared-input layout'
        p = PreparedInput()
        p.shape4 = self.shape4
        p.dtype_code = _capi.F32
        p.torch_dtype = torch.float32
        p.precision = _capi.PREC_BF16
        p.fast = bool(lib.kgdet_dcn_fast_path_supported(ctypes_ref(shape), _capi.PREC_BF16))
        assert p.fast, 'the fused DCN path does not support this shape'
        p.buf = self.buf[:half]
        return p

    def to_dense(self):
        """fp32 [N, C, H, W] value (hi + lo) -- for tests."""
        n, c, h, w = self.shape4
        half = self.buf.numel() // 2
        planes = c // 64
        plane_bytes = half // planes
        guard = (w + 2) * 128
        out = None
        for part in (self.buf[:half], self.buf[half:]):
            v = part.view(planes, plane_bytes)[:, guard:guard + n * h * w * 128].contiguous().view(torch.bfloat16)
            v = v.view(planes, n, h, w, 64).permute(1, 0, 4, 2, 3).reshape(n, c, h, w).float()
            out = v if out is None else out + v
        return out


def conv_supported(in_channels, out_channels, kernel_size):
    return bool(_capi.lib().kgdet_conv_supported(int(in_channels), int(out_channels), int(kernel_size)))


def split_planes(x):
    """fp32 [N, C, H, W] (NCHW-contiguous or channels_last) -> SplitPlanes."""
    lib = _capi.lib()
    _capi.require_cuda(x, 'split_planes')
    x = x.detach()
    assert x.dim() == 4 and x.dtype == torch.float32
    n, c, h, w = x.shape
    sp = SplitPlanes((n, c, h, w), x.device)
    if not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last):
        _capi.check(lib.kgdet_conv_split_planes_from_rows(x.data_ptr(), sp.buf.data_ptr(), n, c, h, w, _capi.stream_of(x)),
                    'kgdet_conv_split_planes_from_rows')
    else:
        x = x.contiguous()
        _capi.check(lib.kgdet_conv_split_planes_from_nchw(x.data_ptr(), sp.buf.data_ptr(), n, c, h, w, _capi.stream_of(x)),
                    'kgdet_conv_split_planes_from_nchw')
    return sp


def pack_conv_weight(weight):
    """fp32 [Cout, Cin, k, k] -> packed operand (cached per parameter version for inference)."""
    def build():
        lib = _capi.lib()
        w = weight.detach().float().contiguous()
        cout, cin, k, k2 = w.shape
        assert k == k2
        packed = torch.empty(int(lib.kgdet_conv_packed_weight_bytes(cout, cin, k)), dtype=torch.uint8, device=w.device)
        _capi.check(lib.kgdet_conv_pack_weight(w.data_ptr(), packed.data_ptr(), cout, cin, k, _capi.stream_of(w)),
                    'kgdet_conv_pack_weight')
        return packed
    return cached((weight,), build, tag='conv_umma')


def conv_planes(planes, weight, bias=None, relu=False):
    """Stride-1 "same" convolution of `planes` with `weight` [Cout, Cin, k, k] (+ bias, ReLU).  Returns a
    channels_last fp32 tensor of logical shape [N, Cout, H, W]."""
    lib = _capi.lib()
    n, c, h, w = planes.shape4
    cout, cin, k, _ = weight.shape
    assert cin == c, 'weight does not match the planes'
    packed = pack_conv_weight(weight)
    out = torch.empty((n, cout, h, w), dtype=torch.float32, device=planes.buf.device,
                      memory_format=torch.channels_last)
    b = None if bias is None else bias.detach().float().contiguous()
    _capi.check(lib.kgdet_conv_forward(planes.buf.data_ptr(), packed.data_ptr(), _capi.ptr(b), out.data_ptr(), n, c, h, w,
                                       cout, k, int(bool(relu)), _capi.stream_of(planes.buf)), 'kgdet_conv_forward')
    return out


def conv_planes_pair(planes0, weight0, planes1, weight1, bias0=None, bias1=None, relu=False):
    """Two stride-1 "same" convolutions of the same geometry in one launch (kgdet_conv_forward_pair): the two towers'
    convolutions of a layer.  Returns the two channels_last fp32 outputs; bit-identical to two `conv_planes` calls."""
    lib = _capi.lib()
    n, c, h, w = planes0.shape4
    assert planes1.shape4 == planes0.shape4 and tuple(weight0.shape) == tuple(weight1.shape)
    cout, cin, k, _ = weight0.shape
    assert cin == c, 'weight does not match the planes'
    p0, p1 = pack_conv_weight(weight0), pack_conv_weight(weight1)
    dev = planes0.buf.device
    out0 = torch.empty((n, cout, h, w), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
    out1 = torch.empty((n, cout, h, w), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
    b0 = None if bias0 is None else bias0.detach().float().contiguous()
    b1 = None if bias1 is None else bias1.detach().float().contiguous()
    _capi.check(lib.kgdet_conv_forward_pair(planes0.buf.data_ptr(), p0.data_ptr(), _capi.ptr(b0), out0.data_ptr(),
                                            planes1.buf.data_ptr(), p1.data_ptr(), _capi.ptr(b1), out1.data_ptr(),
                                            n, c, h, w, cout, k, int(bool(relu)), _capi.stream_of(planes0.buf)),
                'kgdet_conv_forward_pair')
    return out0, out1


def groupnorm_relu_planes(x, gn, relu=True, also_dense=False):
    """GroupNorm (+ ReLU) of a channels_last fp32 activation, written as SplitPlanes (and, with `also_dense`,
    also returned as a channels_last fp32 tensor).  `gn` is the torch.nn.GroupNorm module."""
    lib = _capi.lib()
    _capi.require_cuda(x, 'groupnorm_relu_planes')
    assert x.dim() == 4 and x.dtype == torch.float32 and x.is_contiguous(memory_format=torch.channels_last)
    n, c, h, w = x.shape
    sp = SplitPlanes((n, c, h, w), x.device)
    y = torch.empty_like(x, memory_format=torch.channels_last) if also_dense else None
    _capi.check(lib.kgdet_groupnorm_relu_nhwc_planes(x.data_ptr(), gn.weight.detach().float().contiguous().data_ptr(),
                                                     gn.bias.detach().float().contiguous().data_ptr(), float(gn.eps),
                                                     int(gn.num_groups), int(bool(relu)), _capi.ptr(y), sp.buf.data_ptr(),
                                                     n, h, w, c, _capi.stream_of(x)), 'kgdet_groupnorm_relu_nhwc_planes')
    return (sp, y) if also_dense else sp
