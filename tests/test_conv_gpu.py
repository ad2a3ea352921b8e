"""GPU parity tests of the tensor-core tower convolution (kgdet_conv_forward, csrc/conv_umma.cu) and its
split-plane producers, against torch's fp64 convolution / GroupNorm.  Call path: Python mirror -> ctypes -> C ABI.

Tolerance: the kernel is "bf16x3" (bf16 hi + lo operands, three MMAs per k-step, fp32 accumulation): fp32-grade,
checked at rel 2e-5 of the tensor maximum (cuDNN's TF32 path measures 8e-4 on the same inputs)."""
import pytest
import torch
import torch.nn.functional as F

from tests._data import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-5

CASES = [
    # (N, C, H, W, Cout, k)
    (16, 256, 25, 42, 256, 3),      # the KGDet tower / stage-1 convolutions at the benchmark batch
    (2, 256, 25, 42, 256, 3),       # training batch: 18 tiles
    (1, 64, 7, 11, 64, 3),          # a single (partial) tile, one channel block, odd tile count -> padding CTA
    (3, 128, 13, 21, 192, 3),       # Cout = 192 (TMEM allocation rounded up), odd tile count
    (2, 64, 100, 168, 128, 3),      # FPN P3 width > 128: tiles cut rows
    (2, 128, 9, 10, 64, 1),         # 1x1
    (1, 64, 12, 17, 64, 5),         # 5x5, padding 2
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: 'x'.join(map(str, c)))
@pytest.mark.parametrize('layout', ['nchw', 'channels_last'])
def test_conv_planes_matches_fp64_convolution(case, layout):
    from kgdet_b200.ops import conv
    N, C, H, W, Cout, k = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(Cout, C, k, k, generator=g) * (1.0 / (C * k * k) ** 0.5)
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=k // 2)
    xd = x.cuda()
    if layout == 'channels_last':
        xd = xd.contiguous(memory_format=torch.channels_last)
    planes = conv.split_planes(xd)
    # the planes hold x to 2^-17: hi + lo
    assert rel_err(planes.to_dense(), x) < 2e-5
    out = conv.conv_planes(planes, w.cuda(), b.cuda())
    assert out.shape == (N, Cout, H, W) and out.is_contiguous(memory_format=torch.channels_last)
    assert rel_err(out, ref) < TOL, rel_err(out, ref)
    out_relu = conv.conv_planes(planes, w.cuda(), None, relu=True)
    ref_relu = F.conv2d(x.double(), w.double(), None, padding=k // 2).clamp(min=0)
    assert rel_err(out_relu, ref_relu) < TOL


def test_conv_is_far_closer_than_tf32():
    """The reason this kernel exists: on the same inputs cuDNN's TF32 convolution is ~50x further from the truth."""
    from kgdet_b200.ops import conv
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 256, 25, 42, generator=g)
    w = torch.randn(256, 256, 3, 3, generator=g) * 0.03
    ref = F.conv2d(x.double(), w.double(), padding=1)
    ours = rel_err(conv.conv_planes(conv.split_planes(x.cuda()), w.cuda()), ref)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    try:
        tf32 = rel_err(F.conv2d(x.cuda(), w.cuda(), padding=1), ref)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert ours < TOL and tf32 > 10 * ours, (ours, tf32)


def test_groupnorm_relu_planes_matches_torch_and_feeds_the_dcn():
    """GroupNorm + ReLU written as split planes == torch GroupNorm + ReLU (fp64) to 2e-5; the dense twin output is
    bit-identical to the existing NHWC kernel; the hi half is bit-identical to the DCN's prepared input."""
    from kgdet_b200 import ops
    from kgdet_b200.ops import conv
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(3, 256, 25, 42, generator=g) * 2 + 0.3).cuda().contiguous(memory_format=torch.channels_last)
    gn = torch.nn.GroupNorm(32, 256).cuda()
    with torch.no_grad():
        gn.weight.normal_(1.0, 0.2, generator=None)
        gn.bias.normal_(0.0, 0.2)
        ref = F.relu(F.group_norm(x.double(), 32, gn.weight.double(), gn.bias.double(), gn.eps))
        sp, dense = conv.groupnorm_relu_planes(x, gn, also_dense=True)
        assert torch.equal(dense, ops.groupnorm_relu_nhwc(x, gn))
        assert rel_err(dense, ref) < 1e-5
        assert rel_err(sp.to_dense(), ref) < 2e-5
        pin = sp.as_prepared_input(256)
        want = ops.prepare_input(dense, 256, precision='bf16')
        assert torch.equal(pin.buf, want.buf)
        # ... and a deformable convolution on it equals the one on the separately prepared input
        off = (torch.randn(3, 18, 25, 42, generator=g) * 2).cuda()
        wd = (torch.randn(256, 256, 3, 3, generator=g) * 0.02).cuda()
        plan = ops.prepare_plan(off, (3, 256, 25, 42), 256, 3, 1, 1, 1, precision='bf16')
        a = ops.deform_conv_prepared(pin, plan, wd)
        b = ops.deform_conv_prepared(want, plan, wd)
        assert torch.equal(a, b)


@pytest.mark.parametrize('shape,cout,k,bias,relu', [((16, 256, 25, 42), 256, 3, False, False),
                                                    ((2, 64, 13, 21), 128, 3, True, True),
                                                    ((3, 128, 7, 11), 64, 1, True, False),
                                                    ((1, 256, 100, 168), 256, 3, False, False)])
def test_paired_convolutions_equal_two_single_launches(shape, cout, k, bias, relu):
    """kgdet_conv_forward_pair (two convolutions of one geometry per launch, the second into the other half of
    TMEM) == two kgdet_conv_forward launches, bit for bit."""
    from kgdet_b200.ops import conv as kconv
    n, c, h, w = shape
    g = torch.Generator().manual_seed(c + h + k)
    xs = [torch.randn(n, c, h, w, generator=g).cuda() for _ in range(2)]
    ws = [(torch.randn(cout, c, k, k, generator=g) * 0.05).cuda() for _ in range(2)]
    bs = [torch.randn(cout, generator=g).cuda() if bias else None for _ in range(2)]
    ps = [kconv.split_planes(x) for x in xs]
    want = [kconv.conv_planes(p, wt, b, relu) for p, wt, b in zip(ps, ws, bs)]
    got = kconv.conv_planes_pair(ps[0], ws[0], ps[1], ws[1], bs[0], bs[1], relu)
    for a, b in zip(got, want):
        assert a.is_contiguous(memory_format=torch.channels_last) and torch.equal(a, b)
    # and the same input for both problems (the first tower layer)
    got2 = kconv.conv_planes_pair(ps[0], ws[0], ps[0], ws[1], bs[0], bs[1], relu)
    assert torch.equal(got2[0], want[0])
    assert torch.equal(got2[1], kconv.conv_planes(ps[0], ws[1], bs[1], relu))
