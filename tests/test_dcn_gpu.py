"""GPU parity tests of the deformable convolution (call path: Python mirror -> ctypes -> C ABI).

Tolerances (BASELINE.json north_star): rel 1e-5 for the fp32-grade modes ('fp32', 'tf32x3'),
1e-2 for 'bf16'; 'tf32' (single pass) is checked at 5e-3.  rel = max|d| / max|ref|.
"""
import os

import pytest
import torch

from oracle import build_ref, dcn_oracle
from tests._data import dcn_case, rel_err

pytestmark = pytest.mark.gpu

# 'tf32x3' is fp32-grade per product (dropped term ~2^-22); the tensor core accumulates with truncation
# (~2e-7 of the result per 32-channel k-block absorbed by one TMEM accumulator), so the kernel promotes: no
# accumulator absorbs more than 16 k-blocks and the chunk results are summed in fp32 (dcn_umma.cu) -> 1e-5,
# the default for fp32 tensors.
TOL = {'fp32': 1e-5, 'tf32x3': 1e-5, 'tf32': 5e-3, 'bf16': 1e-2}
# per-tensor bars of the full-size comparison with the reference CUDA op (north_star: rel 1e-5 fp32, 1e-2 bf16).
# 'tf32x3' runs the forward on the tensor cores and the backward on the exact SIMT kernels.
TOL_FULL = {'fp32': dict(out=1e-5, grad_input=1e-5, grad_offset=1e-5, grad_weight=1e-5),
            'tf32x3': dict(out=TOL['tf32x3'], grad_input=1e-5, grad_offset=1e-5, grad_weight=1e-5),
            'bf16': dict(out=1e-2, grad_input=1e-2, grad_offset=1e-2, grad_weight=1e-2)}


def _oracle(d, dtype=torch.float64):
    c = lambda t: None if t is None else t.to(dtype)
    out = dcn_oracle.deform_conv_forward(c(d['x']), c(d['offset']), c(d['weight']), d['stride'], d['padding'],
                                         d['dilation'], d['groups'], d['deformable_groups'],
                                         mask=c(d['mask']), bias=c(d['bias']))
    bw = dcn_oracle.deform_conv_backward(c(d['x']), c(d['offset']), c(d['weight']), c(d['grad_out']),
                                         d['stride'], d['padding'], d['dilation'], d['groups'],
                                         d['deformable_groups'], mask=c(d['mask']),
                                         with_bias=d['bias'] is not None)
    return out, bw


def _ours(d, precision, need_bw=True):
    from kgdet_b200 import ops
    ops.set_precision(precision)
    try:
        dev = 'cuda'
        x = d['x'].to(dev).requires_grad_(need_bw)
        off = d['offset'].to(dev).requires_grad_(need_bw)
        w = d['weight'].to(dev).requires_grad_(need_bw)
        if d['mask'] is None:
            out = ops.deform_conv(x, off, w, d['stride'], d['padding'], d['dilation'], d['groups'],
                                  d['deformable_groups'])
            m = b = None
        else:
            m = d['mask'].to(dev).requires_grad_(need_bw)
            b = None if d['bias'] is None else d['bias'].to(dev).requires_grad_(need_bw)
            out = ops.modulated_deform_conv(x, off, m, w, b, d['stride'], d['padding'], d['dilation'],
                                            d['groups'], d['deformable_groups'])
        res = dict(out=out.detach())
        if need_bw:
            out.backward(d['grad_out'].to(dev))
            res.update(grad_input=x.grad, grad_offset=off.grad, grad_weight=w.grad)
            if m is not None:
                res['grad_mask'] = m.grad
            if b is not None:
                res['grad_bias'] = b.grad
        torch.cuda.synchronize()
        return res
    finally:
        ops.set_precision(None)


GENERIC_CASES = [
    dict(N=2, C=8, H=7, W=9, Cout=6, k=3),
    dict(N=2, C=8, H=7, W=9, Cout=8, k=3, groups=2, dg=2, mask=True, bias=True),
    dict(N=1, C=4, H=9, W=8, Cout=4, k=5, stride=2),
    dict(N=2, C=6, H=6, W=7, Cout=4, k=3, dil=2, dg=3, mask=True),
    dict(N=3, C=70, H=5, W=6, Cout=66, k=3, dg=2),      # ragged tiles: C, Cout not multiples of 32/64
    dict(N=1, C=16, H=3, W=3, Cout=16, k=3, offset_std=6.0),   # most samples outside the map
    dict(N=2, C=8, H=7, W=9, Cout=6, k=1, pad=0),
]


@pytest.mark.parametrize('case', GENERIC_CASES)
def test_simt_exact_path_matches_oracle(case):
    d = dcn_case(**case)
    ref_out, ref_bw = _oracle(d)
    got = _ours(d, 'fp32')
    assert rel_err(got['out'], ref_out) < TOL['fp32']
    for k in ('grad_input', 'grad_offset', 'grad_weight', 'grad_mask', 'grad_bias'):
        if k in ref_bw:
            assert rel_err(got[k], ref_bw[k]) < TOL['fp32'], k


UMMA_CASES = [
    dict(N=2, C=64, H=9, W=11, Cout=64, k=3),
    dict(N=1, C=128, H=13, W=21, Cout=192, k=3),          # TMEM alloc rounded up to 256 columns
    dict(N=2, C=256, H=13, W=21, Cout=256, k=3),          # P6 shape
    dict(N=1, C=256, H=7, W=11, Cout=256, k=5),           # P7 shape, 25 points
    dict(N=1, C=256, H=7, W=11, Cout=256, k=7),           # 49 points
    dict(N=3, C=64, H=25, W=42, Cout=128, k=3, mask=True, bias=True),   # modulated through the fused path
    dict(N=2, C=128, H=13, W=21, Cout=192, k=3),          # Cout not a multiple of 128: uneven epilogue chunks
    dict(N=1, C=64, H=9, W=10, Cout=64, k=1, pad=0),      # a single tap, a single k-block per channel block
]


@pytest.mark.parametrize('precision', ['bf16', 'tf32x3', 'tf32'])
@pytest.mark.parametrize('case', UMMA_CASES)
def test_fused_tensor_core_forward_matches_oracle(case, precision):
    from kgdet_b200.ops import _capi
    import ctypes
    d = dcn_case(**case)
    ref_out, _ = _oracle(d)
    got = _ours(d, precision, need_bw=False)
    assert rel_err(got['out'], ref_out) < TOL[precision]


def test_fast_path_is_selected_for_kgdet_shapes():
    import ctypes
    from kgdet_b200.ops import _capi
    lib = _capi.lib()
    for k in (3, 5, 7):
        s = _capi.DcnShape(N=16, C=256, H=25, W=42, Cout=256, kh=k, kw=k, stride_h=1, stride_w=1,
                           pad_h=k // 2, pad_w=k // 2, dil_h=1, dil_w=1, groups=1, deformable_groups=1)
        for p in (_capi.PREC_TF32X3, _capi.PREC_BF16, _capi.PREC_TF32):
            assert lib.kgdet_dcn_fast_path_supported(ctypes.byref(s), p) == 1
        assert lib.kgdet_dcn_fast_path_supported(ctypes.byref(s), _capi.PREC_FP32) == 0


@pytest.mark.parametrize('k', [3, 5, 7])
def test_full_size_kgdet_call_fused_vs_exact(k):
    """KGDet shape [16,256,25,42]: too big for the CPU oracle in seconds, so the fused tensor-core
    kernel is checked against the exact SIMT kernel (itself oracle-checked above) plus linearity."""
    d = dcn_case(N=16, C=256, H=25, W=42, Cout=256, k=k, seed=k)
    exact = _ours(d, 'fp32', need_bw=False)['out']
    assert rel_err(_ours(d, 'tf32x3', need_bw=False)['out'], exact) < TOL['tf32x3']
    assert rel_err(_ours(d, 'bf16', need_bw=False)['out'], exact) < TOL['bf16']
    # linearity in the input: f(2x) == 2 f(x) exactly in fp32-grade mode (power-of-two scaling)
    d2 = dict(d)
    d2['x'] = d['x'] * 2
    a = _ours(d, 'tf32x3', need_bw=False)['out']
    b = _ours(d2, 'tf32x3', need_bw=False)['out']
    assert torch.equal(a * 2, b)
    # default precision for fp32 tensors is the fp32-grade tensor-core path
    from kgdet_b200 import ops
    assert ops.get_precision(torch.float32) == 'tf32x3' and ops.get_precision(torch.bfloat16) == 'bf16'
    assert torch.equal(_ours(d, None, need_bw=False)['out'], a)


def test_zero_offset_equals_plain_convolution():
    d = dcn_case(N=2, C=64, H=12, W=10, Cout=64, k=3)
    d['offset'] = torch.zeros_like(d['offset'])
    ref = torch.nn.functional.conv2d(d['x'].double(), d['weight'].double(), padding=1)
    for prec in ('fp32', 'tf32x3', 'bf16'):
        assert rel_err(_ours(d, prec, need_bw=False)['out'], ref) < TOL[prec]


def test_full_size_p3_zero_offset_equals_conv2d_forward_and_backward():
    """BASELINE configs[1] at its largest shape (FPN P3, batch 8: 134 400 positions -- too big for the CPU oracle)
    through a size-independent property: with zero offsets the deformable convolution IS the plain convolution, so
    forward, input gradient and weight gradient must match cuDNN's (fp32, TF32 off) in the bf16 mode's tolerance.
    At this size the backward takes its one-tap-per-chunk path and the fused weight gradient splits 2 100 position
    blocks over the SMs."""
    from kgdet_b200 import ops
    g = torch.Generator().manual_seed(31)
    x = torch.randn(8, 256, 100, 168, generator=g).cuda()
    w = (torch.randn(256, 256, 3, 3, generator=g) * 0.02).cuda()
    go = torch.randn(8, 256, 100, 168, generator=g).cuda()
    off = torch.zeros(8, 18, 100, 168, device='cuda')
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
        ref = torch.nn.functional.conv2d(xr, wr, padding=1)
        ref.backward(go)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    ops.set_precision('bf16')
    try:
        xo, wo, oo = x.clone().requires_grad_(), w.clone().requires_grad_(), off.clone().requires_grad_()
        out = ops.deform_conv(xo, oo, wo, 1, 1)
        out.backward(go)
    finally:
        ops.set_precision(None)
    assert rel_err(out, ref) < TOL['bf16']
    assert rel_err(xo.grad, xr.grad) < TOL['bf16']
    assert rel_err(wo.grad, wr.grad) < TOL['bf16']
    assert torch.isfinite(oo.grad).all()


def test_bf16_tensors_end_to_end():
    d = dcn_case(N=2, C=64, H=9, W=11, Cout=64, k=3, dtype=torch.bfloat16)
    d['weight'] = d['weight'].to(torch.bfloat16).float()
    ref_out, ref_bw = _oracle(d)
    got = _ours(d, None)
    assert got['out'].dtype == torch.bfloat16
    assert rel_err(got['out'], ref_out) < 1e-2
    assert rel_err(got['grad_input'], ref_bw['grad_input']) < 1e-2
    assert rel_err(got['grad_offset'], ref_bw['grad_offset']) < 1e-2
    assert rel_err(got['grad_weight'], ref_bw['grad_weight']) < 1e-2


def test_error_behaviour_matches_reference():
    from kgdet_b200 import ops
    x = torch.randn(2, 8, 5, 5)
    with pytest.raises(NotImplementedError):                      # DC.py:44-45
        ops.deform_conv(x, torch.zeros(2, 18, 5, 5), torch.randn(4, 8, 3, 3), 1, 1)
    with pytest.raises(ValueError):                               # DC.py:25-28
        ops.deform_conv(x[0].cuda(), torch.zeros(18, 5, 5).cuda(), torch.randn(4, 8, 3, 3).cuda(), 1, 1)
    with pytest.raises(RuntimeError):                             # DC.cpp:128-135
        ops.deform_conv(x.cuda(), torch.zeros(2, 16, 5, 5).cuda(), torch.randn(4, 8, 3, 3).cuda(), 1, 1)
    with pytest.raises(AssertionError):                           # DC.py:48-49
        ops.deform_conv(x.cuda().repeat(3, 1, 1, 1)[:5], torch.zeros(5, 18, 5, 5).cuda(),
                        torch.randn(4, 8, 3, 3).cuda(), 1, 1, 1, 1, 1, 2)
    m = ops.DeformConv(8, 4, 3, padding=1)
    assert [n for n, _ in m.named_parameters()] == ['weight'] and not hasattr(m, 'bias')


def test_oracle_matches_reference_cuda():
    """Pins oracle/dcn_oracle.py against the reference's own deform_conv_cuda sources compiled
    unmodified for sm_100a (oracle/_ref, built by oracle/build_ref.py)."""
    ref = build_ref.load('deform_conv_cuda')
    if ref is None:
        pytest.skip('oracle/_ref/deform_conv_cuda.so not built (needs /root/reference at build time)')
    d = dcn_case(N=4, C=16, H=9, W=11, Cout=12, k=3, dg=2)
    dev = 'cuda'
    x, off, w, go = (d[k].to(dev) for k in ('x', 'offset', 'weight', 'grad_out'))
    out = x.new_empty(4, 12, 9, 11)
    bufs = [x.new_empty(0), x.new_empty(0)]
    # argument order of DC.py:50-55 (W before H)
    ref.deform_conv_forward_cuda(x, w, off, out, bufs[0], bufs[1], 3, 3, 1, 1, 1, 1, 1, 1, 1, 2, 4)
    gi, goff, gw = torch.zeros_like(x), torch.zeros_like(off), torch.zeros_like(w)
    ref.deform_conv_backward_input_cuda(x, off, go, gi, goff, w, bufs[0], 3, 3, 1, 1, 1, 1, 1, 1, 1, 2, 4)
    # im2col_step = 1 here: with a larger step the reference's zeros_like(transposed).view(...)
    # (deform_conv_cuda.cpp:423-430) raises on torch >= 1.5 (zeros_like keeps the transposed strides)
    ref.deform_conv_backward_parameters_cuda(x, off, go, gw, bufs[0], bufs[1], 3, 3, 1, 1, 1, 1, 1, 1, 1, 2,
                                             1, 1)
    torch.cuda.synchronize()
    o_out, o_bw = _oracle(d, torch.float64)
    assert rel_err(out, o_out) < 1e-5
    assert rel_err(gi, o_bw['grad_input']) < 1e-5
    assert rel_err(goff, o_bw['grad_offset']) < 1e-5
    assert rel_err(gw, o_bw['grad_weight']) < 1e-5
    # and ours against the reference kernel directly
    got = _ours(d, 'fp32')
    assert rel_err(got['out'], out) < 1e-5
    assert rel_err(got['grad_offset'], goff) < 1e-5


@pytest.mark.parametrize('precision', ['fp32', 'bf16', 'tf32x3'])
def test_prepared_api_matches_plain_calls(precision):
    """Shared NHWC copy + shared plan + fused ReLU / channel-slice epilogue == relu(deform_conv) + cat,
    bit for bit (same kernels, same arithmetic).  The plain call would split these tiny maps over k-blocks
    (different fp32 summation order, test_split_k_forward_equals_unsplit); the split is switched off here so that
    the comparison stays bitwise."""
    from kgdet_b200 import ops
    ops.set_precision(precision)
    os.environ['KGDET_UMMA_SPLITS'] = '1'
    try:
        d3 = dcn_case(N=2, C=64, H=12, W=10, Cout=64, k=3, seed=1)
        d5 = dcn_case(N=2, C=64, H=12, W=10, Cout=64, k=5, seed=2)
        x = d3['x'].cuda()
        pin = ops.prepare_input(x, 64)
        out = torch.full((2, 128 + 8, 12, 10), -7.0, device='cuda')
        ref = []
        for i, d in enumerate((d3, d5)):
            k = d['weight'].shape[-1]
            off, w = d['offset'].cuda(), d['weight'].cuda()
            plan = ops.prepare_plan(off, x.shape, 64, k, 1, k // 2, 1)
            ops.deform_conv_prepared(pin, plan, w, out, 4 + i * 64, True)
            ref.append(torch.relu(ops.deform_conv(x, off, w, 1, k // 2)))
        assert torch.equal(out[:, 4:132], torch.cat(ref, 1))
        assert (out[:, :4] == -7).all() and (out[:, 132:] == -7).all()       # neighbours untouched
        plain = ops.deform_conv_prepared(pin, ops.prepare_plan(d3['offset'].cuda(), x.shape, 64, 3, 1, 1, 1),
                                         d3['weight'].cuda())
        assert torch.equal(plain, ops.deform_conv(x, d3['offset'].cuda(), d3['weight'].cuda(), 1, 1))
    finally:
        del os.environ['KGDET_UMMA_SPLITS']
        ops.set_precision(None)


BWD_TC_CASES = [
    dict(N=2, C=64, H=9, W=11, Cout=64, k=3),
    dict(N=2, C=256, H=13, W=21, Cout=256, k=3),
    dict(N=1, C=256, H=7, W=11, Cout=256, k=5),
    dict(N=1, C=128, H=7, W=11, Cout=192, k=7),
    dict(N=3, C=64, H=25, W=42, Cout=128, k=3, mask=True, bias=True),
    dict(N=2, C=128, H=9, W=11, Cout=64, k=3, mask=True, bias=True),     # bulk-reduction col2im (C = 128) with a mask
    dict(N=1, C=256, H=6, W=7, Cout=64, k=5, mask=True, offset_std=5.0),  # ... C = 256, many samples outside the map
]


@pytest.mark.parametrize('case', BWD_TC_CASES)
def test_tensor_core_backward_matches_oracle(case):
    """bf16 mode: column-gradient / weight-gradient GEMMs on tcgen05, col2im with warp-shuffle offset
    reduction.  Tolerance: rel 1e-2 (bf16)."""
    d = dcn_case(**case)
    ref_out, ref_bw = _oracle(d)
    got = _ours(d, 'bf16')
    assert rel_err(got['out'], ref_out) < TOL['bf16']
    for k in ('grad_input', 'grad_offset', 'grad_weight', 'grad_mask', 'grad_bias'):
        if k in ref_bw:
            assert rel_err(got[k], ref_bw[k]) < TOL['bf16'], (k, rel_err(got[k], ref_bw[k]))


@pytest.mark.parametrize('case', [BWD_TC_CASES[1], BWD_TC_CASES[4],
                                  dict(N=2, C=64, H=25, W=42, Cout=64, k=3, offset_std=0.3)])
def test_owned_slice_col2im_matches_oracle_and_default_path(case, monkeypatch):
    """KGDET_COL2IM_OWN=1 (dcn_col2im_own.cu: the CTA owns a 32-channel slice of the input gradient in shared
    memory, lane schedule precomputed per offset tensor; opt-in because it measured slower): same gradients as the
    oracle and as the default red.global col2im, grad_offset / grad_mask bitwise reproducible.  The third case has
    small offsets: neighbouring positions hit the same pixels, the turn-taking path is exercised."""
    d = dcn_case(**case)
    ref_out, ref_bw = _oracle(d)
    default = _ours(d, 'bf16')
    monkeypatch.setenv('KGDET_COL2IM_OWN', '1')
    got = _ours(d, 'bf16')
    again = _ours(d, 'bf16')
    for k in ('grad_input', 'grad_offset', 'grad_mask'):
        if k in ref_bw:
            assert rel_err(got[k], ref_bw[k]) < TOL['bf16'], (k, rel_err(got[k], ref_bw[k]))
            assert rel_err(got[k], default[k]) < 1e-4, (k, rel_err(got[k], default[k]))
    assert torch.equal(got['grad_offset'], again['grad_offset'])
    monkeypatch.setenv('KGDET_COL2IM_OWN_PERMUTE', '0')
    plain = _ours(d, 'bf16')
    assert rel_err(plain['grad_input'], default['grad_input']) < 1e-4


def test_tensor_core_backward_full_size_vs_exact():
    d = dcn_case(N=16, C=256, H=25, W=42, Cout=256, k=5, seed=11)
    exact = _ours(d, 'fp32')
    fast = _ours(d, 'bf16')
    for k in ('grad_input', 'grad_offset', 'grad_weight'):
        assert rel_err(fast[k], exact[k]) < TOL['bf16'], (k, rel_err(fast[k], exact[k]))
    # grad_offset of the tensor-core path has a single writer per element: bitwise reproducible
    again = _ours(d, 'bf16')
    assert torch.equal(fast['grad_offset'], again['grad_offset'])
    # the weight gradient above came from the gather-fused kernel (dcn_wgrad_umma.cu: C = 256); the unfused
    # path (transposed columns + split-K GEMM) must agree with it to bf16 rounding of the sampled values
    os.environ['KGDET_WGRAD_FUSED'] = '0'
    try:
        unfused = _ours(d, 'bf16')
    finally:
        del os.environ['KGDET_WGRAD_FUSED']
    assert rel_err(unfused['grad_weight'], exact['grad_weight']) < TOL['bf16']
    assert rel_err(unfused['grad_weight'], fast['grad_weight']) < 5e-3


@pytest.mark.parametrize('precision,tol', [('bf16', 1e-2), ('tf32x3', 2e-4), ('tf32', 5e-3)])   # pair launches never split: tf32x3 unpromoted
def test_cta_pair_variant_matches_default(monkeypatch, precision, tol):
    """The 2-SM (cta_group::2) variant of the fused kernel -- opt-in via KGDET_UMMA_PAIR=1 -- against the
    fp64 oracle and, bit for bit, against the default 1-CTA launch (same arithmetic, same order).  Odd tile
    count (M = 2*13*21 = 546 -> 5 tiles) so the padded sixth tile of the last pair is exercised."""
    from kgdet_b200 import ops
    d = dcn_case(N=2, C=128, H=13, W=21, Cout=192, k=3, seed=11)
    x, off, w = (d[q].cuda() for q in ('x', 'offset', 'weight'))
    ops.set_precision(precision)
    monkeypatch.setenv('KGDET_UMMA_SPLITS', '1')       # the pair variant never splits the k-blocks; keep both launches whole
    monkeypatch.setenv('KGDET_TF32X3_CHUNK', '0')      # ... and switch the tf32x3 accumulator promotion (a k-block split) off
    try:
        monkeypatch.setenv('KGDET_UMMA_PAIR', '0')
        base = ops.deform_conv(x, off, w, 1, 1)
        monkeypatch.setenv('KGDET_UMMA_PAIR', '1')
        pair = ops.deform_conv(x, off, w, 1, 1)
    finally:
        ops.set_precision(None)
    ref = dcn_oracle.deform_conv_forward(d['x'].double(), d['offset'].double(), d['weight'].double(), 1, 1)
    assert rel_err(pair, ref) < tol
    assert torch.equal(pair, base)


@pytest.mark.parametrize('precision,tol', [('bf16', 1e-2), ('tf32', 5e-3)])
def test_split_k_forward_equals_unsplit(precision, tol):
    """Small maps split the k-blocks of a tile over several CTAs (dcn_umma.cu: umma_splits) and combine the fp32
    partial tiles in a fixed order: same result as the one-CTA-per-tile kernel up to fp32 summation order, and
    both within the mode's tolerance of the oracle; bias, bf16 output and a ragged last tile included."""
    from kgdet_b200 import ops
    d = dcn_case(N=2, C=128, H=13, W=21, Cout=192, k=5, seed=5)
    ref_out, _ = _oracle(d)
    x, off, w = (d[q].cuda() for q in ('x', 'offset', 'weight'))
    ops.set_precision(precision)
    try:
        split = ops.deform_conv(x, off, w, 1, 2)
        os.environ['KGDET_UMMA_SPLITS'] = '1'
        try:
            whole = ops.deform_conv(x, off, w, 1, 2)
        finally:
            del os.environ['KGDET_UMMA_SPLITS']
        os.environ['KGDET_UMMA_SPLITS'] = '7'          # uneven: 50 k-blocks in 7 splits of 8 (last one 2)
        try:
            uneven = ops.deform_conv(x, off, w, 1, 2)
        finally:
            del os.environ['KGDET_UMMA_SPLITS']
    finally:
        ops.set_precision(None)
    assert rel_err(split, ref_out) < tol and rel_err(whole, ref_out) < tol
    assert rel_err(split, whole) < 1e-5 and rel_err(uneven, whole) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# Full-size BASELINE shapes against the REFERENCE'S OWN CUDA op (oracle/_ref/deform_conv_cuda.so: the sources
# of mmdet/ops/dcn/src compiled unmodified for sm_100a by oracle/build_ref.py), forward and all three gradients,
# non-zero offsets with a share of the samples outside the map.  These shapes are too big for the CPU oracle in
# seconds; the reference kernel itself is pinned to the oracle in test_oracle_matches_reference_cuda.
# ---------------------------------------------------------------------------------------------------------
def _reference_cuda(d, dtype=torch.float32):
    """forward / backward_input / backward_parameters of the reference extension, driven with the argument order
    of mmdet/ops/dcn/deform_conv.py:50-55,75-92 (W before H; buffers re-allocated by the C++).  `dtype` float64
    runs the same reference kernels in double (AT_DISPATCH_FLOATING_TYPES_AND_HALF, deform_conv_cuda_kernel.cu:258)."""
    # float64: the build with -maxrregcount=64 (the default build cannot launch its double kernels with 1024 threads)
    ref = build_ref.load('deform_conv_cuda_r64' if dtype == torch.float64 else 'deform_conv_cuda')
    dev = 'cuda'
    x, off, w, go = (d[k].to(dev).to(dtype).contiguous() for k in ('x', 'offset', 'weight', 'grad_out'))
    N, k, pad = x.shape[0], w.shape[-1], d['padding']
    out = x.new_empty(go.shape)
    bufs = [x.new_empty(0), x.new_empty(0)]
    step = min(64, N)                                             # DC.py:47
    ref.deform_conv_forward_cuda(x, w, off, out, bufs[0], bufs[1], k, k, 1, 1, pad, pad, 1, 1, 1, 1, step)
    gi, goff, gw = torch.zeros_like(x), torch.zeros_like(off), torch.zeros_like(w)
    ref.deform_conv_backward_input_cuda(x, off, go, gi, goff, w, bufs[0], k, k, 1, 1, pad, pad, 1, 1, 1, 1, step)
    # im2col_step 1: the reference's zeros_like(transposed).view(...) raises for larger steps on torch >= 1.5
    ref.deform_conv_backward_parameters_cuda(x, off, go, gw, bufs[0], bufs[1], k, k, 1, 1, pad, pad, 1, 1, 1, 1,
                                             1, 1)
    torch.cuda.synchronize()
    return dict(out=out, grad_input=gi, grad_offset=goff, grad_weight=gw)


FULL_SIZE_CASES = [
    # BASELINE configs[1]: isolated 3x3 DeformConv over FPN P3..P7 at 800x1333, batch 8
    dict(N=8, C=256, H=100, W=168, Cout=256, k=3),
    dict(N=8, C=256, H=50, W=84, Cout=256, k=3),
    dict(N=8, C=256, H=25, W=42, Cout=256, k=3),
    dict(N=8, C=256, H=13, W=21, Cout=256, k=3),
    dict(N=8, C=256, H=7, W=11, Cout=256, k=3),
    # the KGDet head's own calls (configs[2]): 9 / 25 / 49 points at batch 16 on the 25x42 map
    dict(N=16, C=256, H=25, W=42, Cout=256, k=3),
    dict(N=16, C=256, H=25, W=42, Cout=256, k=5),
    dict(N=16, C=256, H=25, W=42, Cout=256, k=7),
]


@pytest.mark.parametrize('case', FULL_SIZE_CASES, ids=lambda c: '%dx%dx%d_k%d' % (c['N'], c['H'], c['W'], c['k']))
def test_full_size_shapes_match_reference_cuda_op(case):
    """Truth = the reference kernels run in float64; the reference's own float32 result (cuBLAS SGEMM + fp32
    atomics, what a user of the reference gets) is measured against it too.  Bar per tensor: the north_star
    tolerance of the mode, or -- where the reference's own fp32 result is further than that from the truth (the
    weight gradient of the P3 map sums 134 400 positions in fp32) -- 1.5x the reference's own error."""
    if build_ref.load('deform_conv_cuda') is None or build_ref.load('deform_conv_cuda_r64') is None:
        pytest.skip('oracle/_ref/deform_conv_cuda{,_r64}.so not built (needs /root/reference at build time)')
    d = dcn_case(seed=100 + case['k'] + case['H'], **case)
    # offsets N(0, 2^2) px: the outermost ring of samples falls outside the map (zero-padding rule exercised)
    want = {k: v.cpu() for k, v in _reference_cuda(d, torch.float64).items()}
    torch.cuda.empty_cache()
    ref32 = _reference_cuda(d, torch.float32)
    ref_err = {k: rel_err(ref32[k], want[k]) for k in want}
    del ref32
    torch.cuda.empty_cache()
    report = {}
    for precision in ('fp32', 'tf32x3', 'bf16'):
        got = _ours(d, precision)
        for key in ('out', 'grad_input', 'grad_offset', 'grad_weight'):
            e = rel_err(got[key], want[key])
            report[(precision, key)] = e
            assert e < max(TOL_FULL[precision][key], 1.5 * ref_err[key]), (precision, key, e, ref_err[key])
        del got
        torch.cuda.empty_cache()
    print('full-size parity', case, {'reference_fp32': ref_err}, report)
