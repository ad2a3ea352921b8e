"""GPU tests of the Kp3RepBlock pointwise stage: plan-from-points, the fused DCN
kernel's UMMA-tiled bf16 output, the tcgen05 pointwise GEMM with its NCHW / bias / residual epilogue."""
import os

import pytest
import torch

from tests._data import dcn_case, rel_err

pytestmark = pytest.mark.gpu


def test_nchw_to_tiled_is_a_relu_transpose():
    from kgdet_b200 import ops
    x = torch.randn(3, 128, 5, 9, generator=torch.Generator().manual_seed(0)).cuda()     # M = 135: ragged tile
    for relu in (False, True):
        want = (x.relu() if relu else x).permute(0, 2, 3, 1).reshape(-1, 128)
        got = ops.nchw_to_tiled(x, relu=relu, split=False)
        assert torch.equal(got.to_dense(), want.to(torch.bfloat16).float())
        sp = ops.nchw_to_tiled(x, relu=relu, split=True)
        assert rel_err(sp.to_dense(), want) < 2e-5                                      # hi + lo


@pytest.mark.parametrize('split', [False, True])
@pytest.mark.parametrize('n_hw,K,nouts', [((2, 35), 64, (13,)), ((3, 77), 768, (588, 166)), ((1, 300), 256, (7, 130, 9)),
                                          ((2, 128), 128, (256, 300))])
def test_pointwise_conv_matches_torch(n_hw, K, nouts, split):
    """Plain mode: against fp32 torch on the same bf16-rounded operands (products exact in fp32, only the
    summation order differs).  Split mode ("bf16x3"): against fp64 on the ORIGINAL fp32 operands.  Column
    segments, bias, residual, ragged row / column tiles."""
    from kgdet_b200 import ops
    n, hw = n_hw
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, K, hw, 1, generator=g).cuda()
    w = (torch.randn(sum(nouts), K, generator=g) / K ** 0.5).cuda()
    bias = torch.randn(sum(nouts), generator=g).cuda()
    rows = ops.nchw_to_tiled(x, split=split)
    xr = x.permute(0, 2, 3, 1).reshape(-1, K)
    if split:
        ref = (xr.double() @ w.double().t() + bias.double())
    else:
        ref = xr.to(torch.bfloat16).float() @ w.to(torch.bfloat16).float().t() + bias
    outs, c0 = [], 0
    for i, no in enumerate(nouts):
        res = torch.randn(n, no, hw, 1, generator=g).cuda() if i % 2 == 0 else None
        outs.append((torch.empty(n, no, hw, 1, device='cuda'), res, c0, c0 + no))
        c0 += no
    ops.pointwise_conv(rows, ops.pack_weight(w, split=split), bias, outs, hw)
    for out, res, a, b in outs:
        want = ref[:, a:b].reshape(n, hw, b - a).permute(0, 2, 1).reshape(n, b - a, hw, 1)
        if res is not None:
            want = want + res
        assert rel_err(out, want) < 2e-5
    with pytest.raises(RuntimeError):           # segments must tile [0, Nout)
        ops.pointwise_conv(rows, ops.pack_weight(w, split=split), bias, [(outs[0][0], None, 1, 1 + nouts[0])], hw)


def test_plain_bf16_pointwise_is_only_bf16_grade():
    """What the split buys: the single-pass bf16 GEMM of fp32 operands is ~2e-3 off, the split one 1e-5."""
    from kgdet_b200 import ops
    n, hw, K, no = 2, 130, 768, 200
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, K, hw, 1, generator=g).cuda()
    w = (torch.randn(no, K, generator=g) / K ** 0.5).cuda()
    ref = torch.nn.functional.conv2d(x.double(), w.double().view(no, K, 1, 1))
    errs = {}
    for split in (False, True):
        out = torch.empty(n, no, hw, 1, device='cuda')
        ops.pointwise_conv(ops.nchw_to_tiled(x, split=split), ops.pack_weight(w, split=split), None,
                           [(out, None, 0, no)], hw)
        errs[split] = rel_err(out, ref)
    assert errs[True] < 2e-5 and 1e-4 < errs[False] < 1e-2, errs


@pytest.mark.parametrize('k', [3, 5])
def test_plan_from_points_and_tiled_output(k):
    """prepare_plan_points(points, lo) == prepare_plan(points[:, lo:lo+2K] - base) bit for bit, and the tiled
    bf16 output equals the NCHW fp32 output rounded to bf16 (split: hi + lo reproduces it to 2^-17)."""
    from kgdet_b200 import ops
    d = dcn_case(N=2, C=128, H=9, W=11, Cout=128, k=k, seed=5)
    x, w = d['x'].cuda(), d['weight'].cuda()
    K = k * k
    g = torch.Generator().manual_seed(2)
    pts = (torch.randn(2, 10 + 2 * K + 6, 9, 11, generator=g) * 2).cuda()
    lo = 10
    base = torch.arange(-(k // 2), k // 2 + 1, dtype=torch.float32)
    yx = torch.stack([base.repeat_interleave(k), base.repeat(k)], 1).reshape(1, -1, 1, 1).cuda()
    ops.set_precision('bf16')
    import os
    os.environ['KGDET_UMMA_SPLITS'] = '1'     # the NCHW call would split this tiny map over k-blocks (other fp32 sum order)
    try:
        pin = ops.prepare_input(x, 128, k, 1, k // 2, 1)
        p_ref = ops.prepare_plan(pts[:, lo:lo + 2 * K] - yx, x.shape, 128, k, 1, k // 2, 1)
        p_pts = ops.prepare_plan_points(pts, lo, x.shape, 128, k, 1, k // 2, 1)
        assert torch.equal(p_ref.buf, p_pts.buf)
        # with the head's gradient-mul expression (KP3:135-143) evaluated first, in the reference's operation order
        gm = 0.1
        p_ref_gm = ops.prepare_plan((gm * pts[:, lo:lo + 2 * K] + (1 - gm) * pts[:, lo:lo + 2 * K]) - yx, x.shape, 128, k,
                                    1, k // 2, 1)
        p_pts_gm = ops.prepare_plan_points(pts, lo, x.shape, 128, k, 1, k // 2, 1, gradient_mul=gm)
        assert torch.equal(p_ref_gm.buf, p_pts_gm.buf)
        nchw = ops.deform_conv_prepared(pin, p_pts, w, relu=True)
        rows = ops.TiledRows(2 * 9 * 11, 256, False, 'cuda')
        rows.buf.zero_()
        ops.deform_conv_prepared(pin, p_pts, w, rows, 128, True)
        srows = ops.TiledRows(2 * 9 * 11, 256, True, 'cuda')
        srows.buf.zero_()
        ops.deform_conv_prepared(pin, p_pts, w, srows, 128, True)
    finally:
        ops.set_precision(None)
        del os.environ['KGDET_UMMA_SPLITS']
    full = nchw.permute(0, 2, 3, 1).reshape(-1, 128)
    dense = rows.to_dense()
    assert torch.equal(dense[:, 128:], full.to(torch.bfloat16).float())
    assert not dense[:, :128].any()                                   # only the requested channel slice is written
    sdense = srows.to_dense()
    assert rel_err(sdense[:, 128:], full) < 2e-5 and not sdense[:, :128].any()


def test_groupnorm_relu_nhwc_matches_torch():
    """GroupNorm(32) + ReLU of the towers (conv_module.py:156-164) on channels_last activations, incl. a
    single-pixel-row map and non-trivial affine parameters."""
    from kgdet_b200 import ops
    g = torch.Generator().manual_seed(7)
    gn = torch.nn.GroupNorm(32, 256).cuda()
    with torch.no_grad():
        gn.weight.copy_(1 + 0.3 * torch.randn(256, generator=g))
        gn.bias.copy_(0.2 * torch.randn(256, generator=g))
    for shape in [(2, 256, 25, 42), (3, 256, 1, 5), (1, 256, 7, 11)]:
        x = (torch.randn(*shape, generator=g) * 3 + 1).cuda().contiguous(memory_format=torch.channels_last)
        for relu in (True, False):
            got = ops.groupnorm_relu_nhwc(x, gn, relu=relu)
            want = torch.nn.functional.group_norm(x.double(), 32, gn.weight.double(), gn.bias.double(), gn.eps)
            want = want.relu() if relu else want
            assert got.is_contiguous(memory_format=torch.channels_last)
            assert rel_err(got, want) < 2e-6
    # other group widths (4, 16 and 32 channels per group: every lane grouping of the wide kernel) and a map too
    # large for it (one-group-per-CTA kernel)
    for groups, ch, hw in [(32, 128, (25, 42)), (8, 128, (13, 21)), (2, 64, (7, 11)), (32, 256, (50, 84))]:
        gn2 = torch.nn.GroupNorm(groups, ch).cuda()
        with torch.no_grad():
            gn2.weight.copy_(1 + 0.3 * torch.randn(ch, generator=g))
            gn2.bias.copy_(0.2 * torch.randn(ch, generator=g))
        x = (torch.randn(2, ch, *hw, generator=g) * 2 - 0.5).cuda().contiguous(memory_format=torch.channels_last)
        got = ops.groupnorm_relu_nhwc(x, gn2, relu=True)
        want = torch.nn.functional.group_norm(x.double(), groups, gn2.weight.double(), gn2.bias.double(), gn2.eps).relu()
        assert rel_err(got, want) < 2e-6, (groups, ch, hw)


def test_channels_last_sources_need_no_transpose():
    """prepare_input / nchw_to_tiled from a channels_last activation give the same buffers as from NCHW."""
    from kgdet_b200 import ops
    x = torch.randn(2, 128, 9, 11, generator=torch.Generator().manual_seed(8)).cuda()
    xcl = x.contiguous(memory_format=torch.channels_last)
    for prec in ('bf16', 'tf32'):
        a = ops.prepare_input(x, 128, 3, 1, 1, 1, precision=prec)
        b = ops.prepare_input(xcl, 128, 3, 1, 1, 1, precision=prec)
        assert torch.equal(a.buf, b.buf)
    for split in (False, True):
        assert torch.equal(ops.nchw_to_tiled(x, relu=True, split=split).to_dense(),
                           ops.nchw_to_tiled(xcl, relu=True, split=split).to_dense())


def test_grouped_dcn_launch_equals_single_launches():
    """kgdet_dcn_forward_prepared_group (one persistent launch, dynamic tile scheduler) == the same deformable
    convolutions launched one by one, bit for bit: six jobs of a Kp3RepBlock stage shape (two inputs, three point
    sets, tiled split outputs into channel slices) and a mixed NCHW / tiled group with a ragged last tile."""
    from kgdet_b200 import ops
    g = torch.Generator().manual_seed(17)
    N, C, H, W, F = 3, 256, 13, 21, 256
    xa = torch.randn(N, C, H, W, generator=g).cuda()
    xb = torch.randn(N, C, H, W, generator=g).cuda()
    pts = (torch.randn(N, 166, H, W, generator=g) * 2).cuda()
    ws = {(br, k): (torch.randn(F, C, k, k, generator=g) * (1.0 / (C * k * k) ** 0.5)).cuda()
          for br in 'ab' for k in (3, 5, 7)}
    ops.set_precision('bf16')
    try:
        pa, pb = ops.prepare_input(xa, F), ops.prepare_input(xb, F)
        plans, lo = {}, 0
        for k in (3, 5, 7):
            plans[k] = ops.prepare_plan_points(pts, lo, (N, C, H, W), F, k, 1, k // 2, 1, gradient_mul=0.1)
            lo += 2 * k * k

        def run(grouped):
            rows = {br: ops.TiledRows(N * H * W, 3 * F, True, 'cuda') for br in 'ab'}
            for r in rows.values():
                r.buf.zero_()
            jobs = []
            for i, k in enumerate((3, 5, 7)):
                jobs.append((pa, plans[k], ws[('a', k)], rows['a'], i * F, True))
                jobs.append((pb, plans[k], ws[('b', k)], rows['b'], i * F, True))
            jobs.reverse()
            if grouped:
                ops.deform_conv_prepared_group(jobs)
            else:
                for j in jobs:
                    ops.deform_conv_prepared(*j)
            return rows
        single, grouped = run(False), run(True)
        for br in 'ab':
            assert torch.equal(single[br].buf, grouped[br].buf)
        assert float(grouped['a'].to_dense().abs().sum()) > 0
        # NCHW fp32 outputs in the same group as a tiled one (M = 3 * 13 * 21 = 819: ragged last tile)
        out_s = torch.full((N, 2 * F, H, W), -3.0, device='cuda')
        out_g = out_s.clone()
        rs, rg = ops.TiledRows(N * H * W, F, False, 'cuda'), ops.TiledRows(N * H * W, F, False, 'cuda')
        rs.buf.zero_(); rg.buf.zero_()
        os.environ['KGDET_UMMA_SPLITS'] = '1'      # single NCHW launches would split this small map over k-blocks
        try:
            for out, rows, fn in ((out_s, rs, None), (out_g, rg, ops.deform_conv_prepared_group)):
                jobs = [(pa, plans[5], ws[('a', 5)], out, 0, False), (pb, plans[3], ws[('b', 3)], out, F, True),
                        (pb, plans[7], ws[('b', 7)], rows, 0, True)]
                if fn is None:
                    for j in jobs:
                        ops.deform_conv_prepared(*j)
                else:
                    fn(jobs)
        finally:
            del os.environ['KGDET_UMMA_SPLITS']
        assert torch.equal(out_s, out_g) and torch.equal(rs.buf, rg.buf)
        assert (out_g != -3.0).all()
    finally:
        ops.set_precision(None)


def test_grouped_dcn_256_row_tiles_equal_128_row_tiles_full_size(monkeypatch):
    """The KGDet stage at its full size (batch 16, 25x42, six deformable convolutions, 792 tiles): the grouped kernel's
    256-row tiles (two halves sharing every weight slab, two TMEM accumulators; the default for big groups) give bit
    for bit what its 128-row tiles and the single launches give.  M = 16 800 = 65.6 pairs: ragged last pair."""
    from kgdet_b200 import ops
    g = torch.Generator().manual_seed(23)
    N, C, H, W, F = 16, 256, 25, 42, 256
    xa = torch.randn(N, C, H, W, generator=g).cuda()
    xb = torch.randn(N, C, H, W, generator=g).cuda()
    pts = (torch.randn(N, 166, H, W, generator=g) * 2).cuda()
    ws = {(br, k): (torch.randn(F, C, k, k, generator=g) * (1.0 / (C * k * k) ** 0.5)).cuda()
          for br in 'ab' for k in (3, 5, 7)}
    ops.set_precision('bf16')
    try:
        pa, pb = ops.prepare_input(xa, F), ops.prepare_input(xb, F)
        plans, lo = {}, 0
        for k in (3, 5, 7):
            plans[k] = ops.prepare_plan_points(pts, lo, (N, C, H, W), F, k, 1, k // 2, 1, gradient_mul=0.1)
            lo += 2 * k * k

        def run(mode):
            rows = {br: ops.TiledRows(N * H * W, 3 * F, True, 'cuda') for br in 'ab'}
            for r in rows.values():
                r.buf.zero_()
            jobs = []
            for i, k in enumerate((3, 5, 7)):
                jobs.append((pa, plans[k], ws[('a', k)], rows['a'], i * F, True))
                jobs.append((pb, plans[k], ws[('b', k)], rows['b'], i * F, True))
            jobs.reverse()
            if mode == 'single':
                for j in jobs:
                    ops.deform_conv_prepared(*j)
            else:
                monkeypatch.setenv('KGDET_GROUP_ROWS', mode)
                ops.deform_conv_prepared_group(jobs)
            torch.cuda.synchronize()
            return rows
        single, g128, g256 = run('single'), run('128'), run('256')
        for br in 'ab':
            assert torch.equal(single[br].buf, g128[br].buf)
            assert torch.equal(single[br].buf, g256[br].buf)
        # an odd number of 128-row tiles (N = 15: 15 750 positions = 123.05 tiles -> 124, 62 pairs exactly; N = 13: 107 tiles)
        for n_img in (13,):
            xs = xa[:n_img].contiguous()
            pin = ops.prepare_input(xs, F)
            pl = {k: ops.prepare_plan_points(pts[:n_img].contiguous(), 0 if k == 3 else (18 if k == 5 else 68),
                                             (n_img, C, H, W), F, k, 1, k // 2, 1) for k in (3, 5, 7)}
            outs = []
            for mode in ('128', '256'):
                monkeypatch.setenv('KGDET_GROUP_ROWS', mode)
                rows = ops.TiledRows(n_img * H * W, 6 * F, True, 'cuda')
                rows.buf.zero_()
                jobs = [(pin, pl[k], ws[(br, k)], rows, (i * 2 + j) * F, True)
                        for i, k in enumerate((7, 5, 3)) for j, br in enumerate('ab')]
                ops.deform_conv_prepared_group(jobs)
                torch.cuda.synchronize()
                outs.append(rows.buf.clone())
            assert torch.equal(outs[0], outs[1])
    finally:
        ops.set_precision(None)


@pytest.mark.parametrize('shape,groups,relu', [((2, 256, 25, 42), 32, True), ((3, 64, 13, 21), 16, True),
                                               ((1, 128, 7, 11), 4, False), ((2, 96, 9, 9), 12, True)])
def test_groupnorm_relu_nhwc_autograd_matches_torch(shape, groups, relu):
    """kgdet_groupnorm_relu_nhwc + kgdet_groupnorm_relu_nhwc_backward behind an autograd Function: output and the
    gradients of the input, gamma and beta against torch's GroupNorm (+ ReLU) evaluated in fp64."""
    from kgdet_b200.ops.pointwise import groupnorm_relu_nhwc_autograd
    n, c, h, w = shape
    g = torch.Generator().manual_seed(c + h)
    gn = torch.nn.GroupNorm(groups, c).cuda()
    with torch.no_grad():
        gn.weight.copy_(1 + 0.3 * torch.randn(c, generator=g))
        gn.bias.copy_(0.3 * torch.randn(c, generator=g))
    x = (torch.randn(n, c, h, w, generator=g) * 2 + 0.5).cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    up = torch.randn(n, c, h, w, generator=g).cuda()
    y = groupnorm_relu_nhwc_autograd(x, gn, relu)
    assert y.is_contiguous(memory_format=torch.channels_last)
    y.backward(up)
    x64 = x.detach().double().requires_grad_()
    w64, b64 = gn.weight.detach().double().requires_grad_(), gn.bias.detach().double().requires_grad_()
    y64 = torch.nn.functional.group_norm(x64, groups, w64, b64, gn.eps)
    y64 = torch.relu(y64) if relu else y64
    y64.backward(up.double())

    def close(a, b, tol):
        return (a.double() - b).abs().max().item() <= tol * b.abs().max().item() + 1e-12
    assert close(y, y64, 2e-6)
    assert close(x.grad, x64.grad, 2e-5)
    assert close(gn.weight.grad, w64.grad, 2e-5) and close(gn.bias.grad, b64.grad, 2e-5)
