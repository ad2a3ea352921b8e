"""Size-independent properties of the hot-path operators at the FULL sizes of BASELINE.json (where the CPU oracle is
too slow to be the checker): linearity, batch-permutation equivariance, translation / scale equivariance, greedy-NMS
invariants (idempotence, sorted output, no surviving overlap, maximality), a checksum of checksums for the loss.
Call path: Python mirror -> ctypes -> C ABI."""
import pytest
import torch

from tests._data import dcn_case, random_boxes, rel_err

pytestmark = pytest.mark.gpu


def _iou_matrix(a, b):
    """IoU with the reference's +1 widths (nms_cpu.cpp:18-52), float64."""
    a, b = a.double(), b.double()
    area_a = (a[:, 2] - a[:, 0] + 1) * (a[:, 3] - a[:, 1] + 1)
    area_b = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    w = (torch.min(a[:, None, 2], b[None, :, 2]) - torch.max(a[:, None, 0], b[None, :, 0]) + 1).clamp(min=0)
    h = (torch.min(a[:, None, 3], b[None, :, 3]) - torch.max(a[:, None, 1], b[None, :, 1]) + 1).clamp(min=0)
    inter = w * h
    return inter / (area_a[:, None] + area_b[None, :] - inter)


@pytest.mark.parametrize('k', [3, 7])
@pytest.mark.parametrize('precision,tol', [('tf32x3', 2e-5), ('bf16', 1.5e-2)])
def test_deform_conv_full_size_linearity_and_batch_equivariance(k, precision, tol):
    """KGDet call [16, 256, 25, 42] (K = 9 / 49).  For fixed offsets the operator is linear in the input and in the
    weight; permuting the images of the batch (input and offsets together) permutes the output exactly (bitwise: tiles
    do not mix images' arithmetic)."""
    from kgdet_b200 import ops
    d = dcn_case(N=16, C=256, H=25, W=42, Cout=256, k=k, seed=40 + k)
    g = torch.Generator().manual_seed(7)
    x1, off, w1 = d['x'].cuda(), d['offset'].cuda(), d['weight'].cuda()
    x2 = torch.randn(x1.shape, generator=g).cuda()
    w2 = (torch.randn(w1.shape, generator=g) * float(w1.std())).cuda()
    ops.set_precision(precision)
    try:
        f = lambda x, w: ops.deform_conv(x, off, w, 1, k // 2)
        y11, y21, y12 = f(x1, w1), f(x2, w1), f(x1, w2)
        assert rel_err(f(x1 + x2, w1), y11 + y21) < tol
        assert rel_err(f(x1, w1 + w2), y11 + y12) < tol
        assert rel_err(f(2.0 * x1, w1), 2.0 * y11) < 1e-6          # powers of two commute with every rounding
        perm = torch.randperm(16, generator=g).cuda()
        yp = ops.deform_conv(x1[perm].contiguous(), off[perm].contiguous(), w1, 1, k // 2)
        assert torch.equal(yp, y11[perm])
    finally:
        ops.set_precision(None)


def test_deform_conv_full_size_backward_is_the_adjoint_of_forward():
    """<grad_out, J dx> == <J^T grad_out, dx> for the input path and the weight path of the bf16 tensor-core kernels
    at the KGDet size (K = 25): forward, bulk-reduction col2im and fused weight gradient describe the same linear
    map (rel 1e-2, bf16 operands)."""
    from kgdet_b200 import ops
    k = 5
    d = dcn_case(N=16, C=256, H=25, W=42, Cout=256, k=k, seed=3)
    g = torch.Generator().manual_seed(9)
    off = d['offset'].cuda()
    x, w, go = d['x'].cuda().requires_grad_(), d['weight'].cuda().requires_grad_(), d['grad_out'].cuda()
    dx = torch.randn(x.shape, generator=g).cuda()
    dw = (torch.randn(w.shape, generator=g) * float(w.detach().std())).cuda()
    ops.set_precision('bf16')
    try:
        ops.deform_conv(x, off, w, 1, k // 2).backward(go)
        with torch.no_grad():
            jx = ops.deform_conv(dx, off, w.detach(), 1, k // 2)
            jw = ops.deform_conv(x.detach(), off, dw, 1, k // 2)
    finally:
        ops.set_precision(None)
    lhs_x, rhs_x = float((go.double() * jx.double()).sum()), float((x.grad.double() * dx.double()).sum())
    lhs_w, rhs_w = float((go.double() * jw.double()).sum()), float((w.grad.double() * dw.double()).sum())
    scale_x = float(go.double().norm() * jx.double().norm())
    scale_w = float(go.double().norm() * jw.double().norm())
    assert abs(lhs_x - rhs_x) < 1e-2 * scale_x, (lhs_x, rhs_x)
    assert abs(lhs_w - rhs_w) < 1e-2 * scale_w, (lhs_w, rhs_w)


@pytest.mark.parametrize('n,clustered', [(3350, True), (20011, True), (20011, False)])
@pytest.mark.parametrize('cmp_mode', [0, 1])
def test_nms_invariants_at_size(n, clustered, cmp_mode):
    """Greedy NMS: kept indices ascending (nms_wrapper.py:44-46), no two survivors overlap above the threshold, every
    suppressed box overlaps a better-scored survivor (maximality), and the operator is idempotent."""
    from kgdet_b200.ops.nms import nms_wrapper
    dets = random_boxes(n, seed=n + cmp_mode, clustered=clustered).cuda()
    keep = nms_wrapper._nms_keep_cuda(dets, 0.5, cmp_mode)
    assert keep.dtype == torch.long and bool((keep[1:] > keep[:-1]).all())
    kept = dets[keep]
    over = (lambda m: m >= 0.5) if cmp_mode == 1 else (lambda m: m > 0.5)
    iou_kk = _iou_matrix(kept[:, :4], kept[:, :4])
    iou_kk.fill_diagonal_(0)
    assert not bool(over(iou_kk).any())
    gone = torch.ones(n, dtype=torch.bool, device='cuda')
    gone[keep] = False
    sup = dets[gone]
    for lo in range(0, sup.shape[0], 4096):                      # chunks: the IoU matrix is |suppressed| x |kept|
        blk = sup[lo:lo + 4096]
        better = kept[None, :, 4] > blk[:, None, 4]
        assert bool((over(_iou_matrix(blk[:, :4], kept[:, :4])) & better).any(dim=1).all())
    again = nms_wrapper._nms_keep_cuda(kept.contiguous(), 0.5, cmp_mode)
    assert again.numel() == keep.numel()


def test_moment_transform_equivariance_full_size():
    """points2bbox('moment') on [16, 166, 25, 42] (P = 83): translating all points translates the box, scaling them by
    s > 0 scales it (mean and unbiased std are equivariant); rel 1e-5."""
    from kgdet_b200.ops.moment import points2bbox_moment
    g = torch.Generator().manual_seed(2)
    pts = (torch.randn(16, 166, 25, 42, generator=g) * 8).cuda()
    mt = torch.tensor([0.2, -0.1]).cuda()
    box = points2bbox_moment(pts, mt)
    ty, tx = 3.25, -7.5
    shift = torch.tensor([ty, tx]).repeat(83).view(1, 166, 1, 1).cuda()           # y-first interleave (KP3:353-356)
    moved = points2bbox_moment(pts + shift, mt)
    want = box + torch.tensor([tx, ty, tx, ty]).view(1, 4, 1, 1).cuda()
    assert rel_err(moved, want) < 1e-5
    assert rel_err(points2bbox_moment(pts * 4.0, mt), box * 4.0) < 1e-6          # power of two: exact up to the exp
    assert bool((box[:, 2] >= box[:, 0]).all() and (box[:, 3] >= box[:, 1]).all())


def test_focal_loss_checksum_of_checksums_at_five_level_size():
    """[8 * 22 400, 13] logits (five FPN levels, batch 8): the fused weighted sum equals the sum of the elementwise
    kernel's output times the row weights (fp64 checksum), per-level partial sums add up to the total, losses are
    non-negative, and the gradient pushes the target class up and every other class down."""
    from kgdet_b200.ops import sigmoid_focal_loss, sigmoid_focal_loss_sum
    g = torch.Generator().manual_seed(4)
    M = 8 * 22400
    x = (torch.randn(M, 13, generator=g) * 3).cuda().requires_grad_()
    t = torch.randint(0, 14, (M,), generator=g).cuda()
    wgt = torch.rand(M, generator=g).cuda()
    loss = sigmoid_focal_loss(x, t, 2.0, 0.25)
    assert bool((loss >= 0).all())
    want = float((loss.detach().double() * wgt.double()[:, None]).sum())
    fused = float(sigmoid_focal_loss_sum(x.detach(), t, wgt, 2.0, 0.25))
    assert abs(fused - want) < 1e-5 * abs(want)
    levels = [8 * n for n in (16800, 4200, 1050, 273, 77)]
    parts, lo = 0.0, 0
    for m in levels:
        parts += float(sigmoid_focal_loss_sum(x.detach()[lo:lo + m], t[lo:lo + m], wgt[lo:lo + m], 2.0, 0.25))
        lo += m
    assert lo == M and abs(parts - want) < 1e-5 * abs(want)
    loss.sum().backward()
    onehot = torch.zeros(M, 14, device='cuda', dtype=torch.bool)
    onehot[torch.arange(M, device='cuda'), t] = True
    pos = onehot[:, 1:]                                            # label c > 0 is class c - 1 (sigmoid_focal_loss_cuda.cu:34-37)
    assert bool((x.grad[pos] <= 0).all()) and bool((x.grad[~pos] >= 0).all())
