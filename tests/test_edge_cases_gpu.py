"""Edge cases of the hot-path operators on the GPU (empty and degenerate inputs, maps smaller than a tile), each
against the oracle / the reference's documented behaviour.  Call path: Python mirror -> ctypes -> C ABI."""
import numpy as np
import pytest
import torch

from oracle import dcn_oracle, focal_oracle, moment_oracle, nms_oracle
from tests._data import dcn_case, random_boxes, rel_err

pytestmark = pytest.mark.gpu


def test_focal_loss_with_no_rows_and_single_row():
    """sigmoid_focal_loss on [0, 13] logits returns an empty [0, 13] loss and gradient (the reference kernel is a
    grid-stride loop over 0 elements, sigmoid_focal_loss_cuda.cu:24-58); one row is checked against the oracle."""
    from kgdet_b200.ops import sigmoid_focal_loss
    x = torch.zeros(0, 13, device='cuda', requires_grad=True)
    loss = sigmoid_focal_loss(x, torch.zeros(0, dtype=torch.long, device='cuda'), 2.0, 0.25)
    assert loss.shape == (0, 13)
    loss.sum().backward()
    assert x.grad.shape == (0, 13)
    g = torch.Generator().manual_seed(1)
    x1 = torch.randn(1, 13, generator=g)
    t1 = torch.tensor([13])                                              # the last class
    ours = sigmoid_focal_loss(x1.cuda(), t1.cuda(), 2.0, 0.25).cpu()
    want = torch.as_tensor(focal_oracle.sigmoid_focal_loss_forward(x1.numpy(), t1.numpy(), 2.0, 0.25))
    assert rel_err(ours, want) < 1e-6
    t0 = torch.tensor([0])                                               # background: every class is a negative
    ours0 = sigmoid_focal_loss(x1.cuda(), t0.cuda(), 2.0, 0.25).cpu()
    want0 = torch.as_tensor(focal_oracle.sigmoid_focal_loss_forward(x1.numpy(), t0.numpy(), 2.0, 0.25))
    assert rel_err(ours0, want0) < 1e-6


def test_moment_transform_degenerate_point_sets():
    """P = 1: torch.std of one value is NaN (n - 1 = 0) and so is the half extent, the centre stays finite
    (KP3:374-388); identical points: std 0, box collapses onto the mean, gradient finite; empty batch."""
    from kgdet_b200.ops.moment import points2bbox_moment
    mt = torch.tensor([0.3, -0.2])
    one = torch.tensor([[1.5, -2.0]]).view(1, 2, 1, 1)
    ours = points2bbox_moment(one.cuda(), mt.cuda()).cpu()
    want = moment_oracle.points2bbox_moment(one, mt)
    assert torch.isnan(want).all() and torch.isnan(ours).all()
    same = torch.tensor([2.0, 3.0]).repeat(9).view(1, 18, 1, 1).repeat(2, 1, 3, 2).contiguous()
    p = same.cuda().requires_grad_()
    box = points2bbox_moment(p, mt.cuda())
    want = moment_oracle.points2bbox_moment(same, mt)
    assert torch.equal(box.detach().cpu(), want)
    box.sum().backward()
    assert torch.isfinite(p.grad).all()
    empty = points2bbox_moment(torch.zeros(0, 18, 4, 5, device='cuda'), mt.cuda())
    assert empty.shape == (0, 4, 4, 5)


def test_deform_conv_rejects_maps_smaller_than_the_kernel_like_the_reference():
    """deform_conv_cuda.cpp:126-127: "input image is smaller than kernel" whatever the padding."""
    from kgdet_b200 import ops
    d = dcn_case(N=1, C=8, H=1, W=1, Cout=4, k=3)
    with pytest.raises(RuntimeError, match='smaller than kernel'):
        ops.deform_conv(d['x'].cuda(), d['offset'].cuda(), d['weight'].cuda(), 1, 1)


@pytest.mark.parametrize('case,precision,tol', [
    (dict(N=1, C=8, H=3, W=3, Cout=4, k=3, offset_std=4.0), 'fp32', 1e-5),      # the smallest legal map, most samples outside
    (dict(N=1, C=64, H=3, W=3, Cout=64, k=3), 'bf16', 1e-2),                    # 9 rows of a 128-row tensor-core tile
    (dict(N=1, C=64, H=3, W=4, Cout=64, k=3), 'tf32x3', 1e-5),
    (dict(N=2, C=64, H=5, W=5, Cout=128, k=5, offset_std=0.0), 'bf16', 1e-2),   # zero offsets: a plain 5x5 convolution
])
def test_deform_conv_on_maps_smaller_than_a_tile(case, precision, tol):
    from kgdet_b200 import ops
    d = dcn_case(**case)
    c = lambda t: None if t is None else t.double()
    ref = dcn_oracle.deform_conv_forward(c(d['x']), c(d['offset']), c(d['weight']), d['stride'], d['padding'],
                                         d['dilation'], d['groups'], d['deformable_groups'])
    bw = dcn_oracle.deform_conv_backward(c(d['x']), c(d['offset']), c(d['weight']), c(d['grad_out']), d['stride'],
                                         d['padding'], d['dilation'], d['groups'], d['deformable_groups'])
    ops.set_precision(precision)
    try:
        x, off, w = (d[k].cuda().requires_grad_() for k in ('x', 'offset', 'weight'))
        out = ops.deform_conv(x, off, w, d['stride'], d['padding'], d['dilation'])
        out.backward(d['grad_out'].cuda())
        torch.cuda.synchronize()
    finally:
        ops.set_precision(None)
    assert rel_err(out.detach().cpu(), ref) < tol
    btol = tol if precision != 'tf32x3' else 1e-5
    for got, key in ((x.grad, 'grad_input'), (off.grad, 'grad_offset'), (w.grad, 'grad_weight')):
        if float(bw[key].abs().max()) > 0:
            assert rel_err(got.cpu(), bw[key]) < btol, key
        else:
            assert float(got.abs().max()) == 0, key


def test_nms_single_box_duplicates_and_empty_segments():
    """n = 1 keeps the box; identical boxes keep only the best one under both comparators; the batched launch with a
    segment below the score threshold, a one-box segment and a segment of duplicates returns the oracle's flags."""
    from kgdet_b200.ops import nms
    from kgdet_b200.ops.nms import nms_wrapper
    one = torch.tensor([[3., 4., 50., 60., 0.7]])
    kept, inds = nms(one.cuda(), 0.5)
    assert inds.tolist() == [0] and torch.equal(kept.cpu(), one)
    dup = torch.tensor([[10., 10., 99., 80., s] for s in (0.3, 0.9, 0.5, 0.7)])
    assert nms(dup.cuda(), 0.5)[1].tolist() == [1]
    assert list(nms_oracle.nms_keep(dup, 0.5, 0)) == [1] and list(nms_oracle.nms_keep(dup, 0.5, 1)) == [1]
    # dense batched mode: 4 segments of 8 rows
    rb = random_boxes(8, seed=5)
    low = rb.clone(); low[:, 4] = 0.01                                   # everything under score_thr
    single = rb.clone(); single[:, 4] = 0.0; single[3, 4] = 0.8          # one row present
    dups = dup.repeat(2, 1).clone(); dups[:, 4] = torch.linspace(0.2, 0.9, 8)
    segs = [low, single, dups, rb]
    flags = nms_wrapper.batched_nms_flags(torch.cat(segs).cuda(), None, 8, 0.5, score_thr=0.05).cpu().view(4, 8)
    for f, s in zip(flags, segs):
        present = (s[:, 4] > 0.05).nonzero().flatten()
        want = np.zeros(8, dtype=np.uint8)
        if present.numel():
            keep = nms_oracle.nms_keep(s[present], 0.5, 0)
            want[present[torch.as_tensor(keep, dtype=torch.long)].numpy()] = 1
        assert np.array_equal(f.numpy(), want)
