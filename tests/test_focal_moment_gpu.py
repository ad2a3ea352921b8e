"""GPU parity tests of the sigmoid focal loss and the moment transform."""
import numpy as np
import pytest
import torch

from oracle import build_ref, focal_oracle, moment_oracle
from tests._data import rel_err

pytestmark = pytest.mark.gpu


def _focal_inputs(M, C, seed=0, with_ignore=True):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(M, C, generator=g) * 3
    targets = torch.randint(0, C + 1, (M,), generator=g)
    if with_ignore and M > 4:
        targets[::7] = -1          # ignored rows (sigmoid_focal_loss_cuda.cu:36-37: both terms vanish)
    return logits, targets


@pytest.mark.parametrize('M,C', [(1, 13), (2100, 13), (33600, 13), (257, 80)])
def test_focal_forward_backward_vs_oracle(M, C):
    from kgdet_b200.ops import sigmoid_focal_loss
    logits, targets = _focal_inputs(M, C, seed=M)
    d_loss = torch.rand(M, C, generator=torch.Generator().manual_seed(1))
    x = logits.cuda().requires_grad_()
    loss = sigmoid_focal_loss(x, targets.cuda(), 2.0, 0.25)
    loss.backward(d_loss.cuda())
    ref = focal_oracle.sigmoid_focal_loss_forward(logits.numpy(), targets.numpy(), 2.0, 0.25)
    refb = focal_oracle.sigmoid_focal_loss_backward(logits.numpy(), targets.numpy(), d_loss.numpy(), 2.0, 0.25)
    assert rel_err(loss, torch.from_numpy(ref)) < 1e-5          # tolerance: rel 1e-5 (fp32)
    assert rel_err(x.grad, torch.from_numpy(refb)) < 1e-5


def test_focal_matches_reference_cuda_kernel():
    ref = build_ref.load('sigmoid_focal_loss_cuda')
    if ref is None:
        pytest.skip('oracle/_ref/sigmoid_focal_loss_cuda.so not built')
    from kgdet_b200.ops import sigmoid_focal_loss
    logits, targets = _focal_inputs(2100, 13, seed=5)
    x, t = logits.cuda(), targets.cuda()
    want = ref.forward(x, t, 13, 2.0, 0.25)
    d = torch.rand_like(x)
    wantb = ref.backward(x, t, d, 13, 2.0, 0.25)
    xg = x.clone().requires_grad_()
    got = sigmoid_focal_loss(xg, t, 2.0, 0.25)
    got.backward(d)
    # same promotion pattern, same device math library -> expected identical; allow 1 ulp-ish
    assert rel_err(got, want) < 1e-6
    assert rel_err(xg.grad, wantb) < 1e-6


def test_focal_sum_fused_matches_reduction():
    from kgdet_b200.ops import sigmoid_focal_loss_sum
    logits, targets = _focal_inputs(2100, 13, seed=9)
    w = torch.rand(2100, generator=torch.Generator().manual_seed(2))
    x = logits.cuda().requires_grad_()
    s = sigmoid_focal_loss_sum(x, targets.cuda(), w.cuda(), 2.0, 0.25)
    (s / 37.0).backward()
    want = focal_oracle.focal_loss_reduced(logits.numpy(), targets.numpy(), w.numpy(), reduction='mean',
                                           avg_factor=1.0)
    assert abs(s.item() - want) / abs(want) < 1e-5
    refb = focal_oracle.sigmoid_focal_loss_backward(logits.numpy(), targets.numpy(),
                                                    (w.numpy()[:, None] / 37.0) * np.ones((1, 13), np.float32))
    assert rel_err(x.grad, torch.from_numpy(refb)) < 1e-5


def test_focal_module_and_errors():
    from kgdet_b200.ops import SigmoidFocalLoss, sigmoid_focal_loss
    logits, targets = _focal_inputs(64, 13, seed=3, with_ignore=False)
    m = SigmoidFocalLoss(2.0, 0.25)
    v = m(logits.cuda(), targets.cuda())
    want = focal_oracle.sigmoid_focal_loss_forward(logits.numpy(), targets.numpy()).astype(np.float64).sum()
    assert abs(v.item() - want) / want < 1e-5
    assert repr(m) == 'SigmoidFocalLoss(gamma=2.0, alpha=0.25)'
    with pytest.raises(NotImplementedError):
        sigmoid_focal_loss(logits, targets, 2.0, 0.25)
    e = sigmoid_focal_loss(torch.zeros(0, 13).cuda(), torch.zeros(0, dtype=torch.long).cuda(), 2.0, 0.25)
    assert e.shape == (0, 13)


@pytest.mark.parametrize('shape,y_first', [((2, 166, 25, 42), True), ((16, 18, 100, 168), True),
                                           ((1000, 18), False), ((3, 166, 7, 11), True)])
def test_moment_forward_backward_vs_oracle(shape, y_first):
    from kgdet_b200.ops import points2bbox_moment
    g = torch.Generator().manual_seed(shape[0])
    pts = torch.randn(*shape, generator=g) * 3
    mt = torch.tensor([0.3, -0.2])
    gb = torch.randn(shape[0], 4, *shape[2:], generator=g)
    p64 = pts.double().requires_grad_()
    mt64 = mt.double().requires_grad_()
    ref = moment_oracle.points2bbox_moment(p64, mt64, 0.01, y_first)
    ref.backward(gb.double())
    p = pts.cuda().requires_grad_()
    m = mt.cuda().requires_grad_()
    out = points2bbox_moment(p, m, 0.01, y_first)
    out.backward(gb.cuda())
    assert out.shape == ref.shape
    assert rel_err(out, ref) < 1e-5
    assert rel_err(p.grad, p64.grad) < 1e-5
    assert rel_err(m.grad, mt64.grad) < 1e-4      # fp32 atomics over N*S positions
