"""CPU tests: the oracles against their pins (golden fixtures generated from the reference's own
code, torchvision, finite differences, the reference nms_cpu.cpp in oracle/_ref)."""
import os

import numpy as np
import pytest
import torch

from oracle import build_ref, dcn_oracle, focal_oracle, moment_oracle, nms_oracle
from tests._data import dcn_case, random_boxes, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def gold(name):
    return np.load(os.path.join(GOLD, name))


# ---- NMS ------------------------------------------------------------------------------------
@pytest.mark.parametrize('case', list('abcde'))
def test_nms_oracle_matches_golden_reference_keep(case):
    g = gold('nms.npz')
    keep = nms_oracle.nms_keep(g['dets_' + case], 0.5, cmp_mode=1)
    assert np.array_equal(keep, g['keep_' + case])


def test_nms_oracle_matches_reference_cpu_module():
    ref = build_ref.load('nms_cpu')
    if ref is None:
        pytest.skip('oracle/_ref/nms_cpu.so not built here')
    for seed in range(4):
        dets = random_boxes(1500, seed=100 + seed, clustered=bool(seed % 2))
        for thr in (0.3, 0.5, 0.7):
            assert np.array_equal(nms_oracle.nms_keep(dets, thr, 1), ref.nms(dets, thr).numpy())


def test_nms_oracle_edge_cases():
    assert nms_oracle.nms_keep(np.zeros((0, 5), np.float32), 0.5).size == 0
    one = np.array([[0, 0, 10, 10, 0.5]], np.float32)
    assert list(nms_oracle.nms_keep(one, 0.5)) == [0]
    dets = np.array([[0, 0, 9, 9, 0.9], [0, 0, 9, 4, 0.8]], np.float32)     # IoU exactly 0.5
    assert list(nms_oracle.nms_keep(dets, 0.5, 0)) == [0, 1]
    assert list(nms_oracle.nms_keep(dets, 0.5, 1)) == [0]
    kept, inds = nms_oracle.nms(torch.from_numpy(dets), 0.5)
    assert kept.shape == (1, 5) and inds.dtype == torch.long


# ---- focal loss --------------------------------------------------------------------------------
def test_focal_oracle_matches_reference_debug_twin():
    g = gold('focal_loss.npz')
    loss = focal_oracle.sigmoid_focal_loss_forward(g['logits'], g['targets'], 2.0, 0.25)
    assert rel_err(torch.from_numpy(loss), torch.from_numpy(g['loss'])) < 1e-5
    red = focal_oracle.focal_loss_reduced(g['logits'], g['targets'], g['weight'], 2.0, 0.25, 'mean', 17.0)
    assert abs(red - float(g['reduced'])) / abs(float(g['reduced'])) < 1e-5
    d = np.repeat(g['weight'][:, None] / 17.0, 13, axis=1).astype(np.float32)
    grad = focal_oracle.sigmoid_focal_loss_backward(g['logits'], g['targets'], d, 2.0, 0.25)
    assert rel_err(torch.from_numpy(grad), torch.from_numpy(g['grad'])) < 1e-5


def test_focal_oracle_ignored_rows_and_background():
    x = np.random.RandomState(0).randn(5, 4).astype(np.float32)
    t = np.array([-1, 0, 1, 4, 2])
    loss = focal_oracle.sigmoid_focal_loss_forward(x, t)
    assert np.all(loss[0] == 0)                      # t < 0: both indicator terms vanish (:36-37)
    assert np.all(loss[1] > 0)                       # background row: every column is a negative


# ---- moment transform ---------------------------------------------------------------------------
def test_moment_oracle_matches_reference_points2bbox():
    g = gold('moment.npz')
    mt = torch.from_numpy(g['mt'])
    a = moment_oracle.points2bbox_moment(torch.from_numpy(g['pts83']), mt, 0.01, True)
    b = moment_oracle.points2bbox_moment(torch.from_numpy(g['pts9']), mt, 0.01, False)
    assert rel_err(a, torch.from_numpy(g['bbox83'])) < 1e-6
    assert rel_err(b, torch.from_numpy(g['bbox9'])) < 1e-6


# ---- deformable convolution -----------------------------------------------------------------------
@pytest.mark.parametrize('case', [
    dict(N=2, C=8, H=7, W=9, Cout=6, k=3),
    dict(N=2, C=8, H=7, W=9, Cout=8, k=3, groups=2, dg=2, mask=True),
    dict(N=1, C=4, H=9, W=8, Cout=4, k=5, stride=2),
    dict(N=2, C=4, H=6, W=7, Cout=4, k=3, dil=2, dg=2, mask=True),
    dict(N=1, C=4, H=3, W=3, Cout=4, k=3, offset_std=6.0),
])
def test_dcn_oracle_matches_torchvision(case):
    import torchvision
    d = dcn_case(**case)
    x = d['x'].double().requires_grad_()
    w = d['weight'].double().requires_grad_()
    off = d['offset'].double().requires_grad_()
    m = None if d['mask'] is None else d['mask'].double().requires_grad_()
    ref = torchvision.ops.deform_conv2d(x, off, w, None, stride=d['stride'], padding=d['padding'],
                                        dilation=d['dilation'], mask=m)
    ref.backward(d['grad_out'].double())
    out = dcn_oracle.deform_conv_forward(x.detach(), off.detach(), w.detach(), d['stride'], d['padding'],
                                         d['dilation'], d['groups'], d['deformable_groups'],
                                         mask=None if m is None else m.detach())
    bw = dcn_oracle.deform_conv_backward(x.detach(), off.detach(), w.detach(), d['grad_out'].double(),
                                         d['stride'], d['padding'], d['dilation'], d['groups'],
                                         d['deformable_groups'], mask=None if m is None else m.detach())
    assert rel_err(out, ref) < 1e-12
    assert rel_err(bw['grad_input'], x.grad) < 1e-12
    assert rel_err(bw['grad_offset'], off.grad) < 1e-12
    assert rel_err(bw['grad_weight'], w.grad) < 1e-12
    if m is not None:
        assert rel_err(bw['grad_mask'], m.grad) < 1e-12


def test_dcn_oracle_explicit_backward_equals_autograd():
    d = dcn_case(N=2, C=6, H=5, W=6, Cout=4, k=3, dg=3, mask=True, bias=True)
    x, off, w, m, b = (d[k].double().requires_grad_() for k in ('x', 'offset', 'weight', 'mask', 'bias'))
    out = dcn_oracle.deform_conv_forward(x, off, w, 1, 1, 1, 1, 3, mask=m, bias=b)
    out.backward(d['grad_out'].double())
    bw = dcn_oracle.deform_conv_backward(x.detach(), off.detach(), w.detach(), d['grad_out'].double(), 1, 1, 1,
                                         1, 3, mask=m.detach(), with_bias=True)
    for k, t in dict(grad_input=x, grad_offset=off, grad_weight=w, grad_mask=m, grad_bias=b).items():
        assert rel_err(bw[k], t.grad) < 1e-12, k


def test_dcn_oracle_zero_offset_is_plain_convolution():
    d = dcn_case(N=2, C=8, H=7, W=9, Cout=6, k=3)
    out = dcn_oracle.deform_conv_forward(d['x'].double(), torch.zeros_like(d['offset']).double(),
                                         d['weight'].double(), 1, 1)
    ref = torch.nn.functional.conv2d(d['x'].double(), d['weight'].double(), padding=1)
    assert rel_err(out, ref) < 1e-12


# ---- head mirror (host logic) against the reference head's golden outputs ------------------------
NAMES = ['cls_1', 'cls_2', 'cls_3', 'kpt_1', 'kpt_2', 'kpt_3', 'bbox_1', 'bbox_2', 'bbox_3']


def _cpu_head():
    from tests._cpu_head import make_cpu_head
    from tests.golden.gen_golden import fill_state_dict
    head = make_cpu_head()
    head.load_state_dict(fill_state_dict(head.state_dict()), strict=True)
    return head.eval()


def test_head_mirror_forward_matches_reference_head_golden():
    g = gold('head_p7.npz')
    head = _cpu_head()
    assert len(head.state_dict()) == 53 and sum(p.numel() for p in head.parameters()) == 27852247
    with torch.no_grad():
        out = head.forward_single(torch.from_numpy(g['x']))
    for n, o in zip(NAMES, out):
        assert rel_err(o, torch.from_numpy(g[n])) < 1e-5, n


def test_head_mirror_get_bboxes_matches_reference_golden():
    g = gold('get_bboxes.npz')
    head = _cpu_head()
    dets, labels, kpts = head.get_bboxes([torch.from_numpy(g['logit'])], [torch.from_numpy(g['kpt3'])],
                                         [torch.from_numpy(g['bbox3'])], [(800, 1333)] * 2, 0.05, 0.5, 1000, 100)
    for i in range(2):
        rd, rl, rk = g['dets_%d' % i], g['labels_%d' % i], g['kpts_%d' % i]
        nv = int((labels[i] >= 0).sum())
        assert nv == rd.shape[0]
        o = np.argsort(-rd[:, 4], kind='stable')
        assert np.allclose(dets[i, :nv].numpy(), rd[o], rtol=0, atol=1e-4)
        assert np.array_equal(labels[i, :nv].numpy(), rl[o])
        assert np.allclose(kpts[i, :nv].numpy(), rk[o], rtol=0, atol=1e-3)


@pytest.mark.parametrize('variant,n_params', [('parallel', 6806475), ('serial', 5638523)])
def test_reppoints_kp_mirror_matches_reference_head_golden(variant, n_params):
    """BASELINE.json configs[3]: the RepPoints-Kp parallel / serial head mirrors on the CPU oracles against
    the outputs of the reference's own heads (reppoints_head_kp_{parallel,serial}.py:292-341)."""
    from tests._cpu_head import make_cpu_reppoints_head
    from tests.golden.gen_golden import fill_state_dict
    g = gold('reppoints_%s.npz' % variant)
    head = make_cpu_reppoints_head(variant)
    assert sum(p.numel() for p in head.parameters()) == n_params
    head.load_state_dict(fill_state_dict(head.state_dict(), seed=4321), strict=True)
    head.eval()
    for li in range(3):
        with torch.no_grad():
            outs = head.forward_single(torch.from_numpy(g['x%d' % li]))
            bbox = head.points2bbox(outs[4])
        for n, o in zip(['cls', 'kpt_init', 'kpt_refine', 'rep_init', 'rep_refine'], outs):
            assert rel_err(o, torch.from_numpy(g['%s%d' % (n, li)])) < 1e-5, (n, li)
        assert rel_err(bbox, torch.from_numpy(g['bbox_refine%d' % li])) < 1e-5


def check_reppoints_bboxes(g, variant, dets, labels, kpts, tag=''):
    """(dets, labels, kpts) of RepPointsKpHead.get_bboxes against tests/golden/reppoints_bboxes.npz
    (tag 'rs_': the rescale=True case)."""
    variant = variant + ('_' + tag if tag else '')
    for i in range(dets.shape[0]):
        rd, rl = g['%s_dets_%d' % (variant, i)], g['%s_labels_%d' % (variant, i)]
        rk, rs = g['%s_kpts_head_%d' % (variant, i)], g['%s_kpts_rowsum_%d' % (variant, i)]
        nv = int((labels[i] >= 0).sum())
        assert nv == rd.shape[0]
        # the reference sorts the concatenated per-class survivors by score (bbox_nms_kp.py:64-70); equal scores
        # keep class order in both implementations
        assert np.all(np.diff(rd[:, 4]) <= 0)
        assert np.allclose(dets[i, :nv].cpu().numpy(), rd, rtol=0, atol=1e-4)
        assert np.array_equal(labels[i, :nv].cpu().numpy(), rl)
        kk = kpts[i, :nv].cpu()
        assert np.allclose(kk[:rk.shape[0]].numpy(), rk, rtol=0, atol=1e-3)
        assert np.allclose(kk.double().sum(1).numpy(), rs, rtol=0, atol=0.2)      # 882 values of up to ~400 per row


@pytest.mark.parametrize('variant', ['parallel', 'serial'])
def test_reppoints_kp_mirror_get_bboxes_matches_reference_golden(variant):
    """Multi-level get_bboxes of the RepPoints-Kp heads (PAR/SER:615-752 + multiclass_nms_kp), PyTorch restatement
    with the reference's nms_cpu, against the UNCHANGED reference class (gen_reppoints_bboxes_golden.py) --
    including that head's own keypoint clamp."""
    from tests._cpu_head import make_cpu_reppoints_head
    from tests.golden.gen_reppoints_bboxes_golden import IMG, NMS_PRE, make_case
    g = gold('reppoints_bboxes.npz')
    head = make_cpu_reppoints_head(variant)
    with torch.no_grad():
        head.moment_transfer.copy_(torch.tensor([0.25, -0.15]))
    cls, kpt, rep = make_case()
    dets, labels, kpts = head.get_bboxes(cls, kpt, rep, [IMG] * 2, float(g[variant + '_score_thr']),
                                         float(g[variant + '_iou_thr']), NMS_PRE, int(g[variant + '_max_per_img']))
    check_reppoints_bboxes(g, variant, dets, labels, kpts)
    # rescale=True: boxes / keypoints divided by the scale factor before the NMS (PAR:732-737)
    from tests.golden.gen_reppoints_bboxes_golden import SCALE
    dets, labels, kpts = head.get_bboxes(cls, kpt, rep, [IMG] * 2, float(g[variant + '_score_thr']),
                                         float(g[variant + '_iou_thr']), NMS_PRE, int(g[variant + '_max_per_img']),
                                         scale_factors=[SCALE] * 2)
    check_reppoints_bboxes(g, variant, dets, labels, kpts, tag='rs')


def check_kgdet_bboxes(g, dets, labels, kpts):
    for i in range(dets.shape[0]):
        rd, rl, rk = g['dets_%d' % i], g['labels_%d' % i], g['kpts_%d' % i]
        nv = int((labels[i] >= 0).sum())
        assert nv == rd.shape[0]
        o = np.argsort(-rd[:, 4], kind='stable')
        assert np.allclose(dets[i, :nv].cpu().numpy(), rd[o], rtol=0, atol=1e-4)
        assert np.array_equal(labels[i, :nv].cpu().numpy(), rl[o])
        assert np.allclose(kpts[i, :nv].cpu().numpy(), rk[o], rtol=0, atol=1e-3)


def test_head_mirror_get_bboxes_rescale_matches_reference_golden():
    """rescale=True (KP3:892-898) against the unchanged reference class (gen_kgdet_rescale_golden.py)."""
    from tests.golden.gen_kgdet_rescale_golden import SCALE
    g, gr = gold('get_bboxes.npz'), gold('get_bboxes_rescale.npz')
    head = _cpu_head()
    dets, labels, kpts = head.get_bboxes([torch.from_numpy(g['logit'])], [torch.from_numpy(g['kpt3'])],
                                         [torch.from_numpy(g['bbox3'])], [(800, 1333)] * 2, 0.05, 0.5, 1000, 100,
                                         scale_factors=[SCALE] * 2)
    check_kgdet_bboxes(gr, dets, labels, kpts)
