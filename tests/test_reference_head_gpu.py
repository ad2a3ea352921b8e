"""The UNCHANGED reference head executed on the B200 on top of ``kgdet_b200.mount_as_mmdet_ops()``.

``RepPointsHeadKp3RepCas1AssignOnce`` (mmdet/models/anchor_heads/reppoints_head_kp3rep_cas_1_assign_once.py,
KP3) is imported from the reference's own Python package -- from /root/reference where it exists, otherwise from
the untouched archive ``oracle/_ref/pytree.zip`` that ``oracle/build_ref.py`` stages for the GPU box -- with
``mmdet.ops`` resolved to this package.  Its ``forward_single`` (KP3:412-446), ``get_bboxes`` (KP3:770-914 ->
multiclass_nms_kp -> nms_wrapper.nms) and ``loss`` (KP3:670-768 -> FocalLoss -> sigmoid_focal_loss) then run
through the C-ABI library on the device and are compared with the golden outputs the same class produced on
the CPU (tests/golden/gen_golden.py) and with this package's head mirror.
"""
import os

import numpy as np
import pytest
import torch

from tests._data import rel_err
from tests.golden.gen_golden import fill_state_dict

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
NAMES = ['cls_1', 'cls_2', 'cls_3', 'kpt_1', 'kpt_2', 'kpt_3', 'bbox_1', 'bbox_2', 'bbox_3']


class _fp32_cudnn(object):
    def __enter__(self):
        self.old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *a):
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.old


def _reference_head(cfg_name='kgdet_moment_r50_fpn_1x-demo.py'):
    from tests import refshim
    if not refshim.available():
        pytest.skip('reference python tree not present (neither /root/reference nor oracle/_ref/pytree.zip)')
    from kgdet_b200 import ops
    refshim.install('kgdet')
    head, cfg = refshim.build_head(cfg_name, device='cuda')
    assert type(head).__module__.startswith('mmdet.models.anchor_heads')
    dcns = [m for m in head.modules() if type(m).__name__ == 'DeformConv']
    assert len(dcns) == 12 and all(type(m) is ops.DeformConv for m in dcns)
    head.load_state_dict(fill_state_dict(head.state_dict()), strict=True)
    return head, cfg, refshim


@pytest.mark.parametrize('precision,tol', [(None, 5e-4), ('bf16', 3e-2)])
def test_unchanged_reference_head_forward_on_the_b200(precision, tol):
    """forward_single of the reference class, DeformConv served by this library (default precision for fp32
    tensors = the exact path; 'bf16' = the fused tcgen05 kernel), against the golden from the same class on the
    CPU and against this package's head mirror (module path: identical op sequence -> identical results)."""
    from kgdet_b200 import ops
    from kgdet_b200.head import KGDetHead
    head, _, _ = _reference_head()
    head.eval()
    g = np.load(os.path.join(GOLD, 'head_p7.npz'))
    x = torch.from_numpy(g['x']).cuda()
    mirror = KGDetHead()
    mirror.load_state_dict(head.state_dict(), strict=True)
    mirror = mirror.cuda().eval()
    mirror._fused_inference = False
    from kgdet_b200.ops import _capi
    launches0 = _capi.lib().kgdet_launch_count()
    ops.set_precision(precision)
    try:
        with _fp32_cudnn(), torch.no_grad():
            out = head.forward_single(x)
            mine = mirror.forward_single(x)
    finally:
        ops.set_precision(None)
    assert _capi.lib().kgdet_launch_count() > launches0          # the library did launch kernels
    for n, o, m in zip(NAMES, out, mine):
        e = rel_err(o, torch.from_numpy(g[n]))
        assert e < tol, (n, e)
        assert rel_err(o, m) < 1e-6, (n, rel_err(o, m))           # same ops in the same order


def test_unchanged_reference_head_full_map_on_the_b200():
    """[1,256,25,42] -- BASELINE configs[0]'s input shape -- through the reference class on the device against the
    checksums / strided samples of its own CPU run (head_p5.npz)."""
    head, _, _ = _reference_head()
    head.eval()
    g = np.load(os.path.join(GOLD, 'head_p5.npz'))
    with _fp32_cudnn(), torch.no_grad():
        out = head.forward_single(torch.from_numpy(g['x']).cuda())
    for n, o in zip(NAMES, out):
        f = o.reshape(-1)
        step = max(f.numel() // 4096, 1)
        assert rel_err(f[::step][:4096].cpu(), torch.from_numpy(g[n + '_sample'])) < 3e-3, n
        assert abs(o.double().abs().sum().item() - float(g[n + '_abs'])) / float(g[n + '_abs']) < 1e-4, n


def test_unchanged_reference_get_bboxes_on_the_b200():
    """get_bboxes of the reference class on CUDA tensors: its per-class loop calls nms_wrapper.nms -> this
    library's kgdet_nms (13 launches per image).  Same detections as the golden (reference code + nms_cpu.cpp on
    the CPU; no pair sits at IoU == thr exactly, so the '>' / '>=' comparators agree) and as the mirror's batched,
    sync-free get_bboxes."""
    from kgdet_b200.head import KGDetHead
    head, cfg, refshim = _reference_head()
    head.eval()
    g = np.load(os.path.join(GOLD, 'get_bboxes.npz'))
    g7 = np.load(os.path.join(GOLD, 'head_p7.npz'))
    t = lambda a: torch.from_numpy(a).cuda()
    tc = refshim.AttrDict(cfg['test_cfg'])
    metas = [dict(img_shape=(800, 1333, 3), scale_factor=1.0)] * 2
    dummy = [t(g7['cls_1'])]
    with torch.no_grad():
        res = head.get_bboxes(dummy, dummy, [t(g['logit'])], [t(g7['kpt_1'])], [t(g7['kpt_2'])], [t(g['kpt3'])],
                              [t(g7['bbox_1'])], [t(g7['bbox_2'])], [t(g['bbox3'])], metas, tc, rescale=False)
    mirror = KGDetHead().cuda().eval()
    dets, labels, kpts = mirror.get_bboxes([t(g['logit'])], [t(g['kpt3'])], [t(g['bbox3'])], [(800, 1333)] * 2,
                                           0.05, 0.5, 1000, 100)
    for i, (d, l, k) in enumerate(res):
        rd, rl, rk = g['dets_%d' % i], g['labels_%d' % i], g['kpts_%d' % i]
        assert d.is_cuda and d.shape[0] == rd.shape[0]
        assert np.allclose(d.cpu().numpy(), rd, rtol=0, atol=1e-3)
        assert np.array_equal(l.cpu().numpy(), rl)
        assert np.allclose(k.reshape(d.shape[0], -1).cpu().numpy(), rk, rtol=0, atol=1e-3)
        # the mirror returns the same detections sorted by score in fixed-size slots
        nv = int((labels[i] >= 0).sum())
        o = np.argsort(-rd[:, 4], kind='stable')
        assert nv == rd.shape[0] and np.allclose(dets[i, :nv].cpu().numpy(), rd[o], rtol=0, atol=1e-3)


def test_unchanged_reference_loss_and_backward_on_the_b200():
    """loss() of the reference class on the device (its FocalLoss calls this library's sigmoid focal loss, its
    points2bbox stays PyTorch) + backward through forward_single (DeformConv backward of this library): the nine
    loss values against the golden of the same code on the CPU.  The reference's PointAssigner indexes a CPU
    `arange` with CUDA masks (core/bbox/assigners/point_assigner.py:70-76) -- a torch-1.x idiom outside the ops
    boundary; tests/refshim.py runs that method (body unchanged) with the points' device as the default device."""
    from tests.golden.gen_loss_golden import make_case
    head, cfg, refshim = _reference_head()
    refshim.patch_point_assigner_device()
    head.train()
    g = np.load(os.path.join(GOLD, 'kgdet_loss.npz'))
    outs, gt_bboxes, gt_labels, gt_kps, (ih, iw) = make_case()
    outs = [o.cuda().requires_grad_() for o in outs]
    B = outs[0].shape[0]
    metas = [dict(img_shape=(ih, iw, 3), pad_shape=(ih, iw, 3), scale_factor=1.0, flip=False)] * B
    tc = refshim.AttrDict(uniform=refshim.AttrDict(cfg['train_cfg']['uniform']))
    losses = head.loss(*[[o] for o in outs], [b.clone().cuda() for b in gt_bboxes], [l.cuda() for l in gt_labels],
                       [k.cuda() for k in gt_kps], metas, tc)
    for k, v in losses.items():
        assert abs(float(v[0]) - float(g[k])) < 5e-5 * abs(float(g[k])), (k, float(v[0]), float(g[k]))
    sum(v[0] for v in losses.values()).backward()
    assert all(o.grad is not None and torch.isfinite(o.grad).all() for o in outs)
    # and a forward + backward through the twelve DeformConv of the reference class
    x = torch.randn(1, 256, 7, 8, device='cuda')
    o = head.forward_single(x)
    sum(t.sum() for t in o).backward()
    for n, p in head.named_parameters():
        if 'dfmconv' in n:
            assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().sum()) > 0, n


def test_wire_formats_from_the_batched_gpu_output_match_the_reference_functions():
    """The detections of the batched, sync-free `KGDetHead.get_bboxes` (padded [B, 100, ...] device tensors)
    through this package's wire formats (`results.batch_to_results`, `results.kpt2json`) against the UNCHANGED
    reference functions `bbox2result_kp` (reppoints_detector_kp.py:55-78) and `kpt2json` (coco_utils.py:121-154)
    fed with the reference head's own `get_bboxes` output for the same inputs -- end to end on the GPU box."""
    from kgdet_b200 import results as R
    from kgdet_b200.head import KGDetHead
    head, cfg, refshim = _reference_head()
    head.eval()
    from mmdet.core.evaluation.coco_utils import kpt2json as ref_kpt2json
    from mmdet.models.detectors.reppoints_detector_kp import RepPointsDetectorKp
    g = np.load(os.path.join(GOLD, 'get_bboxes.npz'))
    g7 = np.load(os.path.join(GOLD, 'head_p7.npz'))
    t = lambda a: torch.from_numpy(a).cuda()
    tc = refshim.AttrDict(cfg['test_cfg'])
    metas = [dict(img_shape=(800, 1333, 3), scale_factor=1.0)] * 2
    dummy = [t(g7['cls_1'])]
    with torch.no_grad():
        ref_dets = head.get_bboxes(dummy, dummy, [t(g['logit'])], [t(g7['kpt_1'])], [t(g7['kpt_2'])], [t(g['kpt3'])],
                                   [t(g7['bbox_1'])], [t(g7['bbox_2'])], [t(g['bbox3'])], metas, tc, rescale=True)
    ref_res = [RepPointsDetectorKp.bbox2result_kp(None, d, l, k, 14) for d, l, k in ref_dets]

    class DS(object):
        img_ids = [101, 202]
        cat_ids = list(range(1, 14))

        def __len__(self):
            return 2
    want_bbox, want_kpt = ref_kpt2json(DS(), ref_res)
    mirror = KGDetHead().cuda().eval()
    dets, labels, kpts = mirror.get_bboxes([t(g['logit'])], [t(g['kpt3'])], [t(g['bbox3'])], [(800, 1333)] * 2,
                                           0.05, 0.5, 1000, 100)
    got_bbox, got_kpt = R.kpt2json(DS.img_ids, DS.cat_ids, R.batch_to_results(dets, labels, kpts, 14))
    assert len(got_bbox) == len(want_bbox) > 0 and len(got_kpt) == len(want_kpt)
    key = lambda d: (d['image_id'], d['category_id'], -d['score'])
    for a, b in zip(sorted(got_bbox, key=key), sorted(want_bbox, key=key)):
        assert a['image_id'] == b['image_id'] and a['category_id'] == b['category_id']
        assert abs(a['score'] - b['score']) <= 1e-4 and np.allclose(a['bbox'], b['bbox'], rtol=0, atol=2e-3)
    for a, b in zip(sorted(got_kpt, key=key), sorted(want_kpt, key=key)):
        assert a['image_id'] == b['image_id'] and a['category_id'] == b['category_id']
        assert np.allclose(a['keypoints'], b['keypoints'], rtol=0, atol=2e-3)


def _same_lists(a, b, flat):
    assert len(a) == len(b)
    for (d1, l1, k1), (d2, l2, k2) in zip(a, b):
        assert d1.shape == d2.shape and k1.shape == k2.shape and (k1.dim() == 2) == flat
        assert np.allclose(d1.cpu().numpy(), d2.cpu().numpy(), rtol=0, atol=1e-3)
        assert np.array_equal(l1.cpu().numpy(), l2.cpu().numpy())
        assert np.allclose(k1.cpu().numpy(), k2.cpu().numpy(), rtol=0, atol=1e-3)


def test_accelerate_rebinds_the_unchanged_kgdet_head_to_the_fused_paths():
    """`kgdet_b200.accelerate(ref_head)` on the B200: the reference OBJECT (its own parameters, its own method
    signatures and result formats) now runs forward_single on the fused tensor-core path (bf16 mode: grouped DCN,
    own convolutions, 1x1 GEMMs; no cuDNN kernel), get_bboxes on the decode kernels + ONE batched NMS launch --
    rescale=True included, rows in the reference's order -- and loss() on the assignment / loss kernels.  Compared with
    what the same object returned BEFORE the rebinding (its own methods on top of the mounted ops)."""
    import kgdet_b200
    from kgdet_b200 import ops
    from kgdet_b200.ops import _capi
    from tests.golden.gen_loss_golden import make_case
    head, cfg, refshim = _reference_head()
    refshim.patch_point_assigner_device()
    head.eval()
    g = np.load(os.path.join(GOLD, 'get_bboxes.npz'))
    g7 = np.load(os.path.join(GOLD, 'head_p7.npz'))
    t = lambda a: torch.from_numpy(a).cuda()                               # noqa: E731
    x = t(g7['x'])
    tc = refshim.AttrDict(cfg['test_cfg'])
    metas = [dict(img_shape=(800, 1333, 3), scale_factor=1.0)] * 2
    metas_rs = [dict(img_shape=(800, 1333, 3), scale_factor=1.6)] * 2
    dummy = [t(g7['cls_1'])]
    bb_args = (dummy, dummy, [t(g['logit'])], [t(g7['kpt_1'])], [t(g7['kpt_2'])], [t(g['kpt3'])], [t(g7['bbox_1'])],
               [t(g7['bbox_2'])], [t(g['bbox3'])])
    outs, gt_bboxes, gt_labels, gt_kps, (ih, iw) = make_case()
    lmetas = [dict(img_shape=(ih, iw, 3), pad_shape=(ih, iw, 3), scale_factor=1.0, flip=False)] * outs[0].shape[0]
    trc = refshim.AttrDict(uniform=refshim.AttrDict(cfg['train_cfg']['uniform']))

    def run_loss():
        o = [v.cuda().requires_grad_() for v in outs]
        losses = head.loss(*[[v] for v in o], [b.clone().cuda() for b in gt_bboxes], [l.cuda() for l in gt_labels],
                           [k.cuda() for k in gt_kps], lmetas, trc)
        sum(v[0] for v in losses.values()).backward()
        return {k: float(v[0]) for k, v in losses.items()}, [v.grad for v in o]

    with _fp32_cudnn(), torch.no_grad():
        want = head.forward_single(x)
        ref_plain = head.get_bboxes(*bb_args, metas, tc, rescale=False)
        ref_rs = head.get_bboxes(*bb_args, metas_rs, tc, rescale=True)
    ref_losses, ref_grads = run_loss()

    kgdet_b200.accelerate(head)
    assert all(p is dict(head.named_parameters())[n] for n, p in head.kgdet_mirror.named_parameters())
    n0 = _capi.lib().kgdet_launch_count()
    with torch.no_grad():
        got32 = head.forward_single(x)                    # fp32 tensors: the fp32-grade tensor-core path
        ops.set_precision('bf16')
        try:
            got16 = head.forward_single(x)
        finally:
            ops.set_precision(None)
        plain = head.get_bboxes(*bb_args, metas, tc, rescale=False)
        rs = head.get_bboxes(*bb_args, metas_rs, tc, rescale=True)
    assert _capi.lib().kgdet_launch_count() > n0
    for n, a, b, c in zip(NAMES, got32, got16, want):
        assert rel_err(a, c) < 2e-3, (n, rel_err(a, c))
        assert rel_err(b, c) < 3e-2, (n, rel_err(b, c))
    _same_lists(plain, ref_plain, flat=False)
    _same_lists(rs, ref_rs, flat=True)
    losses, grads = run_loss()
    assert set(losses) == set(ref_losses)
    for k in losses:
        assert abs(losses[k] - ref_losses[k]) < 5e-5 * abs(ref_losses[k]), (k, losses[k], ref_losses[k])
    for a, b in zip(grads, ref_grads):
        assert (a - b).abs().max().item() <= 2e-5 * b.abs().max().item() + 1e-12


@pytest.mark.parametrize('variant', ['parallel', 'serial'])
def test_accelerate_rebinds_the_unchanged_reppoints_heads(variant):
    """The RepPoints-Kp baseline head objects: forward over FPN levels on the position-major fused path and the
    multi-level get_bboxes, against the same object's own methods before the rebinding."""
    import kgdet_b200
    from kgdet_b200 import ops
    from tests import refshim
    from tests.golden.gen_reppoints_bboxes_golden import IMG, NMS_PRE, SCALE, make_case
    if not refshim.available():
        pytest.skip('reference python tree not present')
    refshim.install('kgdet')
    head, cfg = refshim.build_head('reppoints_moment_%s_r50_fpn_1x-deepfashion2.py' % variant, device='cuda')
    head.load_state_dict(fill_state_dict(head.state_dict(), seed=4321), strict=True)
    head.eval()
    gen = torch.Generator().manual_seed(12)
    feats = [torch.randn(2, 256, h, w, generator=gen).cuda() for h, w in [(50, 84), (25, 42), (13, 21), (7, 11), (4, 6)]]
    cls, kpt, rep = [[v.cuda() for v in vs] for vs in make_case()]
    tc = refshim.AttrDict(cfg['test_cfg'])
    tc['nms_pre'] = NMS_PRE
    metas = [dict(img_shape=IMG + (3,), scale_factor=1.0)] * 2
    metas_rs = [dict(img_shape=IMG + (3,), scale_factor=SCALE)] * 2
    call = lambda mt, r: head.get_bboxes([c.clone() for c in cls], [k.clone() for k in kpt], [k.clone() for k in kpt],   # noqa: E731
                                         [v.clone() for v in rep], [v.clone() for v in rep], mt, tc, rescale=r)
    with _fp32_cudnn(), torch.no_grad():
        want = head.forward(feats, None)
        ref_plain, ref_rs = call(metas, False), call(metas_rs, True)
    kgdet_b200.accelerate(head)
    ops.set_precision('bf16')
    try:
        with _fp32_cudnn(), torch.no_grad():
            got = head.forward(feats, None)
    finally:
        ops.set_precision(None)
    assert len(got) == 5 and len(got[0]) == len(feats)
    for j in range(5):
        for li in range(len(feats)):
            assert rel_err(got[j][li], want[j][li]) < 3e-2, (j, li, rel_err(got[j][li], want[j][li]))
    with torch.no_grad():
        _same_lists(call(metas, False), ref_plain, flat=False)
        _same_lists(call(metas_rs, True), ref_rs, flat=True)
