"""`kgdet_b200.accelerate(ref_head)` on the CPU: the UNCHANGED reference head objects (built by the reference's own
builder from its configs, tests/refshim.py) rebound to this package's restatements with oracle-backed operators.
Checks the wiring -- parameter sharing, output formats, rescale, fall-through -- against the reference methods
themselves; the fused CUDA paths behind the same rebinding are checked in tests/test_reference_head_gpu.py."""
import numpy as np
import pytest
import torch

from tests import oracle_ops
from tests._cpu_head import cpu_batched_nms_flags
from tests._data import rel_err


def _inject():
    from oracle import moment_oracle
    return dict(deform_conv_cls=oracle_ops.DeformConv, moment_fn=moment_oracle.points2bbox_moment,
                nms_flags_fn=cpu_batched_nms_flags)


def _build(cfg_name):
    from tests import refshim
    if not refshim.available():
        pytest.skip('reference tree not present')
    from tests.golden.gen_golden import fill_state_dict
    refshim.install('oracle')
    head, cfg = refshim.build_head(cfg_name)
    head.load_state_dict(fill_state_dict(head.state_dict(), seed=99), strict=True)
    return head.eval(), cfg, refshim


def _same_results(a, b, flat):
    assert len(a) == len(b)
    for (d1, l1, k1), (d2, l2, k2) in zip(a, b):
        assert d1.shape == d2.shape and k1.shape == k2.shape and (k1.dim() == 2) == flat
        # the reference concatenates per class and sorts by score (bbox_nms_kp.py:64-70); scores are distinct here
        assert np.allclose(d1.numpy(), d2.numpy(), rtol=0, atol=1e-4)
        assert np.array_equal(l1.numpy(), l2.numpy())
        assert np.allclose(k1.numpy(), k2.numpy(), rtol=0, atol=1e-3)


def test_accelerate_kgdet_reference_head_instance():
    import kgdet_b200
    head, cfg, refshim = _build('kgdet_moment_r50_fpn_1x-demo.py')
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 256, 7, 11, generator=g)
    with torch.no_grad():
        want = head.forward_single(x)
    tc = refshim.AttrDict(cfg['test_cfg'])
    sc = torch.rand(2, 13, 7, 11, generator=g) ** 3
    logit = torch.log(sc / (1 - sc))
    args = ([want[0]], [want[1]], [logit], [want[3]], [want[4]], [want[5]], [want[6]], [want[7]], [want[8]])
    metas = [dict(img_shape=(800, 1333, 3), scale_factor=1.0)] * 2
    metas_rs = [dict(img_shape=(800, 1333, 3), scale_factor=1.6)] * 2
    with torch.no_grad():
        ref_plain = head.get_bboxes(*args, metas, tc, rescale=False)
        ref_rs = head.get_bboxes(*args, metas_rs, tc, rescale=True)
        ref_raw = head.get_bboxes(*args, metas, tc, False, False)          # nms=False
    assert kgdet_b200.accelerate(head, **_inject()) is head
    m = head.kgdet_mirror
    assert type(m).__name__ == 'KGDetHead'
    assert all(p is dict(head.named_parameters())[n] for n, p in m.named_parameters())
    assert m.cls_convs[0].conv.weight is head.cls_convs[0].conv.weight
    with torch.no_grad():
        got = head.forward_single(x)
        got_multi = head.forward((x,))
    for a, b in zip(got, want):
        assert rel_err(a, b) < 1e-5
    assert len(got_multi) == 9 and torch.equal(got_multi[2][0], got[2])
    with torch.no_grad():
        _same_results(head.get_bboxes(*args, metas, tc, rescale=False), ref_plain, flat=False)
        _same_results(head.get_bboxes(*args, metas_rs, tc, rescale=True), ref_rs, flat=True)
        raw = head.get_bboxes(*args, metas, tc, False, False)              # falls through to the reference method
    for a, b in zip(raw[0], ref_raw[0]):
        assert torch.equal(a, b)
    # an optimiser step through the reference object's parameters is what the restatement computes with next
    with torch.no_grad():
        head.kp_rep_block_1.cls_out.bias.add_(1.0)
        after = head.forward_single(x)
    assert torch.allclose(after[0], want[0] + 1.0, atol=1e-5)
    # training mode follows the reference object; gradients land on the reference's parameters
    head.train()
    out = head.forward_single(x)
    assert head.kgdet_mirror.training
    sum(o.sum() for o in out).backward()
    assert head.cls_convs[0].conv.weight.grad is not None and head.kp_rep_block_3.cls_dfmconv_7.weight.grad is not None


@pytest.mark.parametrize('variant', ['parallel', 'serial'])
def test_accelerate_reppoints_reference_head_instance(variant):
    import kgdet_b200
    from tests.golden.gen_reppoints_bboxes_golden import IMG, NMS_PRE, SCALE, make_case
    head, cfg, refshim = _build('reppoints_moment_%s_r50_fpn_1x-deepfashion2.py' % variant)
    g = torch.Generator().manual_seed(4)
    feats = [torch.randn(1, 256, h, w, generator=g) for h, w in [(13, 21), (7, 11)]]
    with torch.no_grad():
        want = [head.forward_single(x) for x in feats]
    cls, kpt, rep = make_case()
    tc = refshim.AttrDict(cfg['test_cfg'])
    tc['nms_pre'] = NMS_PRE
    metas = [dict(img_shape=IMG + (3,), scale_factor=1.0)] * 2
    metas_rs = [dict(img_shape=IMG + (3,), scale_factor=SCALE)] * 2
    call = lambda mt, rs: head.get_bboxes([c.clone() for c in cls], [k.clone() for k in kpt], [k.clone() for k in kpt],   # noqa: E731
                                          [r.clone() for r in rep], [r.clone() for r in rep], mt, tc, rescale=rs)
    with torch.no_grad():
        ref_plain, ref_rs = call(metas, False), call(metas_rs, True)
    kgdet_b200.accelerate(head, **_inject())
    assert type(head.kgdet_mirror).__name__ == 'RepPointsKpHead' and head.kgdet_mirror.variant == variant
    with torch.no_grad():
        for x, w in zip(feats, want):
            for a, b in zip(head.forward_single(x), w):
                assert rel_err(a, b) < 1e-5
        _same_results(call(metas, False), ref_plain, flat=False)
        _same_results(call(metas_rs, True), ref_rs, flat=True)
