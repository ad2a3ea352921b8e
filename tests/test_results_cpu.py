"""kgdet_b200.results against the UNCHANGED reference functions (run through tests/refshim.py when the reference
tree is present; the build container has it, the GPU box does not): bbox2result_kp
(reppoints_detector_kp.py:55-78) and kpt2json (coco_utils.py:121-154)."""
import numpy as np
import pytest
import torch

from tests import refshim


def _padded_batch(seed=0, B=3, k=10, P=294):
    g = torch.Generator().manual_seed(seed)
    dets = torch.rand(B, k, 5, generator=g) * 300
    dets[..., 2:4] += dets[..., :2]
    dets[..., 4] = torch.rand(B, k, generator=g)
    labels = torch.randint(0, 13, (B, k), generator=g)
    kpts = torch.rand(B, k, P * 3, generator=g) * 500
    labels[0, 6:] = -1                      # image 0: 6 detections
    labels[1, :] = -1                       # image 1: nothing
    for t in (dets, kpts):
        t[labels < 0] = 0
    return dets, labels, kpts


@pytest.mark.skipif(not refshim.available(), reason='reference tree not present')
def test_results_and_json_match_reference():
    from kgdet_b200 import results as R
    refshim.install('oracle')
    from mmdet.core.evaluation.coco_utils import kpt2json as ref_kpt2json
    from mmdet.models.detectors.reppoints_detector_kp import RepPointsDetectorKp
    dets, labels, kpts = _padded_batch()
    ours = R.batch_to_results(dets, labels, kpts, 14)
    ref = []
    for b in range(dets.shape[0]):
        keep = labels[b] >= 0
        ref.append(RepPointsDetectorKp.bbox2result_kp(None, dets[b][keep], labels[b][keep], kpts[b][keep], 14))
    for o, r in zip(ours, ref):
        assert len(o) == len(r)
        for a, c in zip(o, r):
            if isinstance(a, list):
                assert len(a) == len(c) and all(np.array_equal(x, y) for x, y in zip(a, c))
            else:
                assert np.array_equal(a, c)

    class DS(object):
        img_ids = [11, 22, 33]
        cat_ids = list(range(1, 14))

        def __len__(self):
            return 3
    got = R.kpt2json(DS.img_ids, DS.cat_ids, ours)
    want = ref_kpt2json(DS(), ref)
    assert got == want and len(got[0]) == 16 and len(got[1]) == 16


def test_empty_and_padding_conventions():
    from kgdet_b200 import results as R
    dets, labels, kpts = _padded_batch()
    res = R.batch_to_results(dets, labels, kpts, 14)
    assert len(res[1]) == 1 and all(a.shape == (0, 5) for a in res[1][0])       # nothing detected: 1-tuple
    assert sum(a.shape[0] for a in res[0][0]) == 6 and res[0][1].shape == (6,)
    bj, kj = R.kpt2json([1, 2, 3], list(range(1, 14)), res)
    assert {d['image_id'] for d in bj} == {1, 3} and len(kj[0]['keypoints']) == 294 * 3
    assert all(d['bbox'][2] >= 1 for d in bj)                                    # +1 pixel width convention
