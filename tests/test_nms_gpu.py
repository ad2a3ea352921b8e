"""GPU parity tests of NMS: keep indices must be BIT-EXACT against the oracle
(oracle/nms_oracle.c, pinned to the reference's nms_cpu.cpp) for distinct scores."""
import numpy as np
import pytest
import torch

from oracle import build_ref, nms_oracle
from tests._data import random_boxes

pytestmark = pytest.mark.gpu


def _ours_keep(dets, thr, cmp_mode):
    from kgdet_b200.ops.nms import nms_wrapper
    return nms_wrapper._nms_keep_cuda(dets.cuda(), thr, cmp_mode).cpu().numpy()


@pytest.mark.parametrize('n', [1, 2, 63, 64, 65, 1000, 3350, 4096])
@pytest.mark.parametrize('clustered', [False, True])
@pytest.mark.parametrize('cmp_mode', [0, 1])
def test_single_cta_path_bit_exact(n, clustered, cmp_mode):
    dets = random_boxes(n, seed=n, clustered=clustered)
    ref = nms_oracle.nms_keep(dets, 0.5, cmp_mode)
    got = _ours_keep(dets, 0.5, cmp_mode)
    assert got.dtype == np.int64 and np.array_equal(got, ref)


@pytest.mark.parametrize('n', [4097, 10000, 20011])
@pytest.mark.parametrize('clustered', [True, False])
def test_large_path_bit_exact(n, clustered):
    """mask kernel + pipelined sweep (chain warp + applier warps); 20 011 boxes = 313 blocks, the last one ragged."""
    dets = random_boxes(n, seed=n, clustered=clustered)
    for cmp_mode in (0, 1):
        assert np.array_equal(_ours_keep(dets, 0.5, cmp_mode), nms_oracle.nms_keep(dets, 0.5, cmp_mode))


def test_comparators_differ_only_at_equality():
    # two boxes with IoU exactly 0.5: (0,0,9,9) area 100; (0,0,9,4) area 50 -> inter 50 / union 100
    dets = torch.tensor([[0., 0., 9., 9., 0.9], [0., 0., 9., 4., 0.8]])
    assert list(_ours_keep(dets, 0.5, 0)) == [0, 1]      # '>'  keeps both (nms_kernel.cu:60)
    assert list(_ours_keep(dets, 0.5, 1)) == [0]         # '>=' suppresses (nms_cpu.cpp:55)
    assert list(nms_oracle.nms_keep(dets, 0.5, 0)) == [0, 1]
    assert list(nms_oracle.nms_keep(dets, 0.5, 1)) == [0]


def test_wrapper_contract_tensor_numpy_empty():
    from kgdet_b200.ops import nms
    dets = random_boxes(500, seed=3)
    kept, inds = nms(dets.cuda(), 0.5)
    assert inds.dtype == torch.long and inds.is_cuda and torch.equal(kept, dets.cuda()[inds])
    assert np.array_equal(inds.cpu().numpy(), nms_oracle.nms_keep(dets, 0.5, 0))
    # numpy in -> numpy out (nms_wrapper.py:29-32,47-48); CPU inputs use the nms_cpu comparator
    kept_np, inds_np = nms(dets.numpy(), 0.5)
    assert isinstance(inds_np, np.ndarray) and np.array_equal(inds_np, nms_oracle.nms_keep(dets, 0.5, 1))
    kept_np2, inds_np2 = nms(dets.numpy(), 0.5, device_id=0)
    assert np.array_equal(inds_np2, nms_oracle.nms_keep(dets, 0.5, 0))
    e, ei = nms(torch.zeros(0, 5).cuda(), 0.5)                 # nms_wrapper.py:39-40
    assert e.shape == (0, 5) and ei.numel() == 0 and ei.dtype == torch.long
    with pytest.raises(TypeError):
        nms([[0, 0, 1, 1, 1]], 0.5)


def test_against_reference_nms_cpu_module():
    ref = build_ref.load('nms_cpu')
    if ref is None:
        pytest.skip('oracle/_ref/nms_cpu.so not built')
    for seed in range(3):
        dets = random_boxes(2000, seed=seed, clustered=bool(seed % 2))
        want = ref.nms(dets, 0.5).numpy()
        assert np.array_equal(_ours_keep(dets, 0.5, 1), want)


def test_batched_segments_match_per_segment_oracle():
    from kgdet_b200.ops import batched_nms_flags
    lens = [0, 1, 700, 64, 0, 1000, 333]
    parts = [random_boxes(max(l, 1), seed=10 + i, clustered=True)[:l] for i, l in enumerate(lens)]
    dets = torch.cat(parts)
    offs = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32)
    flags = batched_nms_flags(dets.cuda(), offs.cuda(), max(lens), 0.5).cpu().numpy()
    want = np.zeros(len(dets), dtype=np.uint8)
    for i, l in enumerate(lens):
        if l:
            want[int(offs[i]) + nms_oracle.nms_keep(parts[i], 0.5, 0)] = 1
    assert np.array_equal(flags, want)


def test_idempotent_and_sorted():
    """Size-independent properties: NMS of the kept set keeps everything; indices ascend."""
    dets = random_boxes(4000, seed=7, clustered=True)
    keep = _ours_keep(dets, 0.5, 0)
    assert np.all(np.diff(keep) > 0)
    again = _ours_keep(dets[torch.from_numpy(keep)], 0.5, 0)
    assert np.array_equal(again, np.arange(len(keep)))


def test_batched_dense_mode_with_score_threshold():
    """Dense segments + in-kernel `score > thr` filter == compact-then-NMS per segment."""
    from kgdet_b200.ops import batched_nms_flags
    nseg, n = 7, 300
    parts = [random_boxes(n, seed=40 + i, clustered=True) for i in range(nseg)]
    for i, p in enumerate(parts):        # push a varying share of each segment below the threshold
        p[:, 4] = p[:, 4] * (0.3 + 0.1 * i)
    dets = torch.cat(parts)
    flags = batched_nms_flags(dets.cuda(), None, n, 0.5, score_thr=0.2).cpu().numpy()
    want = np.zeros(len(dets), dtype=np.uint8)
    for i, p in enumerate(parts):
        rows = np.nonzero(p[:, 4].numpy() > 0.2)[0]
        if len(rows):
            want[i * n + rows[nms_oracle.nms_keep(p[rows], 0.5, 0)]] = 1
    assert np.array_equal(flags, want)
    none = batched_nms_flags(dets.cuda(), None, n, 0.5, score_thr=2.0).cpu().numpy()
    assert none.sum() == 0


def test_topk_flagged_matches_torch_topk():
    """Top-k over the NMS survivors (bbox_nms_kp.py:64-70) in one CTA per image == torch.where + torch.topk on
    distinct scores: plenty of survivors, fewer than k, none at all, and a full-size 13 x 1000 candidate list."""
    import torch
    from kgdet_b200.ops.decode import topk_flagged
    g = torch.Generator().manual_seed(3)
    for B, L, k, p_keep in [(3, 500, 100, 0.5), (2, 300, 100, 0.1), (2, 64, 10, 0.0), (4, 13000, 100, 0.08),
                            (1, 16384, 100, 1.0)]:
        dets = torch.rand(B, L, 5, generator=g).cuda()
        flags = (torch.rand(B, L, generator=g) < p_keep).to(torch.uint8).cuda()
        top_s, top_i = topk_flagged(dets, flags, k)
        masked = torch.where(flags.bool(), dets[..., 4], dets.new_full((), -1.0))
        ws, wi = masked.topk(k, dim=1)
        assert torch.equal(top_s, ws)
        kept = ws > 0
        assert torch.equal(top_i[kept], wi[kept])
        assert (top_i[~kept] == 0).all()
