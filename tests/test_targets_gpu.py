"""The same golden comparison as tests/test_targets_cpu.py on the GPU, with the classification losses through the
package's fused CUDA focal-loss operator, plus gradient flow and CUDA-graph capturability (no host sync)."""
import pytest
import torch

from tests.test_targets_cpu import check_targets, run_case

pytestmark = pytest.mark.gpu


def test_targets_and_losses_match_reference_loss_gpu():
    g, tg, losses = run_case('cuda')
    check_targets(g, tg)
    for k, v in losses.items():
        assert abs(float(v) - float(g[k])) < 5e-5 * abs(float(g[k])), (k, float(v), float(g[k]))


def test_loss_step_is_graph_capturable_and_differentiable():
    """Assignment + losses contain no host synchronisation: the forward is captured into a CUDA graph and replayed on
    new predictions; gradients reach all nine head outputs.  (The full training step -- forward, this, backward,
    clip, SGD -- is captured by `bench.py --mode train`.)"""
    from kgdet_b200 import targets as T
    from tests.golden.gen_loss_golden import make_case
    outs, gt_bboxes, gt_labels, gt_kps, _ = make_case()
    outs = [o.cuda().requires_grad_() for o in outs]
    points = T.grid_points(13, 21, 32, 'cuda')
    boxes, labels, kps, valid = T.pad_ground_truth(gt_bboxes, gt_labels, gt_kps, device='cuda', max_gts=5)

    def total(o):
        tg = T.point_targets(points, boxes, labels, kps, valid)
        return sum(T.kgdet_losses(o, points, 32, tg).values())

    eager = total(outs)
    eager.backward()
    assert all(o.grad is not None and float(o.grad.abs().sum()) > 0 for o in outs)
    static = [o.detach().clone() for o in outs]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        total(static)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph), torch.no_grad():
        captured = total(static)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.allclose(captured, eager.detach(), rtol=1e-6, atol=0)
    with torch.no_grad():
        for st in static:
            st.mul_(0.5)
        want = total(static)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.allclose(captured, want, rtol=1e-6, atol=0) and not torch.allclose(captured, eager.detach(), rtol=1e-3)


def test_fused_assignment_and_loss_kernels_match_reference_loss_gpu():
    """kgdet_point_assign + kgdet_point_losses_forward/backward (csrc/point_loss.cu) against the golden of the
    UNCHANGED reference loss(): the assignment bit for bit (labels, positives, avg_factor), the nine losses to 5e-5,
    and the gradients against autograd through the PyTorch restatement (targets.py) to 1e-5 of each tensor's maximum."""
    import numpy as np
    from kgdet_b200 import targets as T
    from kgdet_b200.ops.point_loss import kgdet_point_losses
    from tests.golden.gen_loss_golden import make_case
    from tests.test_targets_cpu import GOLD
    g = np.load(GOLD)
    outs, gt_bboxes, gt_labels, gt_kps, _ = make_case()
    boxes, labels, kps, valid = T.pad_ground_truth(gt_bboxes, gt_labels, gt_kps, device='cuda', max_gts=5)
    outs_k = [o.cuda().requires_grad_() for o in outs]
    losses, (assigned, avg) = kgdet_point_losses(outs_k, boxes, labels, kps, valid, 32, return_targets=True)
    # assignment == the reference's targets
    lab = torch.where(assigned > 0, labels.gather(1, (assigned.long() - 1).clamp(min=0)), torch.zeros_like(assigned).long())
    assert np.array_equal(lab.cpu().numpy(), g['labels'])
    assert int(avg.item()) == int(g['num_total_pos'])
    points = T.grid_points(13, 21, 32, 'cuda')
    assert torch.equal(assigned.long(), T.point_assign(points, boxes, valid, 4, 25))
    for k, v in losses.items():
        assert abs(float(v) - float(g[k])) < 5e-5 * abs(float(g[k])), (k, float(v), float(g[k]))
    # gradients: arbitrary upstream weights on the nine losses
    wts = torch.linspace(0.5, 1.7, 9, device='cuda')
    sum(w * v for w, v in zip(wts, losses.values())).backward()
    outs_t = [o.cuda().requires_grad_() for o in outs]
    tg = T.point_targets(points, boxes, labels, kps, valid)
    ref = T.kgdet_losses(outs_t, points, 32, tg)
    order = ['loss_cls_1', 'loss_cls_2', 'loss_cls_3', 'loss_bbox_1', 'loss_bbox_2', 'loss_bbox_3', 'loss_kpt_1',
             'loss_kpt_2', 'loss_kpt_3']
    sum(w * ref[n] for w, n in zip(wts, order)).backward()
    for a, b in zip(outs_k, outs_t):
        assert a.grad is not None and (a.grad - b.grad).abs().max().item() <= 1e-5 * b.grad.abs().max().item() + 1e-12
    # the head uses the fused path; a box with no visible keypoint and an image without ground truth do not break it
    kps0 = kps.clone()
    kps0[0, 0, :, 2] = 0
    valid0 = valid.clone()
    valid0[1] = False
    l2 = kgdet_point_losses([o.detach() for o in outs_k], boxes, labels, kps0, valid0, 32)
    tg2 = T.point_targets(points, boxes, labels, kps0, valid0)
    r2 = T.kgdet_losses([o.detach() for o in outs_t], points, 32, tg2)
    for n in order:
        assert torch.isfinite(l2[n]) and abs(float(l2[n]) - float(r2[n])) <= 5e-5 * abs(float(r2[n])) + 1e-9, n
