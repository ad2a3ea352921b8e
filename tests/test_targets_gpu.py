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
