"""Seeded synthetic inputs shared by the parity tests."""
import numpy as np
import torch


def rel_err(a, b):
    """max|a-b| / max|b|  -- the normalised error the tolerances are stated in."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def dcn_case(N, C, H, W, Cout, k, stride=1, pad=None, dil=1, groups=1, dg=1, mask=False, bias=False,
             seed=0, dtype=torch.float32, offset_std=2.0):
    g = torch.Generator().manual_seed(seed)
    pad = (dil * (k - 1)) // 2 if pad is None else pad
    Ho = (H + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    d = dict(
        x=torch.randn(N, C, H, W, generator=g).to(dtype),
        offset=(torch.randn(N, dg * 2 * k * k, Ho, Wo, generator=g) * offset_std),
        weight=(torch.randn(Cout, C // groups, k, k, generator=g) * (1.0 / (C * k * k) ** 0.5)),
        grad_out=torch.randn(N, Cout, Ho, Wo, generator=g).to(dtype),
        mask=torch.rand(N, dg * k * k, Ho, Wo, generator=g) if mask else None,
        bias=torch.randn(Cout, generator=g) if bias else None,
        stride=stride, padding=pad, dilation=dil, groups=groups, deformable_groups=dg)
    return d


def random_boxes(n, seed=0, img_w=1333.0, img_h=800.0, clustered=False):
    """[n,5] float32 boxes with all-distinct scores in (0.05, 1)."""
    g = torch.Generator().manual_seed(seed)
    if clustered:
        ng = max(n // 20, 1)
        cx0 = torch.rand(ng, generator=g) * img_w
        cy0 = torch.rand(ng, generator=g) * img_h
        w0 = torch.rand(ng, generator=g) * 300 + 60
        h0 = torch.rand(ng, generator=g) * 300 + 60
        idx = torch.randint(0, ng, (n,), generator=g)
        cx = cx0[idx] + torch.randn(n, generator=g) * 8
        cy = cy0[idx] + torch.randn(n, generator=g) * 8
        w = w0[idx] + torch.randn(n, generator=g) * 8
        h = h0[idx] + torch.randn(n, generator=g) * 8
    else:
        cx = torch.rand(n, generator=g) * img_w
        cy = torch.rand(n, generator=g) * img_h
        w = torch.rand(n, generator=g) * 384 + 16
        h = torch.rand(n, generator=g) * 384 + 16
    sc = (torch.randperm(n, generator=g).float() + 1) / (n + 1) * 0.95 + 0.05
    x1 = (cx - w / 2).clamp(0, img_w)
    y1 = (cy - h / 2).clamp(0, img_h)
    x2 = (cx + w / 2).clamp(0, img_w)
    y2 = (cy + h / 2).clamp(0, img_h)
    return torch.stack([x1, y1, x2, y2, sc], 1).float().contiguous()
