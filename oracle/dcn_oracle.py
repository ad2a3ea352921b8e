"""CPU oracle for deformable convolution (v1 and modulated v2).

TEST INFRASTRUCTURE ONLY.  Nothing under ``kgdet_b200/`` may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do.  It is the checker, never the thing shipped.

Restates, in plain PyTorch tensor ops on the CPU (any float dtype; use
``torch.float64`` for a tight checker), the algorithm of the reference's CUDA
deformable convolution (mmdet v1.0rc0 DCN is CUDA-only:
``mmdet/ops/dcn/deform_conv.py:44-45``):

* sampling rule, corner validity, bilinear weights:
  ``mmdet/ops/dcn/src/deform_conv_cuda_kernel.cu:83-114`` and ``:226-236``
* offset layout (dy at channel 2k, dx at 2k+1, tap k = i*kw + j), column row
  order ``c*K + k``: ``deform_conv_cuda_kernel.cu:205-224,238``
* GEMM orientation ``out = W[Cout, Cin*K] @ col``:
  ``mmdet/ops/dcn/src/deform_conv_cuda.cpp:225-234``
* backward: column gradient ``W^T @ gO`` (``deform_conv_cuda.cpp:330-331``),
  offset gradient through the bilinear derivative
  (``deform_conv_cuda_kernel.cu:144-187,372-435``), input gradient as the
  bilinear-weighted scatter (``:116-142,278-334``), weight gradient
  ``gO @ col^T`` (``deform_conv_cuda.cpp:443-461``)
* modulated variant (mask multiplies the sampled column, optional bias):
  ``deform_conv_cuda_kernel.cu:569-766`` and ``deform_conv_cuda.cpp:486-679``
* output size: ``mmdet/ops/dcn/deform_conv.py:96-110``

Parity pin: the reference ships no golden vectors for this path.
The pins are (1) ``tests/test_oracle_cpu.py`` — this restatement against
``torchvision.ops.deform_conv2d`` (independent implementation of the same
algorithm lineage) and against ``torch.autograd.gradcheck``-style finite
differences, and (2) ``tests/test_dcn_gpu.py::test_oracle_matches_reference_cuda``
— against the reference's own ``deform_conv_cuda`` sources compiled unmodified
for sm_100a into ``oracle/_ref`` and executed on the B200.
"""
import torch


def _out_size(size, k, stride, pad, dil):
    # mmdet/ops/dcn/deform_conv.py:96-110
    return (size + 2 * pad - (dil * (k - 1) + 1)) // stride + 1


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _sample_geometry(offset, H, W, kh, kw, stride, pad, dil, dg):
    """Sampling coordinates for every (n, dgroup, tap, y, x).

    Returns py, px with shape [N, dg, K, Ho, Wo] in offset's dtype.
    deform_conv_cuda_kernel.cu:210-211,221-227
    """
    N, _, Ho, Wo = offset.shape
    K = kh * kw
    off = offset.reshape(N, dg, K, 2, Ho, Wo)
    dt = offset.dtype
    ys = torch.arange(Ho, dtype=dt).view(1, 1, 1, Ho, 1) * stride[0] - pad[0]
    xs = torch.arange(Wo, dtype=dt).view(1, 1, 1, 1, Wo) * stride[1] - pad[1]
    ki = (torch.arange(K) // kw).to(dt).view(1, 1, K, 1, 1) * dil[0]
    kj = (torch.arange(K) % kw).to(dt).view(1, 1, K, 1, 1) * dil[1]
    py = ys + ki + off[:, :, :, 0]
    px = xs + kj + off[:, :, :, 1]
    return py, px


def _corners(py, px, H, W):
    """Corner indices, bilinear weights and validity masks.

    deform_conv_cuda_kernel.cu:88-110 (corner tests) and :228 (window test).
    Returns a list of 4 tuples (h_idx, w_idx, weight, valid) and the window mask,
    plus (lh, lw) for the derivative.
    """
    inside = (py > -1) & (px > -1) & (py < H) & (px < W)
    h_low = torch.floor(py)
    w_low = torch.floor(px)
    lh = py - h_low
    lw = px - w_low
    hh = 1 - lh
    hw = 1 - lw
    h_low = h_low.long()
    w_low = w_low.long()
    h_high = h_low + 1
    w_high = w_low + 1
    v_hl = h_low >= 0
    v_hh = h_high <= H - 1
    v_wl = w_low >= 0
    v_wh = w_high <= W - 1
    corners = [
        (h_low, w_low, hh * hw, v_hl & v_wl & inside),
        (h_low, w_high, hh * lw, v_hl & v_wh & inside),
        (h_high, w_low, lh * hw, v_hh & v_wl & inside),
        (h_high, w_high, lh * lw, v_hh & v_wh & inside),
    ]
    return corners, inside, (lh, lw, hh, hw)


def _gather(x_g, hi, wi, valid):
    """x_g: [N, dg, Cg, H, W]; hi/wi/valid: [N, dg, K, Ho, Wo] -> [N, dg, Cg, K, Ho, Wo]."""
    N, dg, Cg, H, W = x_g.shape
    K, Ho, Wo = hi.shape[2:]
    flat = (hi.clamp(0, H - 1) * W + wi.clamp(0, W - 1)).reshape(N, dg, 1, K * Ho * Wo)
    flat = flat.expand(N, dg, Cg, K * Ho * Wo)
    v = torch.gather(x_g.reshape(N, dg, Cg, H * W), 3, flat)
    v = v.reshape(N, dg, Cg, K, Ho, Wo)
    return v * valid.unsqueeze(2).to(v.dtype)


def deform_im2col(x, offset, ksize, stride=1, padding=0, dilation=1,
                  deformable_groups=1, mask=None):
    """Column tensor col[N, C, K, Ho, Wo] (deform_conv_cuda_kernel.cu:189-242)."""
    kh, kw = _pair(ksize)
    stride, padding, dilation = _pair(stride), _pair(padding), _pair(dilation)
    N, C, H, W = x.shape
    dg = deformable_groups
    py, px = _sample_geometry(offset, H, W, kh, kw, stride, padding, dilation, dg)
    corners, _, _ = _corners(py, px, H, W)
    x_g = x.reshape(N, dg, C // dg, H, W)
    col = None
    for hi, wi, wt, valid in corners:
        term = _gather(x_g, hi, wi, valid) * wt.unsqueeze(2)
        col = term if col is None else col + term
    if mask is not None:  # deform_conv_cuda_kernel.cu:626
        K = kh * kw
        col = col * mask.reshape(N, dg, 1, K, *mask.shape[2:])
    Ho, Wo = offset.shape[2:]
    return col.reshape(N, C, kh * kw, Ho, Wo)


def deform_conv_forward(x, offset, weight, stride=1, padding=0, dilation=1,
                        groups=1, deformable_groups=1, mask=None, bias=None):
    """out[N, Cout, Ho, Wo]; restates deform_conv_cuda.cpp:151-258 (and :486-564)."""
    Cout, Cin_g, kh, kw = weight.shape
    N, C, H, W = x.shape
    s, p, d = _pair(stride), _pair(padding), _pair(dilation)
    Ho = _out_size(H, kh, s[0], p[0], d[0])
    Wo = _out_size(W, kw, s[1], p[1], d[1])
    assert offset.shape == (N, deformable_groups * 2 * kh * kw, Ho, Wo), offset.shape
    col = deform_im2col(x, offset, (kh, kw), s, p, d, deformable_groups, mask)
    K = kh * kw
    col = col.reshape(N, groups, (C // groups) * K, Ho * Wo)
    w = weight.reshape(groups, Cout // groups, Cin_g * K)
    out = torch.einsum('gok,ngkp->ngop', w, col).reshape(N, Cout, Ho, Wo)
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return out


def deform_conv_backward(x, offset, weight, grad_out, stride=1, padding=0,
                         dilation=1, groups=1, deformable_groups=1, mask=None,
                         with_bias=False):
    """Explicit (non-autograd) backward following the reference kernels.

    Returns dict(grad_input, grad_offset, grad_weight[, grad_mask, grad_bias]).
    """
    Cout, Cin_g, kh, kw = weight.shape
    K = kh * kw
    N, C, H, W = x.shape
    dg = deformable_groups
    s, p, d = _pair(stride), _pair(padding), _pair(dilation)
    Ho, Wo = grad_out.shape[2:]
    # column gradient: W^T @ gO   (deform_conv_cuda.cpp:330-331)
    w = weight.reshape(groups, Cout // groups, Cin_g * K)
    go = grad_out.reshape(N, groups, Cout // groups, Ho * Wo)
    cg = torch.einsum('gok,ngop->ngkp', w, go).reshape(N, C, K, Ho, Wo)
    cg = cg.reshape(N, dg, C // dg, K, Ho, Wo)

    py, px = _sample_geometry(offset, H, W, kh, kw, s, p, d, dg)
    corners, inside, (lh, lw, hh, hw) = _corners(py, px, H, W)
    x_g = x.reshape(N, dg, C // dg, H, W)
    vals = [_gather(x_g, hi, wi, valid) for hi, wi, _, valid in corners]
    v1, v2, v3, v4 = vals
    m = None
    if mask is not None:
        m = mask.reshape(N, dg, 1, K, Ho, Wo)
    cgm = cg if m is None else cg * m
    # offset gradient (deform_conv_cuda_kernel.cu:163-184,405-433)
    dS_dy = (-hw.unsqueeze(2) * v1 - lw.unsqueeze(2) * v2
             + hw.unsqueeze(2) * v3 + lw.unsqueeze(2) * v4)
    dS_dx = (-hh.unsqueeze(2) * v1 + hh.unsqueeze(2) * v2
             - lh.unsqueeze(2) * v3 + lh.unsqueeze(2) * v4)
    g_dy = (cgm * dS_dy).sum(2)
    g_dx = (cgm * dS_dx).sum(2)
    grad_offset = torch.stack([g_dy, g_dx], dim=3).reshape(N, dg * 2 * K, Ho, Wo)
    # input gradient (deform_conv_cuda_kernel.cu:116-142,318-331)
    grad_input = torch.zeros(N, dg, C // dg, H * W, dtype=x.dtype)
    for hi, wi, wt, valid in corners:
        contrib = cgm * (wt * valid.to(wt.dtype)).unsqueeze(2)
        flat = (hi.clamp(0, H - 1) * W + wi.clamp(0, W - 1))
        flat = flat.reshape(N, dg, 1, K * Ho * Wo).expand(N, dg, C // dg, K * Ho * Wo)
        grad_input.scatter_add_(3, flat, contrib.reshape(N, dg, C // dg, K * Ho * Wo))
    grad_input = grad_input.reshape(N, C, H, W)
    # weight gradient (deform_conv_cuda.cpp:443-461)
    col = None
    for (hi, wi, wt, valid), v in zip(corners, vals):
        term = v * wt.unsqueeze(2)
        col = term if col is None else col + term
    out = {}
    if m is not None:
        # grad_mask = sum_c cg * S (deform_conv_cuda_kernel.cu:752,764)
        out['grad_mask'] = (cg * col).sum(2).reshape(N, dg * K, Ho, Wo)
        col = col * m
    colg = col.reshape(N, groups, (C // groups) * K, Ho * Wo)
    grad_weight = torch.einsum('ngop,ngkp->gok', go, colg).reshape(weight.shape)
    out.update(grad_input=grad_input, grad_offset=grad_offset, grad_weight=grad_weight)
    if with_bias:
        out['grad_bias'] = grad_out.sum(dim=(0, 2, 3))
    return out
