"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the Python mirror keeps the reference's surface, the product never touches the oracle,
and the UNCHANGED reference heads resolve their ops to this package."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'kgdet_b200.h')).read()
    return sorted(set(re.findall(r'KGDET_API[^;(]*?\b(kgdet_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from kgdet_b200.ops import _capi
    from kgdet_b200 import build
    if not os.path.exists(_capi.LIB_PATH):
        build.build()
    syms = _header_symbols()
    assert len(syms) >= 23
    handle = ctypes.CDLL(_capi.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), 'missing export %s' % s
    assert sorted(_capi.SIGNATURES) == syms, 'ctypes binding and header disagree'
    lib = _capi.lib()
    assert lib.kgdet_abi_version() == 2
    # argument validation works without a GPU (no kernel is launched on these paths)
    s = _capi.DcnShape(N=1, C=8, H=5, W=5, Cout=8, kh=0, kw=3, stride_h=1, stride_w=1, pad_h=1, pad_w=1,
                       dil_h=1, dil_w=1, groups=1, deformable_groups=1)
    assert lib.kgdet_dcn_forward_workspace_bytes(ctypes.byref(s), 0, 0) == 0
    assert b'kernel size' in lib.kgdet_last_error()
    s.kh = 3
    s.groups = 3
    assert lib.kgdet_dcn_packed_weight_bytes(ctypes.byref(s), 0) == 0
    assert b'groups' in lib.kgdet_last_error()
    assert lib.kgdet_nms_workspace_bytes(1000) >= 1000


def test_built_library_contains_blackwell_instructions():
    """The fused DCN kernel must be tcgen05/TMEM/bulk-copy code, not a recompiled legacy path."""
    from kgdet_b200.ops import _capi
    out = subprocess.run(['cuobjdump', '-sass', _capi.LIB_PATH], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    if not out:
        pytest.skip('cuobjdump unavailable')
    assert 'UTCHMMA' in out and 'LDTM' in out and 'UBLKCP' in out
    assert 'UBLKRED' in out                 # cp.reduce.async.bulk: the input gradient's col2im rows
    assert 'UTMALDG' in out                 # cp.async.bulk.tensor: the 3x3 convolution's tile loads
    assert 'HMMA.16816' not in out          # no mma.sync fallback anywhere


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'kgdet_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r'^\s*(from|import)\s+(oracle|tests)\b', txt, re.M) or 'oracle/' in txt and f.endswith('.py') and 'import' in txt and re.search(r'import.*oracle', txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_ops_fail_loudly_without_cuda():
    from kgdet_b200 import ops
    x = torch.randn(1, 8, 5, 5)
    with pytest.raises(NotImplementedError):
        ops.deform_conv(x, torch.zeros(1, 18, 5, 5), torch.randn(4, 8, 3, 3), 1, 1)
    with pytest.raises(NotImplementedError):
        ops.sigmoid_focal_loss(torch.randn(4, 13), torch.zeros(4, dtype=torch.long), 2.0, 0.25)
    with pytest.raises(NotImplementedError):
        ops.points2bbox_moment(torch.randn(2, 18, 3, 3), torch.zeros(2))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            ops.nms(torch.rand(4, 5), 0.5)
    with pytest.raises(NotImplementedError):
        ops.soft_nms(torch.rand(4, 5), 0.5)
    with pytest.raises(NotImplementedError):
        ops.RoIAlign(7, 1.0)


def test_mmdet_ops_surface_is_complete():
    from kgdet_b200 import ops
    reference_all = ['nms', 'soft_nms', 'RoIAlign', 'roi_align', 'RoIPool', 'roi_pool', 'DeformConv',
                     'DeformConvPack', 'DeformRoIPooling', 'DeformRoIPoolingPack',
                     'ModulatedDeformRoIPoolingPack', 'ModulatedDeformConv', 'ModulatedDeformConvPack',
                     'deform_conv', 'modulated_deform_conv', 'deform_roi_pooling', 'SigmoidFocalLoss',
                     'sigmoid_focal_loss', 'MaskedConv2d', 'ContextBlock']      # mmdet/ops/__init__.py:12-19
    for n in reference_all:
        assert hasattr(ops, n), n
    m = ops.DeformConv(16, 8, 3, padding=1)
    assert list(m.state_dict().keys()) == ['weight'] and m.weight.shape == (8, 16, 3, 3)
    with pytest.raises(AssertionError):
        ops.DeformConv(16, 8, 3, bias=True)                                       # DC.py:204
    md = ops.ModulatedDeformConv(16, 8, 3, padding=1)
    assert sorted(md.state_dict().keys()) == ['bias', 'weight']
    p = ops.ModulatedDeformConvPack(16, 8, 3, padding=1)
    assert p.conv_offset_mask.out_channels == 27
    assert ops.DeformConvPack(16, 8, 3, padding=1).conv_offset.out_channels == 18


@pytest.mark.parametrize('cfg_name,n_params', [
    ('kgdet_moment_r50_fpn_1x-demo.py', 27852247),
    ('kgdet_moment_r50_fpn_1x-deepfashion2.py', 27852247),
    ('reppoints_moment_parallel_r50_fpn_1x-deepfashion2.py', 6806475),
    ('reppoints_moment_serial_r50_fpn_1x-deepfashion2.py', 5638523),
])
def test_unchanged_reference_heads_load_this_package(cfg_name, n_params):
    """`from mmdet.ops import DeformConv` inside the reference heads resolves to kgdet_b200."""
    from tests import refshim
    if not refshim.available():
        pytest.skip('reference tree not present (GPU box)')
    from kgdet_b200 import ops
    refshim.install('kgdet')
    head, cfg = refshim.build_head(cfg_name)
    assert sum(p.numel() for p in head.parameters()) == n_params
    dcns = [m for m in head.modules() if type(m).__name__ == 'DeformConv']
    assert dcns and all(type(m) is ops.DeformConv for m in dcns)
    import mmdet.models.losses.focal_loss as fl
    import mmdet.core.post_processing.bbox_nms_kp as nk
    assert fl._sigmoid_focal_loss is ops.sigmoid_focal_loss
    assert nk.nms_wrapper.nms is ops.nms
    # a state dict written by the reference modules loads into ours and vice versa
    if 'kgdet' in cfg_name:
        from kgdet_b200.head import KGDetHead
        mine = KGDetHead()
        mine.load_state_dict(head.state_dict(), strict=True)
        head.load_state_dict(mine.state_dict(), strict=True)
