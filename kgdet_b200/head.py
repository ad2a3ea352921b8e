"""KGDet point-set head on the kgdet_b200 operators (host-side mirror for the GPU box).

The reference head (``RepPointsHeadKp3RepCas1AssignOnce``,
mmdet/models/anchor_heads/reppoints_head_kp3rep_cas_1_assign_once.py:183-914, abbreviated KP3)
stays the caller in a real deployment: it imports ``DeformConv`` / ``nms`` / focal loss from
``mmdet.ops`` and runs unchanged on top of ``kgdet_b200.mount_as_mmdet_ops()``.  The reference
tree is not present on the benchmark box, so this module restates the head's data flow from
scratch -- same parameter names and shapes (state dicts are interchangeable, 53 entries,
27 852 247 parameters for the KGDet configs), same maths -- to drive the operators end to end:

* ``forward_single``   KP3:412-446  towers -> stage 1 (plain convs) -> stages 2, 3 (6 deformable
                        convolutions each on 3x3 / 5x5 / 7x7 point sets) with the fused moment
                        transform (KP3:342-391) after every stage;
* ``get_bboxes``       KP3:770-914 + multiclass_nms_kp (core/post_processing/bbox_nms_kp.py:6-75):
                        sigmoid, top-k ``nms_pre``, decode, clamp, then ONE batched NMS launch over
                        all (image, class) segments instead of a 13-iteration Python loop with a
                        host sync each, and a fixed-size top-``max_per_img`` selection -- no host
                        synchronisation anywhere before the results are copied out.

Tower convolutions / GroupNorm / 1x1 heads are outside the hot path
and stay PyTorch/cuDNN.
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .ops import (DeformConv, TiledRows, batched_nms_flags, deform_conv_prepared, deform_conv_prepared_group, get_precision,
                  groupnorm_relu_nhwc, nchw_to_tiled, pack_weight, points2bbox_moment, pointwise_conv,
                  prepare_input, prepare_plan, prepare_plan_points)
from .ops.conv import conv_planes, conv_planes_pair, conv_supported, groupnorm_relu_planes, split_planes
from .ops.decode import bbox_decode, bbox_finalize, bbox_select, topk_flagged
from .ops.pointwise import cached, groupnorm_relu_nhwc_autograd, to_channels_last

_POINT_SETS = (3, 5, 7)          # KP3:257: 9 + 25 + 49 points regardless of cfg.num_reppts


class _ConvGNReLU(nn.Module):
    """ConvModule(conv3x3 no-bias + GroupNorm + ReLU) with the reference's child names
    (`conv`, `gn`: mmdet/models/utils/conv_module.py:96-126)."""

    def __init__(self, cin, cout, num_groups=32):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, stride=1, padding=1, bias=False)
        self.gn = nn.GroupNorm(num_groups, cout)

    def forward(self, x):
        return F.relu(self.gn(self.conv(x)))


def _normal(m, std=0.01, bias=0.0):
    nn.init.normal_(m.weight, 0, std)
    if getattr(m, 'bias', None) is not None:
        nn.init.constant_(m.bias, bias)


_side_streams = {}


def _streams(device, n):
    """`n` cached side streams of `device` (created once: stream creation is not allowed during graph capture)."""
    key = (device.type, device.index)
    lst = _side_streams.setdefault(key, [])
    while len(lst) < n:
        lst.append(torch.cuda.Stream(device=device))
    return lst[:n]


class _null_context(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _cl_weight(w):
    """channels_last copy of a convolution weight (cached per parameter version): cuDNN then runs the
    convolution on channels_last activations without inserting layout transposes."""
    return cached((w,), lambda: w.detach().contiguous(memory_format=torch.channels_last), tag='channels_last')


def _conv3x3_nhwc(x_cl, conv, with_bias=True):
    b = None if (conv.bias is None or not with_bias) else conv.bias.detach()
    return F.conv2d(x_cl, _cl_weight(conv.weight), b, conv.stride, conv.padding)


def _pointwise_weights(block):
    """Packed GEMM operands of a block's three 1x1 convolutions, cached per parameter version:
    (W_cls, b_cls, [W_kpt ; W_rep W_kpt], [b_kpt ; W_rep b_kpt + b_rep]).  reppts_out is a linear map of
    keypts_out (KP3:100-106,164-171), so the two are composed on the host in fp32 and keypoints and point set
    come out of ONE GEMM over the keypoint branch's activations -- the point set never sees a rounded
    keypoint tensor."""
    ps = (block.cls_out.weight, block.cls_out.bias, block.keypts_out.weight, block.keypts_out.bias,
          block.reppts_out.weight, block.reppts_out.bias)

    def build():
        with torch.no_grad():
            wc, bc, wk, bk, wr, br = (p.detach().float() for p in ps)
            wc, wk, wr = wc.flatten(1), wk.flatten(1), wr.flatten(1)
            w_kr = torch.cat([wk, wr @ wk], 0)
            b_kr = torch.cat([bk, wr @ bk + br], 0)
            return pack_weight(wc, split=True), bc.contiguous(), pack_weight(w_kr, split=True), b_kr.contiguous()
    return cached(ps, build, tag='pointwise')


def _pointwise_cls(block, cls_rows, n, h, w):
    """cls_out (NCHW fp32) from position-major bf16 activations: one tcgen05 GEMM (64-column tile)."""
    wc, bc, _, _ = _pointwise_weights(block)
    nc = block.cls_out.out_channels
    cls_out = torch.empty((n, nc, h, w), dtype=torch.float32, device=cls_rows.buf.device)
    pointwise_conv(cls_rows, wc, bc, [(cls_out, None, 0, nc)], h * w)
    return cls_out


def _pointwise_kpt(block, kpt_rows, n, h, w, kpt_prev=None, rep_prev=None):
    """keypoint and point-set outputs (NCHW fp32, residuals of the cascade added) from ONE tcgen05 GEMM over
    the keypoint branch's activations, instead of two cuDNN 1x1 convolutions + two adds."""
    _, _, w_kr, b_kr = _pointwise_weights(block)
    nk, nr = block.keypts_out.out_channels, block.reppts_out.out_channels
    dev = kpt_rows.buf.device
    kpt = torch.empty((n, nk, h, w), dtype=torch.float32, device=dev)
    rep = torch.empty((n, nr, h, w), dtype=torch.float32, device=dev)
    pointwise_conv(kpt_rows, w_kr, b_kr, [(kpt, kpt_prev, 0, nk), (rep, rep_prev, nk, nk + nr)], h * w)
    return kpt, rep


def _pointwise_heads(block, cls_rows, kpt_rows, n, h, w, kpt_prev=None, rep_prev=None):
    cls_out = _pointwise_cls(block, cls_rows, n, h, w)
    kpt, rep = _pointwise_kpt(block, kpt_rows, n, h, w, kpt_prev, rep_prev)
    return cls_out, kpt, rep


class _SideBranch(object):
    """Work forked from the current stream onto a cached side stream (a parallel branch of the captured CUDA
    graph).  Tensors the branch reads must stay referenced until ``join`` (the caching allocator orders reuse
    only against the stream a block was allocated on), so the caller hands them to ``keep``."""

    def __init__(self, device, index):
        self.main = torch.cuda.current_stream(device)
        self.side = _streams(device, index + 1)[index]
        self.side.wait_event(self.main.record_event())
        self.kept = []
        self._ctx = torch.cuda.stream(self.side)

    def __enter__(self):
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self._ctx.__exit__(*exc)
        self.done = self.side.record_event()
        return False

    def keep(self, *objs):
        self.kept.extend(objs)

    def join(self):
        self.main.wait_event(self.done)
        self.kept = []


class _PlainBlock(nn.Module):
    """Kp3RepBlock(deform_conv=False), KP3:98-106,173-177."""

    def __init__(self, cls_out, cin, feat, keypts_dim, reppts_dim):
        super().__init__()
        self.cls_conv = nn.Conv2d(cin, feat, 3, 1, 1)
        self.cls_out = nn.Conv2d(feat, cls_out, 1, 1, 0)
        self.keypts_conv = nn.Conv2d(cin, feat, 3, 1, 1)
        self.keypts_out = nn.Conv2d(feat, keypts_dim, 1, 1, 0)
        self.reppts_out = nn.Conv2d(keypts_dim, reppts_dim, 1, 1, 0)
        bias_cls = float(-math.log((1 - 0.01) / 0.01))            # bias_init_with_prob(0.01)
        _normal(self.cls_conv)
        _normal(self.keypts_conv)
        _normal(self.cls_out, bias=bias_cls)
        _normal(self.keypts_out)
        _normal(self.reppts_out)

    def forward(self, cls_feat, pts_feat, reppts_offset=None):
        cls_out = self.cls_out(F.relu(self.cls_conv(cls_feat)))
        keypts_out = self.keypts_out(F.relu(self.keypts_conv(pts_feat)))
        return cls_out, keypts_out, self.reppts_out(keypts_out)

    def forward_tc(self, cls_feat, pts_feat):
        """bf16 inference: the two 3x3 convolutions stay cuDNN; ReLU + layout change is one kernel each and the
        three 1x1 convolutions are two tensor-core GEMMs."""
        cls_out = self.forward_tc_cls(cls_feat)
        return (cls_out,) + self.forward_tc_kpt(pts_feat)

    def forward_tc_cls(self, cls_feat):
        n, _, h, w = cls_feat.shape
        # the convolution's bias is added by the tiling kernel (cuDNN would launch a separate add)
        cls_rows = nchw_to_tiled(_conv3x3_nhwc(cls_feat, self.cls_conv, False), relu=True, split=True,
                                 bias=self.cls_conv.bias)
        return _pointwise_cls(self, cls_rows, n, h, w)

    def forward_tc_kpt(self, pts_feat):
        n, _, h, w = pts_feat.shape
        kpt_rows = nchw_to_tiled(_conv3x3_nhwc(pts_feat, self.keypts_conv, False), relu=True, split=True,
                                 bias=self.keypts_conv.bias)
        return _pointwise_kpt(self, kpt_rows, n, h, w)

    # the same two branches with the 3x3 convolutions on this library's tensor-core kernel (split planes in)
    def forward_planes_cls(self, planes):
        n, _, h, w = planes.shape4
        cls_rows = nchw_to_tiled(conv_planes(planes, self.cls_conv.weight), relu=True, split=True,
                                 bias=self.cls_conv.bias)
        return _pointwise_cls(self, cls_rows, n, h, w)

    def forward_planes_kpt(self, planes):
        n, _, h, w = planes.shape4
        kpt_rows = nchw_to_tiled(conv_planes(planes, self.keypts_conv.weight), relu=True, split=True,
                                 bias=self.keypts_conv.bias)
        return _pointwise_kpt(self, kpt_rows, n, h, w)


class _DeformBlock(nn.Module):
    """Kp3RepBlock(deform_conv=True), KP3:35-96,126-171: six DeformConv on three point sets."""

    def __init__(self, cls_out, cin, feat, keypts_dim, reppts_dim, gradient_mul, deform_conv_cls=DeformConv):
        super().__init__()
        self.gradient_mul = gradient_mul
        self.concurrent_dcn = True          # forward_tc: the six DCNs of the stage on six streams
        self.grouped_dcn = True             # forward_tc: ... or, better, in one persistent grouped launch
        self.concurrent_training = os.environ.get('KGDET_TRAIN_STREAMS', '1') != '0'   # forward (autograd): six streams
        for k in _POINT_SETS:
            pad = (k - 1) // 2
            setattr(self, 'cls_dfmconv_%d' % k, deform_conv_cls(cin, feat, k, 1, pad))
            setattr(self, 'keypts_dfmconv_%d' % k, deform_conv_cls(cin, feat, k, 1, pad))
            base = np.arange(-pad, pad + 1).astype(np.float64)
            yx = np.stack([np.repeat(base, k), np.tile(base, k)], axis=1).reshape(-1)   # KP3:38-46
            self.register_buffer('_dcn_base_%d' % k, torch.tensor(yx, dtype=torch.float32).view(1, -1, 1, 1),
                                 persistent=False)
        self.cls_out = nn.Conv2d(feat * 3, cls_out, 1, 1, 0)
        self.keypts_out = nn.Conv2d(feat * 3, keypts_dim, 1, 1, 0)
        self.reppts_out = nn.Conv2d(keypts_dim, reppts_dim, 1, 1, 0)
        bias_cls = float(-math.log((1 - 0.01) / 0.01))
        for k in _POINT_SETS:
            _normal(getattr(self, 'cls_dfmconv_%d' % k))
            _normal(getattr(self, 'keypts_dfmconv_%d' % k))
        _normal(self.cls_out, bias=bias_cls)
        _normal(self.keypts_out)
        _normal(self.reppts_out)

    def forward_fused(self, cls_prep, pts_prep, reppts_offset):
        """Inference path on the prepared API: the NHWC copies of the two tower outputs are shared by all
        six DCNs (and by both stages), each offset tensor's sample plan is shared by the cls and keypoint
        branches, and every DCN applies its ReLU and writes its 256-channel slice of the concatenated
        tensor in its epilogue (no relu / cat kernels)."""
        n, c, h, w = cls_prep.shape4
        feat = self.cls_dfmconv_3.out_channels
        cls_cat = reppts_offset.new_empty((n, 3 * feat, h, w))
        kpt_cat = reppts_offset.new_empty((n, 3 * feat, h, w))
        lo = 0
        for i, k in enumerate(_POINT_SETS):
            n2 = 2 * k * k
            pts = reppts_offset[:, lo:lo + n2]
            pts = self.gradient_mul * pts + (1 - self.gradient_mul) * pts       # KP3:135-143, value-wise (no autograd here)
            dcn_offset = pts - getattr(self, '_dcn_base_%d' % k).to(reppts_offset.dtype)
            lo += n2
            plan = prepare_plan(dcn_offset, (n, c, h, w), feat, k, 1, (k - 1) // 2, 1, like_dtype=cls_cat.dtype)
            deform_conv_prepared(cls_prep, plan, getattr(self, 'cls_dfmconv_%d' % k).weight, cls_cat, i * feat, True)
            deform_conv_prepared(pts_prep, plan, getattr(self, 'keypts_dfmconv_%d' % k).weight, kpt_cat, i * feat,
                                 True)
        cls_out = self.cls_out(cls_cat)
        keypts_out = self.keypts_out(kpt_cat)
        return cls_out, keypts_out, self.reppts_out(keypts_out)

    def forward_tc(self, cls_prep, pts_prep, rep_prev, kpt_prev, branches=None):
        """bf16 inference, fully on this package's kernels: per point set one
        sample plan read straight from the channel slice of the previous stage's point tensor, two fused DCN
        launches that write ReLU-ed position-major bf16 rows, then two GEMMs whose epilogue adds bias and the
        cascade residuals (KP3:431-432,440-441) and writes NCHW fp32.  12 launches per stage."""
        n, c, h, w = cls_prep.shape4
        feat = self.cls_dfmconv_3.out_channels
        dev = rep_prev.device
        cls_rows = TiledRows(n * h * w, 3 * feat, True, dev)         # split precision: [hi | lo]
        kpt_rows = TiledRows(n * h * w, 3 * feat, True, dev)
        lo = 0
        jobs = []
        # the three sample plans are independent 8 us kernels: three parallel branches instead of a chain
        main = torch.cuda.current_stream(dev)
        plan_streams = [main] + _streams(dev, 2)
        fork = main.record_event()
        for i, k in enumerate(_POINT_SETS):
            st = plan_streams[i]
            if st is not main:
                st.wait_event(fork)
            with torch.cuda.stream(st):
                plan = prepare_plan_points(rep_prev, lo, (n, c, h, w), feat, k, 1, (k - 1) // 2, 1, precision='bf16',
                                           gradient_mul=self.gradient_mul)
            lo += 2 * k * k
            jobs.append((cls_prep, plan, getattr(self, 'cls_dfmconv_%d' % k).weight, cls_rows, i * feat))
            jobs.append((pts_prep, plan, getattr(self, 'keypts_dfmconv_%d' % k).weight, kpt_rows, i * feat))
        for st in plan_streams[1:]:
            main.wait_event(st.record_event())
        jobs.reverse()                                   # longest first (49, 49, 25, 25, 9, 9 points)
        if self.grouped_dcn:
            # ONE persistent launch for the six DCNs of the stage: 792 tiles handed out longest first to one CTA per
            # SM (kgdet_dcn_forward_prepared_group) -- no idle SMs (a single call has 132 tiles for 148 SMs), one
            # launch / TMEM allocation / barrier set-up instead of six
            deform_conv_prepared_group([job + (True,) for job in jobs])
        elif self.concurrent_dcn:
            # The six DCNs of a stage are independent and each fills only 132 of the 148 SMs (M = 16 800 ->
            # 132 tiles, one CTA per SM): issued on six streams, the block scheduler packs the tiles of all six
            # onto whatever SM is free (forks/joins become parallel branches of the captured CUDA graph).
            main = torch.cuda.current_stream(dev)
            fork = main.record_event()
            for job, st in zip(jobs, _streams(dev, len(jobs))):
                st.wait_event(fork)
                with torch.cuda.stream(st):
                    deform_conv_prepared(*job, True)
                    done = st.record_event()
                main.wait_event(done)
        else:
            for job in jobs:
                deform_conv_prepared(*job, True)
        if (self.concurrent_dcn or self.grouped_dcn) and branches is not None:
            # the 13-column cls GEMM is a parallel branch: it fills the SMs the keypoint GEMM's last wave leaves idle
            # and is only joined at the end of the head
            with _SideBranch(dev, len(jobs)) as br:
                cls_out = _pointwise_cls(self, cls_rows, n, h, w)
            br.keep(cls_rows)
            branches.append(br)
        else:
            cls_out = _pointwise_cls(self, cls_rows, n, h, w)
        kpt, rep = _pointwise_kpt(self, kpt_rows, n, h, w, kpt_prev, rep_prev)
        return cls_out, kpt, rep

    def forward(self, cls_feat, pts_feat, reppts_offset):
        cls_feats, kpt_feats = [], []
        lo = 0
        # Training at batch 2 (2 100 positions): every deformable convolution and each kernel of its backward is a
        # handful of CTAs.  The six calls of a stage are independent, so they are issued on six streams; autograd runs
        # every backward node on its forward's stream, so the captured training graph gets six parallel branches in
        # both directions.
        fork = (self.concurrent_training and cls_feat.is_cuda and torch.is_grad_enabled())
        if fork:
            dev = cls_feat.device
            main = torch.cuda.current_stream(dev)
            streams = _streams(dev, 6)
            start = main.record_event()
        jobs = []
        for i, k in enumerate(_POINT_SETS):
            n2 = 2 * k * k
            pts = reppts_offset[:, lo:lo + n2]                                     # KP3:131-133
            lo += n2
            # KP3:135-143 -- evaluated in inference too (the identity up to fp32 rounding), exactly as the reference
            pts = self.gradient_mul * pts + (1 - self.gradient_mul) * pts.detach()
            dcn_offset = pts - getattr(self, '_dcn_base_%d' % k).to(pts.dtype)
            if fork:
                jobs.append((k, dcn_offset))
                continue
            cls_feats.append(F.relu(getattr(self, 'cls_dfmconv_%d' % k)(cls_feat, dcn_offset)))
            kpt_feats.append(F.relu(getattr(self, 'keypts_dfmconv_%d' % k)(pts_feat, dcn_offset)))
        if fork:
            ready = main.record_event()                 # offsets computed
            outs = {}
            for j, (k, dcn_offset) in enumerate(reversed(jobs)):       # longest (49 points) first
                for b, (name, feat) in enumerate((('cls_dfmconv_%d', cls_feat), ('keypts_dfmconv_%d', pts_feat))):
                    st = streams[2 * j + b]
                    st.wait_event(ready)
                    with torch.cuda.stream(st):
                        o = F.relu(getattr(self, name % k)(feat, dcn_offset))
                        done = st.record_event()
                    o.record_stream(main)
                    outs[(k, b)] = (o, done)
            for k in _POINT_SETS:
                for b, lst in ((0, cls_feats), (1, kpt_feats)):
                    o, done = outs[(k, b)]
                    main.wait_event(done)
                    lst.append(o)
            del start
        if fork:
            st = streams[0]
            st.wait_event(main.record_event())
            with torch.cuda.stream(st):
                cls_out = self.cls_out(torch.cat(cls_feats, dim=1))                # KP3:152-156
                cls_done = st.record_event()
            keypts_out = self.keypts_out(torch.cat(kpt_feats, dim=1))              # KP3:164-170
            rep_out = self.reppts_out(keypts_out)
            main.wait_event(cls_done)
            cls_out.record_stream(main)
            return cls_out, keypts_out, rep_out
        cls_out = self.cls_out(torch.cat(cls_feats, dim=1))                        # KP3:152-156
        keypts_out = self.keypts_out(torch.cat(kpt_feats, dim=1))                  # KP3:164-170
        return cls_out, keypts_out, self.reppts_out(keypts_out)


class KGDetHead(nn.Module):
    """State-dict compatible restatement of RepPointsHeadKp3RepCas1AssignOnce (inference path)."""

    def __init__(self, num_classes=14, in_channels=256, feat_channels=256, point_feat_channels=256,
                 stacked_convs=3, num_keypts=294, gradient_mul=0.1, point_strides=(32,),
                 moment_mul=0.01, num_groups=32, deform_conv_cls=None, moment_fn=None, nms_flags_fn=None):
        # The three *_cls / *_fn arguments exist for bench.py's CPU reference arm, which injects
        # oracle-backed stand-ins; the defaults are the CUDA operators of this package.
        super().__init__()
        self._fused_inference = deform_conv_cls is None     # prepared API only with the CUDA operators
        self._tensor_core_heads = True                      # bf16 mode: 1x1 convolutions as tcgen05 GEMMs
        self._fused_decode = True                           # get_bboxes: decode kernels instead of PyTorch glue
        self._own_convs = True                              # bf16 inference: 3x3 convolutions on conv_umma.cu, not cuDNN
        self._fused_loss = True                             # loss(): assignment + losses as CUDA kernels (point_loss.cu)
        # training towers position-major (cuDNN on channels_last + this library's GroupNorm fwd / bwd kernels): measured
        # 2.97 -> 2.92 ms per step only -- the towers are not on the critical path of the six-arm graph -- so off by default
        self._nhwc_training = os.environ.get('KGDET_TRAIN_NHWC', '0') == '1'
        self.concurrent_branches = True                     # bf16 inference: cls / point branches on two streams
        # ... their convolutions of a layer in ONE launch (kgdet_conv_forward_pair): measured neutral (1.648 vs 1.652 ms --
        # the convolution's main loop is bound by L2 -> SM bytes, 2 100 clk per k-block, not by the prologue / epilogue
        # the pairing hides), so off by default
        self.paired_convs = os.environ.get('KGDET_PAIRED_CONVS', '0') == '1'
        deform_conv_cls = deform_conv_cls or DeformConv
        self._moment_fn = moment_fn or points2bbox_moment
        self._nms_flags_fn = nms_flags_fn or batched_nms_flags
        self._lim_cache = {}
        self.num_classes = num_classes
        self.cls_out_channels = num_classes - 1                                    # sigmoid cls, KP3:283-284
        self.num_keypts = num_keypts
        self.num_reppts = sum(k * k for k in _POINT_SETS)                          # KP3:257
        self.point_strides = list(point_strides)
        self.moment_mul = moment_mul
        self.moment_transfer = nn.Parameter(torch.zeros(2))                        # KP3:278-281
        self.cls_convs = nn.ModuleList()
        self.reg_convs = nn.ModuleList()
        for i in range(stacked_convs):
            chn = in_channels if i == 0 else feat_channels
            self.cls_convs.append(_ConvGNReLU(chn, feat_channels, num_groups))
            self.reg_convs.append(_ConvGNReLU(chn, feat_channels, num_groups))
        kd, rd = 2 * num_keypts, 2 * self.num_reppts
        self.kp_rep_block_1 = _PlainBlock(self.cls_out_channels, feat_channels, point_feat_channels, kd, rd)
        self.kp_rep_block_2 = _DeformBlock(self.cls_out_channels, feat_channels, point_feat_channels, kd, rd,
                                           gradient_mul, deform_conv_cls)
        self.kp_rep_block_3 = _DeformBlock(self.cls_out_channels, feat_channels, point_feat_channels, kd, rd,
                                           gradient_mul, deform_conv_cls)
        for m in list(self.cls_convs) + list(self.reg_convs):                      # init_weights, KP3:336-340
            _normal(m.conv)

    def points2bbox(self, pts, y_first=True):
        return self._moment_fn(pts, self.moment_transfer, self.moment_mul, y_first)

    def forward_single(self, x):
        fused = self._fused_inference and not torch.is_grad_enabled() and x.is_cuda
        if fused and self._tensor_core_heads and get_precision(x.dtype) == 'bf16' and x.dtype == torch.float32:
            # bf16 mode: everything except the eight plain 3x3 convolutions (cuDNN, channels_last) runs on this
            # package's kernels.  Towers: position-major activations end to end,
            # GroupNorm + ReLU fused -- no layout transposes, no separate ReLU kernels.
            feat = self.kp_rep_block_2.cls_dfmconv_3.out_channels
            own = self._own_convs and all(conv_supported(m.conv.in_channels, m.conv.out_channels, 3)
                                          for m in list(self.cls_convs) + list(self.reg_convs)) and \
                x.shape[2] * x.shape[3] <= 1600
            # own: every 3x3 convolution on this library's fp32-grade tensor-core kernel (conv_umma.cu), activations
            # travel as split planes (the last tower layer's hi planes ARE the deformable stage's prepared input);
            # otherwise cuDNN in channels_last (TF32 unless the caller switched it off)
            cls_feat = pts_feat = split_planes(x) if own else to_channels_last(x)

            def cls_branch_own(p):
                for m in self.cls_convs:
                    p = groupnorm_relu_planes(conv_planes(p, m.conv.weight), m.gn)
                return self.kp_rep_block_1.forward_planes_cls(p), p.as_prepared_input(feat)

            def pts_branch_own(p):
                for m in self.reg_convs:
                    p = groupnorm_relu_planes(conv_planes(p, m.conv.weight), m.gn)
                kpt1, rep1 = self.kp_rep_block_1.forward_planes_kpt(p)
                return kpt1, rep1, self.points2bbox(rep1), p.as_prepared_input(feat)

            def cls_branch(cls_feat):
                for m in self.cls_convs:
                    cls_feat = groupnorm_relu_nhwc(_conv3x3_nhwc(cls_feat, m.conv), m.gn)
                return self.kp_rep_block_1.forward_tc_cls(cls_feat), prepare_input(cls_feat, feat, precision='bf16')

            def pts_branch(pts_feat):
                for m in self.reg_convs:
                    pts_feat = groupnorm_relu_nhwc(_conv3x3_nhwc(pts_feat, m.conv), m.gn)
                kpt1, rep1 = self.kp_rep_block_1.forward_tc_kpt(pts_feat)
                return kpt1, rep1, self.points2bbox(rep1), prepare_input(pts_feat, feat, precision='bf16')

            # packed 1x1 weights are cached per block and shared by its cls and keypoint GEMMs, which run on different
            # streams below: build (or refresh) them here, on the stream everything forks from
            for blk in (self.kp_rep_block_1, self.kp_rep_block_2, self.kp_rep_block_3):
                _pointwise_weights(blk)
            branches = [] if self.concurrent_branches else None
            if own and self.paired_convs and self.concurrent_branches:
                # The two towers' convolutions of a layer in ONE launch (kgdet_conv_forward_pair): on two streams they
                # only take turns -- 132 one-per-SM CTAs each -- and every CTA pays prologue and epilogue around a
                # single tile; paired, a CTA runs the classification tile and then the point tile into the other half
                # of TMEM, the first epilogue under the second main loop.  The two GroupNorms are parallel branches.
                p_cls = p_pts = cls_feat
                for mc, mp in zip(self.cls_convs, self.reg_convs):
                    y_c, y_p = conv_planes_pair(p_cls, mc.conv.weight, p_pts, mp.conv.weight)
                    with _SideBranch(x.device, 7) as br:
                        p_cls = groupnorm_relu_planes(y_c, mc.gn)
                    br.keep(y_c)
                    p_pts = groupnorm_relu_planes(y_p, mp.gn)
                    br.join()
                b1 = self.kp_rep_block_1
                y_c, y_p = conv_planes_pair(p_cls, b1.cls_conv.weight, p_pts, b1.keypts_conv.weight)
                n1, _, h1, w1 = p_cls.shape4
                with _SideBranch(x.device, 7) as br:
                    cls1 = _pointwise_cls(b1, nchw_to_tiled(y_c, relu=True, split=True, bias=b1.cls_conv.bias), n1, h1, w1)
                br.keep(y_c)
                kpt1, rep1 = _pointwise_kpt(b1, nchw_to_tiled(y_p, relu=True, split=True, bias=b1.keypts_conv.bias),
                                            n1, h1, w1)
                bbox1 = self.points2bbox(rep1)
                cls_prep, pts_prep = p_cls.as_prepared_input(feat), p_pts.as_prepared_input(feat)
                br.join()
            elif self.concurrent_branches:
                # The classification and point branches are independent up to the first deformable stage, and a
                # 3x3 convolution on a 25x42 map is 66 CTAs of cuDNN's 256-row tile -- under half of the SMs:
                # the two towers run as parallel branches (two streams / two arms of the captured graph).
                with _SideBranch(x.device, 7) as br:
                    cls1, cls_prep = (cls_branch_own if own else cls_branch)(cls_feat)
                br.keep(cls_feat)
                kpt1, rep1, bbox1, pts_prep = (pts_branch_own if own else pts_branch)(pts_feat)
                br.join()
            else:
                cls1, cls_prep = (cls_branch_own if own else cls_branch)(cls_feat)
                kpt1, rep1, bbox1, pts_prep = (pts_branch_own if own else pts_branch)(pts_feat)
            cls2, kpt2, rep2 = self.kp_rep_block_2.forward_tc(cls_prep, pts_prep, rep1, kpt1, branches)
            bbox2 = self.points2bbox(rep2)
            cls3, kpt3, rep3 = self.kp_rep_block_3.forward_tc(cls_prep, pts_prep, rep2, kpt2, branches)
            bbox3 = self.points2bbox(rep3)
            for br in branches or ():
                br.join()
            return cls1, cls2, cls3, kpt1, kpt2, kpt3, bbox1, bbox2, bbox3
        if (fused and self._tensor_core_heads and self._own_convs and get_precision(x.dtype) == 'tf32x3'
                and x.dtype == torch.float32 and x.shape[2] * x.shape[3] <= 1600
                and all(conv_supported(m.conv.in_channels, m.conv.out_channels, 3)
                        for m in list(self.cls_convs) + list(self.reg_convs))):
            return self._forward_single_fp32_grade(x)
        cls_feat = pts_feat = x
        if x.is_cuda and torch.is_grad_enabled() and self.kp_rep_block_2.concurrent_training:
            # training: the classification tower + stage-1 classifier run on a side stream next to the point branch
            # (autograd replays each backward node on its forward's stream: two parallel arms in both directions)
            main = torch.cuda.current_stream(x.device)
            side = _streams(x.device, 7)[6]
            side.wait_event(main.record_event())
            nhwc = self._nhwc_training and x.dtype == torch.float32
            if nhwc:
                # position-major towers: cuDNN runs fprop / dgrad / wgrad on channels_last tensors without layout
                # transposes, GroupNorm + ReLU forward and backward are one kernel each of this library
                cls_feat = pts_feat = x.contiguous(memory_format=torch.channels_last)

            def tower(convs, t):
                for m in convs:
                    if nhwc:
                        wcl = m.conv.weight.contiguous(memory_format=torch.channels_last)
                        t = groupnorm_relu_nhwc_autograd(F.conv2d(t, wcl, None, 1, 1), m.gn)
                    else:
                        t = m(t)
                return t

            with torch.cuda.stream(side):
                cls_feat = tower(self.cls_convs, cls_feat)
                b1 = self.kp_rep_block_1
                cls1 = b1.cls_out(F.relu(b1.cls_conv(cls_feat)))
                if nhwc:
                    cls_feat = cls_feat.contiguous()          # the deformable stages read NCHW
                    cls1 = cls1.contiguous()
                cls_done = side.record_event()
            pts_feat = tower(self.reg_convs, pts_feat)
            kpt1 = b1.keypts_out(F.relu(b1.keypts_conv(pts_feat)))
            rep1 = b1.reppts_out(kpt1)
            if nhwc:
                pts_feat, kpt1, rep1 = pts_feat.contiguous(), kpt1.contiguous(), rep1.contiguous()
            main.wait_event(cls_done)
            cls_feat.record_stream(main)
            cls1.record_stream(main)
        else:
            for m in self.cls_convs:
                cls_feat = m(cls_feat)
            for m in self.reg_convs:
                pts_feat = m(pts_feat)
            cls1, kpt1, rep1 = self.kp_rep_block_1(cls_feat, pts_feat)
        bbox1 = self.points2bbox(rep1)
        if fused:
            feat = self.kp_rep_block_2.cls_dfmconv_3.out_channels
            cls_prep = prepare_input(cls_feat, feat)
            pts_prep = prepare_input(pts_feat, feat)
            cls2, kpt2, rep2 = self.kp_rep_block_2.forward_fused(cls_prep, pts_prep, rep1)
            kpt2 = kpt2 + kpt1
            rep2 = rep2 + rep1
            bbox2 = self.points2bbox(rep2)
            cls3, kpt3, rep3 = self.kp_rep_block_3.forward_fused(cls_prep, pts_prep, rep2)
            kpt3 = kpt3 + kpt2
            rep3 = rep3 + rep2
            bbox3 = self.points2bbox(rep3)
            return cls1, cls2, cls3, kpt1, kpt2, kpt3, bbox1, bbox2, bbox3
        cls2, kpt2, rep2 = self.kp_rep_block_2(cls_feat, pts_feat, rep1)
        kpt2 = kpt2 + kpt1.detach()                                                # KP3:431-432
        rep2 = rep2 + rep1.detach()
        bbox2 = self.points2bbox(rep2)
        cls3, kpt3, rep3 = self.kp_rep_block_3(cls_feat, pts_feat, rep2)
        kpt3 = kpt3 + kpt2.detach()                                                # KP3:440-441
        rep3 = rep3 + rep2.detach()
        bbox3 = self.points2bbox(rep3)
        return cls1, cls2, cls3, kpt1, kpt2, kpt3, bbox1, bbox2, bbox3

    def _forward_single_fp32_grade(self, x):
        """fp32 inference (the reference's precision; DCN mode 'tf32x3', the default for fp32 tensors) with every
        contraction on this library's tensor-core kernels at fp32-grade accuracy: 3x3 convolutions and 1x1 GEMMs at
        split precision (bf16x3), GroupNorm fused, deformable convolutions in the tf32x3 mode with accumulator
        promotion (NCHW fp32 outputs, then one re-tiling pass per branch for the GEMMs).  cuDNN's fp32 path takes
        1.5 ms per 3x3 convolution at [16, 256, 25, 42] (TF32 off); this path 56 us."""
        feat = self.kp_rep_block_2.cls_dfmconv_3.out_channels
        n, _, h, w = x.shape
        planes = split_planes(x)

        def tower(convs, p):
            dense = None
            for i, m in enumerate(convs):
                y = conv_planes(p, m.conv.weight)
                if i + 1 < len(convs):
                    p = groupnorm_relu_planes(y, m.gn)
                else:
                    p, dense = groupnorm_relu_planes(y, m.gn, also_dense=True)
            return p, dense

        for blk in (self.kp_rep_block_1, self.kp_rep_block_2, self.kp_rep_block_3):
            _pointwise_weights(blk)
        cls_p, cls_dense = tower(self.cls_convs, planes)
        pts_p, pts_dense = tower(self.reg_convs, planes)
        cls1 = self.kp_rep_block_1.forward_planes_cls(cls_p)
        kpt1, rep1 = self.kp_rep_block_1.forward_planes_kpt(pts_p)
        bbox1 = self.points2bbox(rep1)
        cls_prep = prepare_input(cls_dense, feat, precision='tf32x3')
        pts_prep = prepare_input(pts_dense, feat, precision='tf32x3')
        outs = [cls1, kpt1, bbox1]
        rep_prev, kpt_prev = rep1, kpt1
        for blk in (self.kp_rep_block_2, self.kp_rep_block_3):
            cls_cat = x.new_empty((n, 3 * feat, h, w))
            kpt_cat = x.new_empty((n, 3 * feat, h, w))
            lo = 0
            for i, k in enumerate(_POINT_SETS):
                plan = prepare_plan_points(rep_prev, lo, (n, cls_dense.shape[1], h, w), feat, k, 1, (k - 1) // 2, 1,
                                           precision='tf32x3', gradient_mul=blk.gradient_mul)      # KP3:131-143
                lo += 2 * k * k
                deform_conv_prepared(cls_prep, plan, getattr(blk, 'cls_dfmconv_%d' % k).weight, cls_cat, i * feat, True)
                deform_conv_prepared(pts_prep, plan, getattr(blk, 'keypts_dfmconv_%d' % k).weight, kpt_cat, i * feat, True)
            cls_k, kpt_k, rep_k = _pointwise_heads(blk, nchw_to_tiled(cls_cat, split=True), nchw_to_tiled(kpt_cat, split=True),
                                                   n, h, w, kpt_prev, rep_prev)                    # KP3:431-432,440-441
            outs += [cls_k, kpt_k, self.points2bbox(rep_k)]
            rep_prev, kpt_prev = rep_k, kpt_k
        cls1, kpt1, bbox1, cls2, kpt2, bbox2, cls3, kpt3, bbox3 = outs
        return cls1, cls2, cls3, kpt1, kpt2, kpt3, bbox1, bbox2, bbox3

    def forward(self, feats):
        outs = [self.forward_single(x) for x in feats]
        return tuple(map(list, zip(*outs)))                                        # multi_apply layout

    def loss(self, outs, gt_bboxes, gt_labels, gt_keypoints, gt_valid, assigner_scale=4, pos_num=25,
             point_base_scale=4):
        """The nine training losses of the reference head (KP3:670-768) for the single KGDet level, from the
        9-tuple of `forward_single` and PADDED ground truth (`targets.pad_ground_truth`): target assignment
        (PointAssigner + point_target_kp, "assign once": the same targets serve all three stages) and the loss
        reductions run batched on the device without host synchronisation (kgdet_b200/targets.py), so the whole
        training step can be replayed as a CUDA graph.  Loss weights are the reference configs' (0.5, 0.5, 1.0)."""
        from . import targets as T
        h, w = outs[0].shape[-2:]
        stride = self.point_strides[0]
        if self._fused_loss and outs[0].is_cuda and h * w <= 4096 and len(self.point_strides) == 1:
            # assignment + the nine losses as three kernels of the library (csrc/point_loss.cu)
            from .ops.point_loss import kgdet_point_losses
            return kgdet_point_losses(outs, gt_bboxes, gt_labels, gt_keypoints, gt_valid, stride, assigner_scale, pos_num,
                                      point_base_scale)
        key = (h, w, stride, outs[0].device)
        pts = self._lim_cache.get(('points',) + key)
        if pts is None:
            pts = T.grid_points(h, w, stride, outs[0].device)
            self._lim_cache[('points',) + key] = pts
        tg = T.point_targets(pts, gt_bboxes, gt_labels, gt_keypoints, gt_valid, assigner_scale, pos_num)
        return T.kgdet_losses(outs, pts, stride, tg, point_base_scale)

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def get_bboxes(self, cls_scores, keypts_preds, bbox_preds, img_shapes, score_thr=0.05, iou_thr=0.5,
                   nms_pre=1000, max_per_img=100, score_override=None, scale_factors=None, return_kept=False):
        """Batched, sync-free form of KP3:770-914 + multiclass_nms_kp.

        return_kept: also return the number of boxes the NMS kept per image, [B], and the flat (class, candidate)
        index of every result row, [B, k] (the reference sorts by score only when more than max_per_img boxes were
        kept, bbox_nms_kp.py:64-70; otherwise its rows stay in class / candidate order -- kgdet_b200.adopt restores
        that ordering from these two).

        scale_factors: optional per-image floats -- the reference's `rescale=True` (KP3:892-898): boxes and keypoint
        coordinates are divided by the image's scale factor BEFORE the NMS, as the reference does.

        cls_scores / keypts_preds / bbox_preds: per-level lists of the stage-3 outputs
        ([B,13,H,W], [B,588,H,W], [B,4,H,W]).  img_shapes: list of (h, w) per image.
        score_override: optional per-level [B,13,H,W] sigmoid scores used instead of
        sigmoid(cls_scores) (benchmarks with random-init weights: every real score is ~0.01).
        Returns dets [B, max_per_img, 5], labels [B, max_per_img] (-1 = empty slot),
        kpts [B, max_per_img, num_keypts*3], all sorted by score; no host sync.
        """
        B = cls_scores[0].shape[0]
        dev = cls_scores[0].device
        key = (tuple(tuple(s[:2]) for s in img_shapes), str(dev))
        lim = self._lim_cache.get(key)
        if lim is None:       # built once per (shapes, device): keeps H2D copies out of graph capture
            lim = torch.tensor([[s[1], s[0], s[1], s[0]] for s in img_shapes], dtype=torch.float32, device=dev)
            self._lim_cache[key] = lim
        sf = None
        if scale_factors is not None:
            skey = ('scale', tuple(float(f) for f in scale_factors), str(dev))
            sf = self._lim_cache.get(skey)
            if sf is None:
                sf = torch.tensor([float(f) for f in scale_factors], dtype=torch.float32, device=dev)
                self._lim_cache[skey] = sf
        if (self._fused_decode and len(cls_scores) == 1 and cls_scores[0].is_cuda
                and self._nms_flags_fn is batched_nms_flags and cls_scores[0].shape[-2] * cls_scores[0].shape[-1] <= 16384):
            # one head level on the GPU: three decode kernels + the batched NMS + one top-k (section 8(f) rank 1)
            cs, kp, bp = cls_scores[0], keypts_preds[0], bbox_preds[0]
            H, W = cs.shape[-2:]
            src = (cs if score_override is None else score_override[0]).float().contiguous()
            sig = score_override is None
            n = min(nms_pre, H * W) if nms_pre > 0 else H * W
            wh = lim[:, :2].contiguous() if lim.shape[1] == 4 else lim
            stride = self.point_strides[0]
            order = bbox_select(src, sig, n)
            boxes, dets = bbox_decode(src, sig, bp.float().contiguous(), order, wh, stride)
            C = dets.shape[1]
            if sf is not None:                                                          # KP3:892-893
                boxes = boxes / sf.view(B, 1, 1)
                dets[..., :4] = dets[..., :4] / sf.view(B, 1, 1, 1)
            flags = self._nms_flags_fn(dets.view(-1, 5), None, n, iou_thr, score_thr=score_thr)
            if C * n <= 16384:
                # survivors compacted + sorted inside one CTA per image (bbox_nms_kp.py:64-70)
                top_s, top_i = topk_flagged(dets.view(B, C * n, 5), flags.view(B, C * n), min(max_per_img, C * n))
            else:
                masked = torch.where(flags.view(B, C * n).bool(), dets[..., 4].reshape(B, C * n),
                                     dets.new_full((), -1.0))
                top_s, top_i = masked.topk(min(max_per_img, C * n), dim=1)
            res = bbox_finalize(boxes, kp.float().contiguous(), order, top_i, top_s, wh, stride, (H, W))
            if sf is not None:                                                          # KP3:894-896
                kv = res[2].view(B, res[2].shape[1], -1, 3)
                kv[..., :2] = kv[..., :2] / sf.view(B, 1, 1, 1)
            return res + (flags.view(B, -1).sum(1), top_i) if return_kept else res
        boxes_l, scores_l, kpts_l = [], [], []
        for lvl, (cs, kp, bp) in enumerate(zip(cls_scores, keypts_preds, bbox_preds)):
            stride = self.point_strides[lvl]
            H, W = cs.shape[-2:]
            scores = cs.permute(0, 2, 3, 1).reshape(B, H * W, -1).float()
            scores = scores.sigmoid() if score_override is None else \
                score_override[lvl].permute(0, 2, 3, 1).reshape(B, H * W, -1).float()
            bbox = bp.permute(0, 2, 3, 1).reshape(B, H * W, 4).float()
            # points2kpt (KP3:393-410): (y, x) pairs -> (x, y); visibility padded with ones (KP3:856-861)
            kpt = kp.permute(0, 2, 3, 1).reshape(B, H * W, self.num_keypts, 2).float().flip(-1)
            ys, xs = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32) * stride,
                                    torch.arange(W, device=dev, dtype=torch.float32) * stride, indexing='ij')
            centers = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1)                 # point_generator.py:14-23
            if nms_pre > 0 and H * W > nms_pre:                                        # KP3:863-874
                _, topk = scores.max(dim=2)[0].topk(nms_pre, dim=1)
                scores = scores.gather(1, topk[..., None].expand(-1, -1, scores.shape[2]))
                bbox = bbox.gather(1, topk[..., None].expand(-1, -1, 4))
                kpt = kpt.gather(1, topk[..., None, None].expand(-1, -1, self.num_keypts, 2))
                ctr = centers[topk]
            else:
                ctr = centers[None].expand(B, -1, -1)
            boxes = bbox * stride + torch.cat([ctr, ctr], -1)                           # KP3:875-877
            boxes = torch.min(boxes.clamp(min=0), lim[:, None, :])                      # KP3:882-886
            kpt = kpt * stride + ctr[:, :, None, :]                                     # KP3:878-880
            kpt = torch.min(kpt.clamp(min=0), lim[:, None, None, :2])                   # KP3:887-888
            boxes_l.append(boxes)
            scores_l.append(scores)
            kpts_l.append(kpt)
        boxes = torch.cat(boxes_l, 1)
        scores = torch.cat(scores_l, 1)
        kpts = torch.cat(kpts_l, 1)
        if sf is not None:                                                              # KP3:892-896
            boxes = boxes / sf.view(B, 1, 1)
            kpts = kpts / sf.view(B, 1, 1, 1)
        n, C = scores.shape[1], scores.shape[2]

        # one dense segment of n rows per (image, class), in the order of the reference's per-class loop
        # (bbox_nms_kp.py:38-52); its `scores > score_thr` filter (:39) is applied inside the NMS op, so
        # every shape is static: no nonzero(), no host sync, CUDA-graph capturable
        sc_t = scores.transpose(1, 2).contiguous()                                      # [B, C, n]
        dets = torch.cat([boxes[:, None].expand(B, C, n, 4), sc_t[..., None]], -1).reshape(B * C * n, 5)
        flags = self._nms_flags_fn(dets, None, n, iou_thr, score_thr=score_thr)
        masked = torch.where(flags.view(B, C * n).bool(), sc_t.reshape(B, C * n), sc_t.new_full((), -1.0))
        k = min(max_per_img, C * n)
        top_s, top_i = masked.topk(k, dim=1)                                            # bbox_nms_kp.py:64-70
        valid = top_s > 0
        cls_i = top_i // n
        row_i = top_i % n
        out_boxes = boxes.gather(1, row_i[..., None].expand(-1, -1, 4))
        out_dets = torch.cat([out_boxes, top_s[..., None]], -1) * valid[..., None]
        out_labels = torch.where(valid, cls_i, torch.full_like(cls_i, -1))
        out_kpts = kpts.gather(1, row_i[..., None, None].expand(-1, -1, self.num_keypts, 2))
        vis = torch.ones_like(out_kpts[..., :1])
        out_kpts = (torch.cat([out_kpts, vis], -1) * valid[..., None, None]).reshape(B, k, -1)
        if return_kept:
            return out_dets, out_labels, out_kpts, flags.view(B, -1).sum(1), top_i
        return out_dets, out_labels, out_kpts


class RepPointsKpHead(nn.Module):
    """State-dict compatible restatement of the two RepPoints-Kp baselines of the reference
    (`RepPointsHeadKpParallel`, reppoints_head_kp_parallel.py:17-341, 6 806 475 parameters, and
    `RepPointsHeadKpSerial`, reppoints_head_kp_serial.py:17-341, 5 638 523 parameters): 5 FPN levels,
    9 learned points, one shared 3x3 offset tensor per level feeding 3 (parallel) or 2 (serial)
    deformable convolutions.  Inference runs them through the prepared API: one NHWC copy per tower
    output, ONE sample plan per level, ReLU fused in the DCN epilogue."""

    def __init__(self, variant='parallel', num_classes=14, in_channels=256, feat_channels=256,
                 point_feat_channels=256, stacked_convs=3, num_reppts=9, num_keypts=294, gradient_mul=0.1,
                 point_strides=(8, 16, 32, 64, 128), moment_mul=0.01, num_groups=32, deform_conv_cls=None,
                 moment_fn=None, nms_flags_fn=None):
        super().__init__()
        assert variant in ('parallel', 'serial')
        self.variant = variant
        self._fused_inference = deform_conv_cls is None
        self._fused_decode = True           # get_bboxes: per-level select / decode kernels instead of PyTorch glue
        self._nms_flags_fn = nms_flags_fn or batched_nms_flags
        self._lim_cache = {}
        deform_conv_cls = deform_conv_cls or DeformConv
        self._moment_fn = moment_fn or points2bbox_moment
        self.cls_out_channels = num_classes - 1
        self.num_keypts, self.num_reppts = num_keypts, num_reppts
        self.gradient_mul = gradient_mul
        self.point_strides = list(point_strides)
        self.moment_mul = moment_mul
        self.moment_transfer = nn.Parameter(torch.zeros(2))
        k = int(round(num_reppts ** 0.5))
        assert k * k == num_reppts and k % 2 == 1
        self.dcn_kernel, self.dcn_pad = k, (k - 1) // 2
        base = np.arange(-self.dcn_pad, self.dcn_pad + 1).astype(np.float64)
        yx = np.stack([np.repeat(base, k), np.tile(base, k)], axis=1).reshape(-1)      # PAR:108-113
        self.register_buffer('_dcn_base', torch.tensor(yx, dtype=torch.float32).view(1, -1, 1, 1), persistent=False)
        self.cls_convs = nn.ModuleList()
        self.reg_convs = nn.ModuleList()
        for i in range(stacked_convs):
            chn = in_channels if i == 0 else feat_channels
            self.cls_convs.append(_ConvGNReLU(chn, feat_channels, num_groups))
            self.reg_convs.append(_ConvGNReLU(chn, feat_channels, num_groups))
        kd, rd, pf = 2 * num_keypts, 2 * num_reppts, point_feat_channels
        self.grouped_dcn = True             # bf16 inference: the DCNs of all levels in grouped persistent launches
        self.nhwc_towers = True             # ... with position-major towers / 1x1 GEMMs around them (_forward_nhwc)
        self.concurrent_levels = True       # ... and the FPN levels as parallel branches around the grouped launches
        self.cls_refine_dfmconv = deform_conv_cls(feat_channels, pf, k, 1, self.dcn_pad)
        self.cls_refine_out = nn.Conv2d(pf, self.cls_out_channels, 1, 1, 0)
        self.keypts_init_conv = nn.Conv2d(feat_channels, pf, 3, 1, 1)
        self.keypts_init_out = nn.Conv2d(pf, kd, 1, 1, 0)
        self.keypts_refine_dfmconv = deform_conv_cls(feat_channels, pf, k, 1, self.dcn_pad)
        self.keypts_refine_out = nn.Conv2d(pf, kd, 1, 1, 0)
        if variant == 'parallel':                                                 # PAR:147-171
            self.reppts_init_conv = nn.Conv2d(feat_channels, pf, 3, 1, 1)
            self.reppts_init_out = nn.Conv2d(pf, rd, 1, 1, 0)
            self.reppts_refine_dfmconv = deform_conv_cls(feat_channels, pf, k, 1, self.dcn_pad)
            self.reppts_refine_out = nn.Conv2d(pf, rd, 1, 1, 0)
        else:                                                                     # SER:156-168
            self.reppts_init_out = nn.Conv2d(kd, rd, 1, 1, 0)
            self.reppts_refine_out = nn.Conv2d(kd, rd, 1, 1, 0)
        bias_cls = float(-math.log((1 - 0.01) / 0.01))
        for m in list(self.cls_convs) + list(self.reg_convs):
            _normal(m.conv)
        for n, m in self.named_children():
            if n.endswith(('_conv', '_out', '_dfmconv')):
                _normal(m, bias=bias_cls if n == 'cls_refine_out' else 0.0)

    def points2bbox(self, pts, y_first=True):
        return self._moment_fn(pts, self.moment_transfer, self.moment_mul, y_first)

    def forward_single(self, x):
        cls_feat = pts_feat = x
        for m in self.cls_convs:
            cls_feat = m(cls_feat)
        for m in self.reg_convs:
            pts_feat = m(pts_feat)
        kpt_init = self.keypts_init_out(F.relu(self.keypts_init_conv(pts_feat)))
        if self.variant == 'parallel':
            rep_init = self.reppts_init_out(F.relu(self.reppts_init_conv(pts_feat)))     # PAR:314-315
        else:
            rep_init = self.reppts_init_out(kpt_init)                                    # SER:314
        base = self._dcn_base.to(x.dtype)
        if self._fused_inference and not torch.is_grad_enabled() and x.is_cuda:
            n, c, h, w = cls_feat.shape
            pf = self.cls_refine_dfmconv.out_channels
            pts = self.gradient_mul * rep_init + (1 - self.gradient_mul) * rep_init       # PAR:322-325, value-wise
            plan = prepare_plan(pts - base, (n, c, h, w), pf, self.dcn_kernel, 1, self.dcn_pad, 1,
                                like_dtype=x.dtype)
            cls_prep = prepare_input(cls_feat, pf)
            pts_prep = prepare_input(pts_feat, pf)
            cls_out = self.cls_refine_out(deform_conv_prepared(cls_prep, plan, self.cls_refine_dfmconv.weight,
                                                               relu=True))
            kpt_ref = self.keypts_refine_out(deform_conv_prepared(pts_prep, plan,
                                                                  self.keypts_refine_dfmconv.weight, relu=True))
            if self.variant == 'parallel':
                rep_ref = self.reppts_refine_out(deform_conv_prepared(pts_prep, plan,
                                                                      self.reppts_refine_dfmconv.weight, relu=True))
            else:
                rep_ref = self.reppts_refine_out(kpt_ref)
            return cls_out, kpt_init, kpt_ref + kpt_init, rep_init, rep_ref + rep_init
        # PAR:322-325 -- evaluated in inference too (the identity up to fp32 rounding), exactly as the reference
        pts = self.gradient_mul * rep_init + (1 - self.gradient_mul) * rep_init.detach()
        dcn_offset = pts - base
        cls_out = self.cls_refine_out(F.relu(self.cls_refine_dfmconv(cls_feat, dcn_offset)))
        kpt_ref = self.keypts_refine_out(F.relu(self.keypts_refine_dfmconv(pts_feat, dcn_offset)))
        if self.variant == 'parallel':
            rep_ref = self.reppts_refine_out(F.relu(self.reppts_refine_dfmconv(pts_feat, dcn_offset)))
        else:
            rep_ref = self.reppts_refine_out(kpt_ref)                                    # SER:330
        kpt_ref = kpt_ref + kpt_init.detach()                                            # PAR:337-338
        rep_ref = rep_ref + rep_init.detach()
        return cls_out, kpt_init, kpt_ref, rep_init, rep_ref

    def forward(self, feats):
        if (self._fused_inference and self.grouped_dcn and not torch.is_grad_enabled() and feats[0].is_cuda
                and get_precision(feats[0].dtype) == 'bf16' and feats[0].dtype == torch.float32):
            if self.nhwc_towers and self._nhwc_supported():
                return self._forward_nhwc(feats)
            return self._forward_grouped(feats)
        outs = [self.forward_single(x) for x in feats]
        return tuple(map(list, zip(*outs)))

    def _nhwc_supported(self):
        gns = [m.gn for m in list(self.cls_convs) + list(self.reg_convs)]
        pf = self.cls_refine_dfmconv.out_channels
        return (all(g.num_channels % 128 == 0 and g.num_channels // g.num_groups in (4, 8, 16, 32) for g in gns)
                and pf % 64 == 0 and self.cls_refine_dfmconv.in_channels % 64 == 0)

    def _head_weights(self):
        """Packed operands of the 1x1 convolutions (cached per parameter version).  Serial variant: the point-set
        convolutions are linear maps of the keypoint outputs (SER:314,330), composed on the host in fp32 so keypoints
        and point set come out of one GEMM."""
        names = ['cls_refine_out', 'keypts_init_out', 'keypts_refine_out', 'reppts_init_out', 'reppts_refine_out']
        ps = tuple(p for n in names for p in (getattr(self, n).weight, getattr(self, n).bias))

        def build():
            with torch.no_grad():
                w = {n: getattr(self, n).weight.detach().float().flatten(1) for n in names}
                b = {n: getattr(self, n).bias.detach().float() for n in names}
                out = {'cls': (pack_weight(w['cls_refine_out'], split=True), b['cls_refine_out'].contiguous())}
                for st in ('init', 'refine'):
                    wk, bk = w['keypts_%s_out' % st], b['keypts_%s_out' % st]
                    wr, br = w['reppts_%s_out' % st], b['reppts_%s_out' % st]
                    if self.variant == 'parallel':
                        out['kpt_' + st] = (pack_weight(wk, split=True), bk.contiguous())
                        out['rep_' + st] = (pack_weight(wr, split=True), br.contiguous())
                    else:
                        out['kpt_rep_' + st] = (pack_weight(torch.cat([wk, wr @ wk], 0), split=True),
                                                torch.cat([bk, wr @ bk + br], 0).contiguous())
                return out
        return cached(ps, build, tag='reppoints_pointwise')

    def _forward_nhwc(self, feats):
        """bf16 inference with position-major activations end to end (BASELINE.json configs[3] on the fast path):

        * towers: cuDNN 3x3 convolutions in channels_last (no layout transposes) + this library's GroupNorm + ReLU
          kernel (resident for maps of <= 1600 positions, two streaming passes for P3 / P4); the last tower layer's
          normalisation writes the deformable stage's prepared bf16 planes directly;
        * the init branches' ReLU + bias + re-tiling is one kernel, every 1x1 convolution a tcgen05 GEMM of this
          library at split ("bf16x3") precision whose epilogue adds the refine stage's residual (PAR:337-338);
        * one sample plan per level straight from the point tensor (PAR:322-325), the deformable convolutions of ALL
          levels in grouped persistent launches writing ReLU-ed bf16 rows for those GEMMs.
        Same values as `forward_single` up to the arithmetic of the 1x1 convolutions (fp32-grade here, cuDNN there)."""
        pf = self.cls_refine_dfmconv.out_channels
        kd, rd, nc = 2 * self.num_keypts, 2 * self.num_reppts, self.cls_out_channels
        W = self._head_weights()
        par = self.variant == 'parallel'
        last = len(self.cls_convs) - 1
        per_level, jobs = [], []
        # levels are independent up to the grouped launches: level i > 0 runs on its own stream (a parallel branch of the
        # captured graph) next to P3, whose kernels are the only ones that fill the machine
        dev0 = feats[0].device
        main = torch.cuda.current_stream(dev0)
        lvl_streams = [main] + (_streams(dev0, 8 + len(feats))[9:8 + len(feats)] if self.concurrent_levels else
                                [main] * (len(feats) - 1))
        start = main.record_event()
        for li, x in enumerate(feats):
          st = lvl_streams[li]
          if st is not main:
              st.wait_event(start)
          with torch.cuda.stream(st):                          # (body indented for the `with`)
            n, _, h, w = x.shape
            dev = x.device
            cls_feat = pts_feat = to_channels_last(x)
            cls_prep = pts_prep = None
            for i, m in enumerate(self.cls_convs):
                y = _conv3x3_nhwc(cls_feat, m.conv)
                if i == last:
                    _, cls_prep = groupnorm_relu_nhwc(y, m.gn, dense=False, prepared_for=pf)
                else:
                    cls_feat = groupnorm_relu_nhwc(y, m.gn)
            for i, m in enumerate(self.reg_convs):
                y = _conv3x3_nhwc(pts_feat, m.conv)
                if i == last:
                    pts_feat, pts_prep = groupnorm_relu_nhwc(y, m.gn, prepared_for=pf)
                else:
                    pts_feat = groupnorm_relu_nhwc(y, m.gn)
            c = pts_feat.shape[1]
            kpt_init = x.new_empty((n, kd, h, w))
            rep_init = x.new_empty((n, rd, h, w))
            kpt_rows = nchw_to_tiled(_conv3x3_nhwc(pts_feat, self.keypts_init_conv, False), relu=True, split=True,
                                     bias=self.keypts_init_conv.bias)
            if par:
                pointwise_conv(kpt_rows, W['kpt_init'][0], W['kpt_init'][1], [(kpt_init, None, 0, kd)], h * w)
                rep_rows = nchw_to_tiled(_conv3x3_nhwc(pts_feat, self.reppts_init_conv, False), relu=True, split=True,
                                         bias=self.reppts_init_conv.bias)                    # PAR:314-315
                pointwise_conv(rep_rows, W['rep_init'][0], W['rep_init'][1], [(rep_init, None, 0, rd)], h * w)
            else:
                pointwise_conv(kpt_rows, W['kpt_rep_init'][0], W['kpt_rep_init'][1],
                               [(kpt_init, None, 0, kd), (rep_init, None, kd, kd + rd)], h * w)   # SER:314
            plan = prepare_plan_points(rep_init, 0, (n, c, h, w), pf, self.dcn_kernel, 1, self.dcn_pad, 1,
                                       precision='bf16', gradient_mul=self.gradient_mul)       # PAR:322-325
            rows = [TiledRows(n * h * w, pf, True, dev) for _ in range(3 if par else 2)]
            jobs.append((cls_prep, plan, self.cls_refine_dfmconv.weight, rows[0], 0, True))
            jobs.append((pts_prep, plan, self.keypts_refine_dfmconv.weight, rows[1], 0, True))
            if par:
                jobs.append((pts_prep, plan, self.reppts_refine_dfmconv.weight, rows[2], 0, True))
            per_level.append((n, h, w, kpt_init, rep_init, rows))
        for st in lvl_streams[1:]:
            if st is not main:
                main.wait_event(st.record_event())
        jobs.sort(key=lambda j: -j[3].M)                                           # biggest maps first
        for i in range(0, len(jobs), 6):
            deform_conv_prepared_group(jobs[i:i + 6])
        res = []
        grouped = main.record_event()
        for li, (n, h, w, kpt_init, rep_init, rows) in enumerate(per_level):
          st = lvl_streams[li]
          if st is not main:
              st.wait_event(grouped)
          with torch.cuda.stream(st):
            cls_out = kpt_init.new_empty((n, nc, h, w))
            kpt_ref = torch.empty_like(kpt_init)
            rep_ref = torch.empty_like(rep_init)
            pointwise_conv(rows[0], W['cls'][0], W['cls'][1], [(cls_out, None, 0, nc)], h * w)
            if par:
                pointwise_conv(rows[1], W['kpt_refine'][0], W['kpt_refine'][1], [(kpt_ref, kpt_init, 0, kd)], h * w)
                pointwise_conv(rows[2], W['rep_refine'][0], W['rep_refine'][1], [(rep_ref, rep_init, 0, rd)], h * w)
            else:
                pointwise_conv(rows[1], W['kpt_rep_refine'][0], W['kpt_rep_refine'][1],
                               [(kpt_ref, kpt_init, 0, kd), (rep_ref, rep_init, kd, kd + rd)], h * w)  # SER:330
            res.append((cls_out, kpt_init, kpt_ref, rep_init, rep_ref))
        for st in lvl_streams[1:]:
            if st is not main:
                main.wait_event(st.record_event())
        for lvl in res:
            for t in lvl:
                t.record_stream(main)
        return tuple(map(list, zip(*res)))

    def _forward_grouped(self, feats):
        """bf16 inference over all levels with the deformable convolutions of EVERY level (3 x 5 parallel, 2 x 5
        serial) in grouped persistent launches of up to six (kgdet_dcn_forward_prepared_group): one such call is
        a few tiles on the small levels (P6: 5 tiles per image batch of 8, P7: 1..5) and leaves the machine idle
        when launched alone; grouped, their tiles fill the SMs the big levels' tails leave free.  Same arithmetic
        as `forward_single` level by level (bit-identical)."""
        base = self._dcn_base.to(feats[0].dtype)
        pf = self.cls_refine_dfmconv.out_channels
        per_level, jobs = [], []
        for x in feats:
            cls_feat = pts_feat = x
            for m in self.cls_convs:
                cls_feat = m(cls_feat)
            for m in self.reg_convs:
                pts_feat = m(pts_feat)
            kpt_init = self.keypts_init_out(F.relu(self.keypts_init_conv(pts_feat)))
            if self.variant == 'parallel':
                rep_init = self.reppts_init_out(F.relu(self.reppts_init_conv(pts_feat)))     # PAR:314-315
            else:
                rep_init = self.reppts_init_out(kpt_init)                                    # SER:314
            n, c, h, w = cls_feat.shape
            pts = self.gradient_mul * rep_init + (1 - self.gradient_mul) * rep_init           # PAR:322-325, value-wise
            plan = prepare_plan(pts - base, (n, c, h, w), pf, self.dcn_kernel, 1, self.dcn_pad, 1, like_dtype=x.dtype)
            cls_prep = prepare_input(cls_feat, pf)
            pts_prep = prepare_input(pts_feat, pf)
            outs = [x.new_empty((n, pf, h, w)) for _ in range(3 if self.variant == 'parallel' else 2)]
            jobs.append((cls_prep, plan, self.cls_refine_dfmconv.weight, outs[0], 0, True))
            jobs.append((pts_prep, plan, self.keypts_refine_dfmconv.weight, outs[1], 0, True))
            if self.variant == 'parallel':
                jobs.append((pts_prep, plan, self.reppts_refine_dfmconv.weight, outs[2], 0, True))
            per_level.append((kpt_init, rep_init, outs))
        jobs.sort(key=lambda j: -j[3].shape[0] * j[3].shape[2] * j[3].shape[3])      # biggest maps first
        for i in range(0, len(jobs), 6):
            deform_conv_prepared_group(jobs[i:i + 6])
        res = []
        for kpt_init, rep_init, outs in per_level:
            cls_out = self.cls_refine_out(outs[0])
            kpt_ref = self.keypts_refine_out(outs[1])
            rep_ref = self.reppts_refine_out(outs[2]) if self.variant == 'parallel' else self.reppts_refine_out(kpt_ref)
            res.append((cls_out, kpt_init, kpt_ref + kpt_init, rep_init, rep_ref + rep_init))
        return tuple(map(list, zip(*res)))


    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def get_bboxes(self, cls_scores, keypts_preds_refine, reppts_preds_refine, img_shapes, score_thr=0.05, iou_thr=0.5,
                   nms_pre=1000, max_per_img=100, score_override=None, scale_factors=None, return_kept=False):
        """Batched, sync-free form of get_bboxes + get_bboxes_single of the two baseline heads (PAR:615-752, the
        same lines of SER) + multiclass_nms_kp (core/post_processing/bbox_nms_kp.py:6-75), over ALL levels.

        cls_scores / keypts_preds_refine / reppts_preds_refine: per-level lists ([B,13,H,W], [B,588,H,W],
        [B,18,H,W]: outputs 0, 2 and 4 of `forward`).  Per level the `nms_pre` best positions are selected and
        decoded (kgdet_bbox_select / kgdet_bbox_decode on the GPU), the candidates of all levels form one segment
        per (image, class) for ONE batched NMS launch, and keypoints are decoded only for the `max_per_img`
        survivors.  The keypoint clamp is this head's own (PAR:721-722 index the keypoint axis: keypoints 0, 3, 6...
        are limited to the image width in x AND y, keypoints 1, 4, 7... to the height, the others not at all).
        scale_factors: optional per-image floats = the reference's `rescale=True` (PAR:732-737: boxes and keypoint
        coordinates divided by the scale factor before the NMS).
        Returns dets [B, max_per_img, 5], labels [B, max_per_img] (-1 = empty slot), kpts [B, max_per_img, 294*3]."""
        B, dev = cls_scores[0].shape[0], cls_scores[0].device
        key = (tuple(tuple(s[:2]) for s in img_shapes), str(dev))
        wh = self._lim_cache.get(key)
        if wh is None:            # built once per (shapes, device): keeps H2D copies out of graph capture
            wh = torch.tensor([[s[1], s[0]] for s in img_shapes], dtype=torch.float32, device=dev)
            self._lim_cache[key] = wh
        fused = (self._fused_decode and cls_scores[0].is_cuda and self._nms_flags_fn is batched_nms_flags
                 and all(c.shape[-2] * c.shape[-1] <= 40960 for c in cls_scores) and 0 < nms_pre <= 4096)
        C = cls_scores[0].shape[1]
        boxes_l, dets_l, pos_l, lvl_of = [], [], [], []
        # per-level selection + decode as parallel branches (one CTA per image each: latency, not throughput)
        par = fused and self.concurrent_levels and len(cls_scores) > 1
        main = torch.cuda.current_stream(dev) if par else None
        lvl_streams = ([main] + _streams(dev, 8 + len(cls_scores))[9:8 + len(cls_scores)]) if par else None
        start = main.record_event() if par else None
        for lvl, (cs, rp) in enumerate(zip(cls_scores, reppts_preds_refine)):
          st = lvl_streams[lvl] if par else None
          if par and st is not main:
              st.wait_event(start)
          with (torch.cuda.stream(st) if par else _null_context()):
            bp = self.points2bbox(rp)                                                   # PAR:628-631
            stride = self.point_strides[lvl]
            H, W = cs.shape[-2:]
            n_l = min(nms_pre, H * W) if nms_pre > 0 else H * W
            src = (cs if score_override is None else score_override[lvl]).float().contiguous()
            sig = score_override is None
            if fused:
                order = bbox_select(src, sig, n_l)
                boxes, dets = bbox_decode(src, sig, bp.float().contiguous(), order, wh, stride)
                order = order.long()
            else:
                scores = src.reshape(B, C, H * W)
                scores = scores.sigmoid() if sig else scores
                if n_l < H * W:                                                         # PAR:703-713
                    order = scores.max(dim=1)[0].topk(n_l, dim=1)[1]
                else:
                    order = torch.arange(H * W, device=dev)[None].expand(B, -1)
                ctr = torch.stack([(order % W).float() * stride, (order // W).float() * stride], -1)
                bbox = bp.float().reshape(B, 4, H * W).gather(2, order[:, None].expand(-1, 4, -1)).transpose(1, 2)
                boxes = bbox * stride + torch.cat([ctr, ctr], -1)                       # PAR:714-716
                boxes = torch.min(boxes.clamp(min=0), torch.cat([wh, wh], 1)[:, None])  # PAR:723-727
                sc = scores.gather(2, order[:, None].expand(-1, C, -1))                # [B, C, n_l]
                dets = torch.cat([boxes[:, None].expand(B, C, n_l, 4), sc[..., None]], -1)
            boxes_l.append(boxes)
            dets_l.append(dets)
            pos_l.append(order)
            lvl_of.append(torch.full((n_l,), lvl, dtype=torch.long, device=dev))
        if par:
            for st in lvl_streams[1:]:
                main.wait_event(st.record_event())
        boxes = torch.cat(boxes_l, 1)                                                   # [B, n, 4]
        dets = torch.cat(dets_l, 2).contiguous()                                        # [B, C, n, 5]
        pos = torch.cat(pos_l, 1)                                                       # [B, n] position in its level
        lvl_of = torch.cat(lvl_of)                                                      # [n]
        n = boxes.shape[1]
        sf = None
        if scale_factors is not None:                                                   # PAR:732-733
            skey = ('scale', tuple(float(f) for f in scale_factors), str(dev))
            sf = self._lim_cache.get(skey)
            if sf is None:
                sf = torch.tensor([float(f) for f in scale_factors], dtype=torch.float32, device=dev)
                self._lim_cache[skey] = sf
            boxes = boxes / sf.view(B, 1, 1)
            dets[..., :4] = dets[..., :4] / sf.view(B, 1, 1, 1)
        flags = self._nms_flags_fn(dets.view(-1, 5), None, n, iou_thr, score_thr=score_thr)
        masked = torch.where(flags.view(B, C * n).bool(), dets[..., 4].reshape(B, C * n), dets.new_full((), -1.0))
        k = min(max_per_img, C * n)
        top_s, top_i = masked.topk(k, dim=1)                                            # bbox_nms_kp.py:64-70
        valid = top_s > 0
        cls_i, row_i = top_i // n, top_i % n
        out_boxes = boxes.gather(1, row_i[..., None].expand(-1, -1, 4))
        out_dets = torch.cat([out_boxes, top_s[..., None]], -1) * valid[..., None]
        out_labels = torch.where(valid, cls_i, torch.full_like(cls_i, -1))
        # keypoints of the survivors only: gathered from their level's map
        p_i = pos.gather(1, row_i)                                                      # [B, k]
        l_i = lvl_of[row_i]                                                             # [B, k]
        P = self.num_keypts
        kp = top_s.new_zeros((B, k, 2 * P))
        ctr = top_s.new_zeros((B, k, 2))
        strd = top_s.new_zeros((B, k))
        picked = []
        if par:
            chosen = main.record_event()
        for lvl, kmap in enumerate(keypts_preds_refine):
          st = lvl_streams[lvl] if par else None
          if par and st is not main:
              st.wait_event(chosen)
          with (torch.cuda.stream(st) if par else _null_context()):
            H, W = kmap.shape[-2:]
            here = l_i == lvl
            pl = torch.where(here, p_i, torch.zeros_like(p_i))
            v = kmap.float().reshape(B, 2 * P, H * W).gather(2, pl[:, None].expand(-1, 2 * P, -1)).transpose(1, 2)
            c_l = torch.stack([(pl % W).float(), (pl // W).float()], -1) * self.point_strides[lvl]
            picked.append((here, v, c_l))
        if par:
            for st in lvl_streams[1:]:
                main.wait_event(st.record_event())
        for lvl, (here, v, c_l) in enumerate(picked):
            kp = torch.where(here[..., None], v, kp)
            ctr = torch.where(here[..., None], c_l, ctr)
            strd = torch.where(here, torch.full_like(strd, float(self.point_strides[lvl])), strd)
        kxy = kp.view(B, k, P, 2).flip(-1)                                              # points2kpt: (y, x) -> (x, y)
        kxy = kxy * strd[..., None, None] + ctr[:, :, None, :]                          # PAR:717-719
        idx = torch.arange(P, device=dev) % 3
        w_, h_ = wh[:, 0].view(B, 1, 1, 1), wh[:, 1].view(B, 1, 1, 1)
        by_w = torch.min(kxy.clamp(min=0), w_)
        by_h = torch.min(kxy.clamp(min=0), h_)
        kxy = torch.where((idx == 0).view(1, 1, P, 1), by_w, torch.where((idx == 1).view(1, 1, P, 1), by_h, kxy))
        if sf is not None:                                                              # PAR:734-736
            kxy = kxy / sf.view(B, 1, 1, 1)
        out_kpts = (torch.cat([kxy, torch.ones_like(kxy[..., :1])], -1) * valid[..., None, None]).reshape(B, k, -1)
        if return_kept:         # boxes the NMS kept per image (the reference sorts only above max_per_img)
            return out_dets, out_labels, out_kpts, flags.view(B, -1).sum(1), top_i
        return out_dets, out_labels, out_kpts


class RepPointsKpDetect(nn.Module):
    """`head(feats)` + `head.get_bboxes` of a RepPointsKpHead as one callable (simple_test of the reference's
    detector minus backbone / neck: mmdet/models/detectors/single_stage_kp.py:75-86), e.g. for GraphedForward.
    score_override: optional per-level sigmoid scores (benchmarks with random-init weights)."""

    def __init__(self, head, img_shapes, score_thr=0.05, iou_thr=0.5, nms_pre=1000, max_per_img=100, score_override=None):
        super().__init__()
        self.head = head
        self.img_shapes = list(img_shapes)
        self.cfg = (score_thr, iou_thr, nms_pre, max_per_img)
        self.score_override = score_override

    def forward(self, feats):
        outs = self.head(feats)
        return self.head.get_bboxes(outs[0], outs[2], outs[4], self.img_shapes, *self.cfg,
                                    score_override=self.score_override)


class GraphedForward(object):
    """`head(feats)` of any head of this module captured ONCE into a CUDA graph for fixed input shapes (the forward
    has static shapes and no host synchronisation).  ``__call__(feats)`` copies the inputs into the static ones and
    replays; the returned structure holds the graph's static output tensors.  A snapshot of the weights at capture
    time, like GraphedInference."""

    def __init__(self, head, example_feats, warmup=3):
        assert all(x.is_cuda for x in example_feats)
        self.head = head
        self.static_in = [x.detach().clone() for x in example_feats]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):
                head(self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = head(self.static_in)
        from .ops.pointwise import cache_values
        self._pinned_buffers = cache_values()

    def __call__(self, feats=None):
        if feats is not None:
            for s, x in zip(self.static_in, feats):
                if x.data_ptr() != s.data_ptr():
                    s.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out


class GraphedInference(object):
    """forward_single + get_bboxes of a KGDetHead captured ONCE into a CUDA graph for a fixed input shape.

    The whole step (~150 kernels: cuDNN towers, 2 layout transforms, 6 sample plans, 12 fused tcgen05
    deformable convolutions, 3 moment transforms, decode, one batched NMS, top-k) has static shapes and
    no host synchronisation, so a replay costs one launch on the host instead of ~4 ms of Python/launch
    overhead.  ``__call__(x)`` copies ``x`` (device or pinned host tensor) into the static input and
    replays; the returned (dets, labels, kpts) are the graph's static output tensors.
    """

    def __init__(self, head, example_x, img_shapes, score_thr=0.05, iou_thr=0.5, nms_pre=1000, max_per_img=100,
                 score_override=None, warmup=3):
        assert example_x.is_cuda, 'GraphedInference needs a CUDA example input'
        self.head = head
        self.args = (img_shapes, score_thr, iou_thr, nms_pre, max_per_img)
        self.score_override = score_override
        self.static_x = example_x.detach().clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():       # packs weights, picks cuDNN algorithms, warms the allocator
            for _ in range(max(warmup, 1)):
                self._run()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self._capture()

    def _capture(self):
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = self._run()
        # the captured kernels hold raw pointers into the derived-weight caches (packed DCN / 1x1 weights,
        # channels_last copies): this object keeps those buffers alive whatever happens to the caches
        from .ops.pointwise import cache_values
        self._pinned_buffers = cache_values()
        self._twin = None
        self._more = None          # further instances of serve(instances > 2): captured again on demand

    def refresh_weights(self):
        """Re-derive every packed weight from the head's CURRENT parameters and capture the step again.  A
        GraphedInference is a snapshot of the weights at capture time; call this after the parameters changed
        (optimizer steps, `load_state_dict`, in-place `.data` writes, replays of a captured training graph)."""
        from .ops import invalidate_weight_caches
        torch.cuda.synchronize()
        invalidate_weight_caches()
        with torch.no_grad():
            self._run()
        torch.cuda.synchronize()
        self._capture()

    def _run(self):
        o = self.head.forward_single(self.static_x)
        shapes, score_thr, iou_thr, nms_pre, max_per_img = self.args
        return self.head.get_bboxes([o[2]], [o[5]], [o[8]], shapes, score_thr, iou_thr, nms_pre, max_per_img,
                                    score_override=None if self.score_override is None else [self.score_override])

    def __call__(self, x=None):
        if x is not None and x.data_ptr() != self.static_x.data_ptr():
            self.static_x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out

    def _new_instance(self):
        twin = object.__new__(GraphedInference)
        twin.head, twin.args, twin.score_override = self.head, self.args, self.score_override
        twin.static_x = self.static_x.clone()
        torch.cuda.synchronize()
        twin.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(twin.graph), torch.no_grad():
            twin.static_out = twin._run()
        from .ops.pointwise import cache_values
        twin._pinned_buffers = cache_values()
        twin._twin = None
        return twin

    def _second_instance(self):
        """A second capture of the same step with its own static input / outputs (lazy: only `serve` needs it)."""
        if getattr(self, '_twin', None) is None:
            self._twin = self._new_instance()
        return self._twin

    def _instances(self, n):
        """This capture plus n - 1 further captures of the same step (own static input / outputs, own scratch)."""
        more = getattr(self, '_more', None)
        if more is None:
            more = self._more = []
        while len(more) < n - 2:
            more.append(self._new_instance())
        return [self, self._second_instance()][:max(n, 1)] + more[:max(n - 2, 0)]

    def serve(self, host_batches, host_outputs=None, before_step=None, concurrent=None, instances=None):
        """Throughput path for HOST inputs: a three-stage software pipeline over the batches --

            copy-in stream   pinned host batch i+1 -> static input of graph instance (i+1) % 2   (PCIe, host -> device)
            compute stream   replay of instance i % 2
            copy-out stream  static results of instance (i-1) % 2 -> pinned host buffers         (PCIe, device -> host)

        -- the step is captured twice (two instances with their own static input and outputs), so both copies of
        neighbouring batches run under the replay of the current one and no device-side staging copy is needed.
        `host_batches`: iterable of pinned CPU tensors shaped like the example input.  `host_outputs`: optional list
        (one entry per batch) of pinned (dets, labels, keypoints) triples to fill; allocated when omitted.
        `before_step(i)`: optional callable issued on the compute stream before replay i (bench.py flushes the L2
        there).  `concurrent` (default on; environment KGDET_SERVE_CONCURRENT=0 turns it off): the two instances replay on their
        own compute streams, so batch i + 1 starts while batch i is still running and the narrow phases of one step
        (towers, stage 1, the point-refinement gap, decode + NMS) overlap the wide kernels of the other; the
        instances share nothing but read-only weights.  Returns the list of host triples; everything has completed
        on return."""
        dev = self.static_x.device
        main = torch.cuda.current_stream(dev)
        if concurrent is None:
            concurrent = os.environ.get('KGDET_SERVE_CONCURRENT', '1') != '0'
        if instances is None:
            instances = int(os.environ.get('KGDET_SERVE_INSTANCES', '2'))
        ni = max(2, int(instances)) if concurrent else 2
        inst = self._instances(ni)
        if not hasattr(self, '_pipe'):
            self._pipe = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        cin, cout = self._pipe
        if concurrent:
            if len(getattr(self, '_compute', ())) < ni:
                self._compute = tuple(torch.cuda.Stream(device=dev) for _ in range(ni))
            comp = self._compute[:ni]
        else:
            comp = (main,) * ni
        start = main.record_event()
        cin.wait_event(start)
        cout.wait_event(start)
        if concurrent:
            for c in comp:
                c.wait_event(start)
        replayed = [None] * ni         # instance b has consumed its static input (and rewritten its outputs)
        drained = [None] * ni          # copy-out has read instance b's static outputs
        results = []
        batches = list(host_batches)

        def copy_in(i):
            b = i % ni
            if replayed[b] is not None:
                cin.wait_event(replayed[b])
            with torch.cuda.stream(cin):
                inst[b].static_x.copy_(batches[i], non_blocking=True)
                return cin.record_event()

        ready = copy_in(0) if batches else None
        for i in range(len(batches)):
            b = i % ni
            nxt = copy_in(i + 1) if i + 1 < len(batches) else None     # overlaps replay i
            cs = comp[b]
            cs.wait_event(ready)
            if drained[b] is not None:
                cs.wait_event(drained[b])                              # results of batch i - 2 have left the device
            with torch.cuda.stream(cs):
                if before_step is not None:
                    before_step(i)
                inst[b].graph.replay()
                replayed[b] = cs.record_event()
            host = host_outputs[i] if host_outputs is not None else \
                tuple(torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in self.static_out)
            cout.wait_event(replayed[b])
            with torch.cuda.stream(cout):
                for h, d in zip(host, inst[b].static_out):
                    h.copy_(d, non_blocking=True)
                drained[b] = cout.record_event()
            results.append(host)
            ready = nxt
        if concurrent:
            for c in comp:
                main.wait_stream(c)
        main.wait_stream(cout)
        main.wait_stream(cin)
        main.synchronize()
        return results
