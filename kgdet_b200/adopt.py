"""Put the fused paths of this package behind an UNCHANGED reference head object.

`kgdet_b200.mount_as_mmdet_ops()` makes the reference heads run on this library's operators one call at a time.
The fused paths (prepared inputs, grouped persistent deformable convolutions, tcgen05 1x1 GEMMs, batched NMS, the
assignment + loss kernels) live in `kgdet_b200.head`, whose modules are state-dict compatible restatements of

    RepPointsHeadKp3RepCas1AssignOnce   mmdet/models/anchor_heads/reppoints_head_kp3rep_cas_1_assign_once.py:183-914
    RepPointsHeadKpParallel / ...Serial reppoints_head_kp_parallel.py:17-752, reppoints_head_kp_serial.py:17-752

`accelerate(ref_head)` builds the matching restatement AROUND THE REFERENCE OBJECT'S OWN PARAMETERS (the very
`nn.Parameter` objects: optimisers, checkpoints, `.to()` and gradients keep working through the reference object)
and rebinds, on that instance only,

    forward_single / forward   -> the restatement's (same 9- / 5-tuples, KP3:412-446, PAR:292-341)
    get_bboxes                 -> the batched, sync-free post-processing, returned in the reference's format: a list of
                                  (det_bboxes [k, 5], det_labels [k], det_kpts) per image (KP3:770-914, PAR:615-752),
                                  `rescale=True` included
    loss (KGDet head)          -> target assignment + the nine losses as three kernels, same dict of per-level lists
                                  (KP3:670-768)

Anything the fused paths do not cover (`nms=False`, a point_strides / loss configuration other than the reference
configs') falls through to the reference's own method, still running on this package's operators.
"""
import types

import torch

from . import head as _head


def _share_parameters(mirror, ref_head):
    """Make every parameter of `mirror` BE the reference head's parameter of the same name."""
    ref_params = dict(ref_head.named_parameters())
    names = [n for n, _ in mirror.named_parameters()]
    missing = [n for n in names if n not in ref_params]
    extra = [n for n in ref_params if n not in set(names)]
    if missing or extra:
        raise ValueError('the reference head does not match the restatement: missing %r, unexpected %r'
                         % (missing[:5], extra[:5]))
    for n in names:
        mod = mirror
        parts = n.split('.')
        for q in parts[:-1]:
            mod = getattr(mod, q)
        if tuple(mod._parameters[parts[-1]].shape) != tuple(ref_params[n].shape):
            raise ValueError('parameter %s: shape %r != %r' % (n, tuple(mod._parameters[parts[-1]].shape),
                                                               tuple(ref_params[n].shape)))
        mod._parameters[parts[-1]] = ref_params[n]


def mirror_of(ref_head, **inject):
    """The `kgdet_b200.head` restatement of an unchanged reference head, sharing its parameters.
    `inject`: optional `deform_conv_cls` / `moment_fn` / `nms_flags_fn` stand-ins (CPU tests)."""
    kind = type(ref_head).__name__
    norm = getattr(ref_head, 'norm_cfg', None) or {}
    common = dict(num_classes=ref_head.num_classes, in_channels=ref_head.in_channels,
                  feat_channels=ref_head.feat_channels, point_feat_channels=ref_head.point_feat_channels,
                  stacked_convs=ref_head.stacked_convs, num_keypts=ref_head.num_keypts,
                  gradient_mul=ref_head.gradient_mul, point_strides=tuple(ref_head.point_strides),
                  moment_mul=ref_head.moment_mul, num_groups=norm.get('num_groups', 32))
    if getattr(ref_head, 'transform_method', 'moment') != 'moment' or not getattr(ref_head, 'use_sigmoid_cls', True):
        raise ValueError('only transform_method="moment" with sigmoid classification is restated')
    if norm.get('type', 'GN') != 'GN':
        raise ValueError('only GroupNorm towers are restated (norm_cfg %r)' % (norm,))
    if kind == 'RepPointsHeadKp3RepCas1AssignOnce':
        mirror = _head.KGDetHead(**common, **inject)
    elif kind in ('RepPointsHeadKpParallel', 'RepPointsHeadKpSerial'):
        mirror = _head.RepPointsKpHead('parallel' if kind.endswith('Parallel') else 'serial',
                                       num_reppts=ref_head.num_reppts, **common, **inject)
    else:
        raise ValueError('no restatement of %s in kgdet_b200.head' % kind)
    _share_parameters(mirror, ref_head)
    p = next(ref_head.parameters())
    for name, buf in list(mirror.named_buffers()):            # own constant buffers follow the parameters' device
        mod = mirror
        parts = name.split('.')
        for q in parts[:-1]:
            mod = getattr(mod, q)
        mod._buffers[parts[-1]] = buf.to(p.device)
    mirror.train(ref_head.training)
    return mirror


def _cfg_get(cfg, key, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def _to_reference_lists(dets, labels, kpts, kept, flat_index, max_per_img, num_keypts, flat):
    """Padded, score-sorted batch results -> the reference's list of (det_bboxes, det_labels, det_kpts) per image.
    multiclass_nms_kp concatenates the per-class survivors -- inside a class in candidate order, `nms` returns its
    keep indices ascending (nms_wrapper.py:8-49) -- and sorts them by score ONLY when there are more than
    max_per_img of them (bbox_nms_kp.py:64-70)."""
    counts = (labels >= 0).sum(1).tolist()                     # the one host synchronisation ...
    kept = kept.tolist()                                       # (... and its second half)
    out = []
    for i, k in enumerate(counts):
        d, l, kp = dets[i, :k], labels[i, :k], kpts[i, :k]
        if kept[i] <= max_per_img and k > 1:
            o = torch.sort(flat_index[i, :k])[1]               # class-major, candidate order inside a class
            d, l, kp = d[o], l[o], kp[o]
        out.append((d, l, kp if flat else kp.reshape(k, num_keypts, 3)))
    return out


def accelerate(ref_head, **inject):
    """Rebind `forward_single`, `forward`, `get_bboxes` (and, for the KGDet head, `loss`) of this reference head
    INSTANCE to the fused paths.  Returns the same object; `ref_head.kgdet_mirror` is the restatement."""
    mirror = mirror_of(ref_head, **inject)
    ref_head.kgdet_mirror = mirror
    kgdet = isinstance(mirror, _head.KGDetHead)
    cls = type(ref_head)

    def sync():
        if mirror.training != ref_head.training:
            mirror.train(ref_head.training)
        dev = next(ref_head.parameters()).device
        for name, buf in list(mirror.named_buffers()):
            if buf.device != dev:
                mod = mirror
                parts = name.split('.')
                for q in parts[:-1]:
                    mod = getattr(mod, q)
                mod._buffers[parts[-1]] = buf.to(dev)

    def forward_single(self, x):
        sync()
        return mirror.forward_single(x)

    def forward(self, feats, img_metas=None):                  # KP3:490-495, PAR:343-344
        if getattr(self, 'flip_forward', False):
            return cls.forward(self, feats, img_metas)         # the flip-fusion wrapper calls the rebound forward_single
        sync()
        return mirror.forward(feats)

    def get_bboxes(self, *args, **kwargs):
        n_out = 9 if kgdet else 5
        outs, rest = args[:n_out], list(args[n_out:])
        names = ['img_metas', 'cfg', 'rescale', 'nms']
        vals = dict(rescale=False, nms=True)
        vals.update(dict(zip(names, rest)))
        vals.update(kwargs)
        img_metas, cfg = vals['img_metas'], vals['cfg']
        nms_cfg = _cfg_get(cfg, 'nms', {}) or {}
        scales = [m.get('scale_factor', 1.0) for m in img_metas]
        supported = (vals['nms'] and _cfg_get(nms_cfg, 'type', 'nms') == 'nms'
                     and all(isinstance(f, (int, float)) for f in scales))
        if not supported:
            return cls.get_bboxes(self, *args, **kwargs)
        sync()
        shapes = [tuple(m['img_shape'][:2]) for m in img_metas]
        common = (shapes, float(_cfg_get(cfg, 'score_thr', 0.05)), float(_cfg_get(nms_cfg, 'iou_thr', 0.5)),
                  int(_cfg_get(cfg, 'nms_pre', -1)), int(_cfg_get(cfg, 'max_per_img', 100)))
        sf = [float(f) for f in scales] if vals['rescale'] else None
        with torch.no_grad():
            if kgdet:
                res = mirror.get_bboxes(list(outs[2]), list(outs[5]), list(outs[8]), *common, scale_factors=sf,
                                        return_kept=True)
            else:
                res = mirror.get_bboxes(list(outs[0]), list(outs[2]), list(outs[4]), *common, scale_factors=sf,
                                        return_kept=True)
        # KP3:894-898 / PAR:734-737: the keypoints come back flat [k, 3P] only on the rescale branch
        return _to_reference_lists(*res, max_per_img=common[4], num_keypts=mirror.num_keypts, flat=bool(vals['rescale']))

    ref_head.forward_single = types.MethodType(forward_single, ref_head)
    ref_head.forward = types.MethodType(forward, ref_head)
    ref_head.get_bboxes = types.MethodType(get_bboxes, ref_head)

    def standard_losses():
        # the fused loss kernels carry the reference configs' loss settings (kgdet_moment_r50_fpn_1x-*.py:38-62)
        try:
            for st, w in ((1, 0.5), (2, 0.5), (3, 1.0)):
                lc, lb, lk = (getattr(ref_head, 'loss_%s_%d' % (k, st)) for k in ('cls', 'bbox', 'kpt'))
                if (abs(lc.loss_weight - w) > 1e-12 or abs(lb.loss_weight - w) > 1e-12 or abs(lk.loss_weight - w) > 1e-12
                        or abs(lc.gamma - 2.0) > 1e-12 or abs(lc.alpha - 0.25) > 1e-12
                        or abs(lb.beta - 1.0 / 9.0) > 1e-9 or abs(lk.beta - 1.0 / 9.0) > 1e-9
                        or not lc.use_sigmoid or lc.reduction != 'mean'):
                    return False
            return True
        except AttributeError:
            return False

    if kgdet and standard_losses():
        def loss(self, *args, **kwargs):
            names = ['gt_bboxes', 'gt_labels', 'gt_keypoints', 'img_metas', 'cfg', 'gt_bboxes_ignore']
            outs, rest = args[:9], list(args[9:])
            vals = dict(gt_bboxes_ignore=None)
            vals.update(dict(zip(names, rest)))
            vals.update(kwargs)
            cfg = vals['cfg']
            assigner = _cfg_get(cfg, 'assigner', {}) or {}
            single = all(len(o) == 1 for o in outs)
            cuda = outs[0][0].is_cuda
            if (not single or not cuda or vals['gt_bboxes_ignore'] is not None
                    or _cfg_get(assigner, 'type', 'PointAssigner') != 'PointAssigner'):
                return cls.loss(self, *args, **kwargs)
            from . import targets as T
            sync()
            dev = outs[0][0].device
            max_gts = max(max(int(b.shape[0]) for b in vals['gt_bboxes']), 1)
            boxes, labels, kps, valid = T.pad_ground_truth(vals['gt_bboxes'], vals['gt_labels'], vals['gt_keypoints'],
                                                           device=dev, max_gts=max_gts)
            flat = [o[0] for o in outs]
            losses = mirror.loss(flat, boxes, labels, kps, valid, int(_cfg_get(assigner, 'scale', 4)),
                                 int(_cfg_get(assigner, 'pos_num', 25)), int(self.point_base_scale))
            return {k: [v] for k, v in losses.items()}               # per-level lists, as multi_apply returns them

        ref_head.loss = types.MethodType(loss, ref_head)
    return ref_head
