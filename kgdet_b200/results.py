"""Wire formats of the detections: the per-class result tuple of
`RepPointsDetectorKp.bbox2result_kp` (mmdet/models/detectors/reppoints_detector_kp.py:55-78) and the
DeepFashion2 / COCO-style JSON records of `kpt2json` (mmdet/core/evaluation/coco_utils.py:121-154), from the
fixed-size, padded output of `KGDetHead.get_bboxes` / `GraphedInference` (labels == -1 marks an empty slot).
Host-side Python like the reference's; parity: tests/test_results_cpu.py runs the unchanged reference functions.
"""
import numpy as np


def detections_to_result(dets, labels, kpts, num_classes):
    """One image: dets [k, 5], labels [k] (-1 = empty slot), kpts [k, P*3] (tensors or arrays) -> the tuple
    bbox2result_kp returns: ([per-class [n_c, 5] arrays], scores [n], [per-class [n_c, P*3] arrays]); a 1-tuple of
    empty per-class arrays when nothing was detected (reppoints_detector_kp.py:66-70)."""
    dets, labels, kpts = (np.asarray(t.detach().cpu().numpy() if hasattr(t, 'detach') else t) for t in (dets, labels, kpts))
    keep = labels >= 0
    dets, labels, kpts = dets[keep], labels[keep], kpts[keep]
    if dets.shape[0] == 0:
        return ([np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes - 1)],)
    return ([dets[labels == i, :] for i in range(num_classes - 1)], dets[:, 4],
            [kpts[labels == i, :] for i in range(num_classes - 1)])


def batch_to_results(dets, labels, kpts, num_classes):
    """[B, k, ...] padded batch output -> list of per-image result tuples."""
    return [detections_to_result(dets[b], labels[b], kpts[b], num_classes) for b in range(dets.shape[0])]


def _xyxy2xywh(b):
    b = b.tolist()
    return [b[0], b[1], b[2] - b[0] + 1, b[3] - b[1] + 1]                # coco_utils.py:79-86 (+1 pixel convention)


def kpt2json(img_ids, cat_ids, results, num_digits=4):
    """coco_utils.py:121-154: (bbox records, keypoint records); images whose result is not a 3-tuple (nothing
    detected) are skipped, values are rounded to 4 digits, a keypoint record carries its box's score."""
    bbox_json, kpt_json = [], []
    for img_id, res in zip(img_ids, results):
        if len(res) != 3:
            continue
        det, _, kpt = res
        for label in range(len(det)):
            bboxes = det[label]
            for i in range(bboxes.shape[0]):
                bbox_json.append(dict(image_id=img_id, bbox=[round(v, num_digits) for v in _xyxy2xywh(bboxes[i])],
                                      score=round(float(bboxes[i][4]), num_digits), category_id=cat_ids[label]))
            kpts = kpt[label]
            for i in range(kpts.shape[0]):
                kpt_json.append(dict(image_id=img_id,
                                     keypoints=np.round(kpts[i].astype(np.float64), num_digits).tolist(),
                                     score=round(float(bboxes[i][4]), num_digits), category_id=cat_ids[label]))
    return bbox_json, kpt_json
