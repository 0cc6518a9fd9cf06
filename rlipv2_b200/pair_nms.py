"""Pair-wise NMS of the HOI evaluation path.

Mirror of `HICOEvaluator.triplet_nms_filter` / `pairwise_nms` (/root/reference/datasets/hico_eval.py:493-564, switched on by
`--use_nms_filter --thres_nms 0.7 --nms_alpha 1.0 --nms_beta 0.5`, main.py:375-378): predictions of one image that share
the (subject category, object category, verb) triplet are visited in descending score order and a prediction suppresses
every later one whose  IoU(subjects)^alpha * IoU(objects)^beta  exceeds the threshold (boxes as inclusive pixel
rectangles: +1 on widths / heights).  Same inputs, same outputs (dicts in, dicts out, kept predictions in the
reference's order), so it drops in for the evaluator's method.

The reference recomputes the overlaps of the current head against the remaining list in every round of its while loop
(O(n^2) numpy calls of shrinking size); here the full overlap matrix of a triplet group is formed once and the greedy pass
is a scan over its rows - identical decisions (every entry is the same expression on the same operands)."""
import numpy as np


def overlap_matrix(subs, objs, alpha, beta):
    """[n, 4] xyxy subject / object boxes -> [n, n] pair overlap  IoU_sub^alpha * IoU_obj^beta  (hico_eval.py:539-560)"""
    def iou(b):
        x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
        area = (x2 - x1 + 1) * (y2 - y1 + 1)
        w = np.maximum(0.0, np.minimum(x2[:, None], x2[None, :]) - np.maximum(x1[:, None], x1[None, :]) + 1)
        h = np.maximum(0.0, np.minimum(y2[:, None], y2[None, :]) - np.maximum(y1[:, None], y1[None, :]) + 1)
        inter = w * h
        return inter / (area[:, None] + area[None, :] - inter)
    return np.power(iou(subs), alpha) * np.power(iou(objs), beta)


def pairwise_nms(subs, objs, scores, thres_nms=0.7, nms_alpha=1.0, nms_beta=0.5):
    """-> indices kept, in the order the reference returns them (descending score; hico_eval.py:525-564)"""
    subs, objs, scores = np.asarray(subs), np.asarray(objs), np.asarray(scores)
    if scores.size == 0:
        return []
    order = scores.argsort()[::-1]                       # the reference's own call: same tie order
    ovr = overlap_matrix(subs, objs, nms_alpha, nms_beta)
    alive = np.ones(order.size, dtype=bool)              # by position in `order`
    keep = []
    for pos in range(order.size):
        if not alive[pos]:
            continue
        i = order[pos]
        keep.append(i)
        rest = order[pos + 1:]
        alive[pos + 1:] &= ovr[i, rest] <= thres_nms
    return keep


def triplet_nms_filter(preds, thres_nms=0.7, nms_alpha=1.0, nms_beta=0.5):
    """preds: [{'filename', 'predictions': [{'bbox', 'category_id'}], 'hoi_prediction': [{'subject_id', 'object_id',
    'category_id', 'score'}]}] -> the same list with suppressed HOI predictions removed (hico_eval.py:493-523)"""
    out = []
    for img in preds:
        boxes, hois = img["predictions"], img["hoi_prediction"]
        groups = {}
        for index, hoi in enumerate(hois):
            key = (boxes[hoi["subject_id"]]["category_id"], boxes[hoi["object_id"]]["category_id"], hoi["category_id"])
            groups.setdefault(key, []).append(index)
        keep_all = []
        for idxs in groups.values():                     # dict order = first appearance, as in the reference
            subs = np.array([boxes[hois[i]["subject_id"]]["bbox"] for i in idxs])
            objs = np.array([boxes[hois[i]["object_id"]]["bbox"] for i in idxs])
            scores = np.array([hois[i]["score"] for i in idxs])
            keep_all.extend(idxs[k] for k in pairwise_nms(subs, objs, scores, thres_nms, nms_alpha, nms_beta))
        out.append({"filename": img["filename"], "predictions": boxes, "hoi_prediction": [hois[i] for i in keep_all]})
    return out
