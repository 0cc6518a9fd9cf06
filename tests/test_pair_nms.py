"""Pair-wise NMS (rlipv2_b200/pair_nms.py) against the reference's own `HICOEvaluator.triplet_nms_filter`
(/root/reference/datasets/hico_eval.py:493-564): golden fixture from oracle/gen_golden_pair_nms.py (two settings, images
with and without score ties), plus the greedy rule restated element by element."""
import json
import os

import numpy as np

from tests.golden_util import GOLDEN


def test_matches_the_reference_evaluator():
    from oracle.gen_golden_pair_nms import make_preds
    from rlipv2_b200.pair_nms import triplet_nms_filter
    gold = json.load(open(os.path.join(GOLDEN, "pair_nms.json")))
    for name, g in gold.items():
        thres, alpha, beta = g["settings"]
        got = triplet_nms_filter(make_preds(), thres, alpha, beta)
        assert len(got) == len(g["kept"])
        for img, want in zip(got, g["kept"]):
            mine = [[h["subject_id"], h["object_id"], h["category_id"], h["score"]] for h in img["hoi_prediction"]]
            assert mine == want, name
        assert any(len(k) < 60 for k in g["kept"])                   # the fixture really suppresses something


def test_greedy_rule_and_edge_cases():
    from rlipv2_b200.pair_nms import pairwise_nms, triplet_nms_filter
    rng = np.random.RandomState(1)
    subs = np.concatenate([rng.rand(30, 2) * 50, rng.rand(30, 2) * 50 + 60], 1)
    objs = np.concatenate([rng.rand(30, 2) * 50, rng.rand(30, 2) * 50 + 60], 1)
    scores = rng.rand(30)
    keep = pairwise_nms(subs, objs, scores, 0.5, 1.0, 0.5)

    def iou(a, b):
        w = max(0.0, min(a[2], b[2]) - max(a[0], b[0]) + 1)
        h = max(0.0, min(a[3], b[3]) - max(a[1], b[1]) + 1)
        area = lambda r: (r[2] - r[0] + 1) * (r[3] - r[1] + 1)
        return w * h / (area(a) + area(b) - w * h)

    want = []
    for i in np.argsort(scores)[::-1]:                                # the textbook form of the same rule
        if all(iou(subs[i], subs[k]) ** 1.0 * iou(objs[i], objs[k]) ** 0.5 <= 0.5 for k in want):
            want.append(i)
    assert list(keep) == want and 0 < len(keep) < 30
    assert pairwise_nms(np.zeros((0, 4)), np.zeros((0, 4)), np.zeros(0)) == []
    assert list(pairwise_nms(subs[:1], objs[:1], scores[:1])) == [0]
    empty = triplet_nms_filter([{"filename": "x", "predictions": [], "hoi_prediction": []}])
    assert empty == [{"filename": "x", "predictions": [], "hoi_prediction": []}]
