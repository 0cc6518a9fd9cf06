"""Golden fixture for rlipv2_b200/pair_nms.py from the reference's own `HICOEvaluator.triplet_nms_filter`
(/root/reference/datasets/hico_eval.py:493-564), called unbound on a stand-in `self` that carries the three NMS settings.
    python oracle/gen_golden_pair_nms.py   -> tests/golden/pair_nms.json"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_preds(seed=0, n_img=4, n_box=12, n_hoi=60):
    rng = np.random.RandomState(seed)
    preds = []
    for i in range(n_img):
        centres = rng.rand(4, 2) * 300 + 100                       # a few clusters, so that pairs really overlap
        boxes = []
        for b in range(n_box):
            c = centres[rng.randint(4)] + rng.randn(2) * 6
            wh = 80 + rng.rand(2) * 20
            boxes.append({"bbox": [float(c[0] - wh[0] / 2), float(c[1] - wh[1] / 2), float(c[0] + wh[0] / 2), float(c[1] + wh[1] / 2)],
                          "category_id": int(rng.randint(3))})
        hois = [{"subject_id": int(rng.randint(n_box)), "object_id": int(rng.randint(n_box)), "category_id": int(rng.randint(2)),
                 "score": float(np.round(rng.rand(), 2 if i % 2 else 6))} for _ in range(n_hoi)]   # odd images: score ties
        preds.append({"filename": f"img{i}.jpg", "predictions": boxes, "hoi_prediction": hois})
    return preds


def main():
    sys.path.insert(0, ROOT)
    from oracle import ref_import
    ref_import.install()                                           # stubs for pycocotools / matplotlib etc.
    sys.path.insert(0, "/root/reference")
    from datasets.hico_eval import HICOEvaluator
    out = {}
    for name, (thres, alpha, beta) in {"default": (0.7, 1.0, 0.5), "tight": (0.3, 1.0, 1.0)}.items():
        me = types.SimpleNamespace(thres_nms=thres, nms_alpha=alpha, nms_beta=beta)
        me.pairwise_nms = types.MethodType(HICOEvaluator.pairwise_nms, me)
        preds = make_preds()
        got = HICOEvaluator.triplet_nms_filter(me, preds)
        out[name] = {"settings": [thres, alpha, beta],
                     "kept": [[[h["subject_id"], h["object_id"], h["category_id"], h["score"]] for h in g["hoi_prediction"]] for g in got]}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "pair_nms.json"), "w"))
    print({k: [len(x) for x in v["kept"]] for k, v in out.items()})


if __name__ == "__main__":
    main()
