"""TEST INFRASTRUCTURE - generator of tests/golden/parseda_postprocess.npz.

Runs HERE (build container): imports the reference's own PostProcessHOI / PostProcessSGG
(/root/reference/models/hoi.py:4769-4938) through oracle/ref_import.py, feeds them seeded synthetic
model outputs and stores inputs + per-image results (plain, temperature, zero-shot and SGG variants).

    python oracle/gen_golden_postprocess.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "parseda_postprocess.npz")


def make_outputs(seed=5, bs=3, Q=12, n_obj=9, n_verb=7):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    out = {"pred_obj_logits": r(bs, Q, n_obj + 1) * 2, "pred_sub_logits": r(bs, Q, n_obj + 1) * 2,
           "pred_verb_logits": r(bs, Q, n_verb) * 2,
           "pred_sub_boxes": torch.rand(bs, Q, 4, generator=g) * 0.5 + 0.2,
           "pred_obj_boxes": torch.rand(bs, Q, 4, generator=g) * 0.5 + 0.2}
    out["pred_sub_logits"][:, ::2, 0] += 6.0            # about half of the subjects are class 0 ("person")
    sizes = torch.tensor([[480, 640], [427, 640], [600, 393]])
    return out, sizes


def main():
    ref_import.install()
    with ref_import.chdir(ref_import.REF):
        import models.hoi as hoi
        variants = {
            "hoi": hoi.PostProcessHOI(0, sigmoid=True),
            "hoi_tem": hoi.PostProcessHOI(0, sigmoid=True, temperature=True),
            "hoi_zs": hoi.PostProcessHOI(0, sigmoid=True, zero_shot_hoi_eval=True),
            "hoi_nosig": hoi.PostProcessHOI(3, sigmoid=False),
            "sgg": hoi.PostProcessSGG(sigmoid=True),
        }
    outputs, sizes = make_outputs()
    blob = {"in_" + k: v.numpy() for k, v in outputs.items()}
    blob["in_sizes"] = sizes.numpy()
    for name, pp in variants.items():
        res = pp({k: v.clone() for k, v in outputs.items()}, sizes)
        for b, r in enumerate(res):
            for k, v in r.items():
                blob[f"{name}/{b}/{k}"] = v.numpy()
    np.savez_compressed(OUT, **blob)
    print("written", OUT, len(blob), "arrays")


if __name__ == "__main__":
    main()
