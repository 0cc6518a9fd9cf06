"""`--fusion_type MDETR_attn` (late fusion; scripts/RLIP_ParSeDA/*_MDETR.sh) and `no_fusion` (main.py's default) of the ParSeDA
transformer (/root/reference/models/dab_deformable/deformable_transformer.py:252-256, 278-291, 552-562, 703-733) against
fixtures from the reference's own modules (oracle/gen_golden_fusion_ablations.py): identical state_dict keys, outputs of
the last and the first decoder level, matcher indices (exact), every loss, gradient norms.  Tolerance 1e-3 (north_star)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.detfill import det_fill_
from tests.golden_util import GOLDEN


def _load(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("fusion_type,fixture", [("MDETR_attn", "parseda_mdetr_step.npz"), ("no_fusion", "parseda_nofusion_step.npz")])
def test_fusion_ablation_step_matches_reference(fusion_type, fixture, msda_cpu_stub):
    from oracle.gen_golden_model import OBJ_NAMES, VERB_NAMES
    from rlipv2_b200 import dense, models
    dense.set_matmul_precision("fp32")
    g, gs = _load(fixture), _load("parseda_step.npz")
    model, criterion, _ = models.build_model(models.default_args(device="cpu", num_queries=16, synthetic_text_encoder=True,
                                                                 fusion_type=fusion_type))
    ref_keys = json.load(open(os.path.join(GOLDEN, "parseda_fusion_ablation_keys.json")))[fusion_type]
    assert {k: list(v.shape) for k, v in model.state_dict().items()} == ref_keys
    det_fill_(model, seed=3)
    model.eval()
    criterion.eval()
    imgs = [torch.from_numpy(gs["img0"]), torch.from_numpy(gs["img1"])]
    targets = [{k: torch.from_numpy(gs[f"tgt{i}_{k}"]) for k in ("obj_labels", "sub_labels", "verb_labels", "sub_boxes", "obj_boxes")}
               for i in range(2)]
    text = [(OBJ_NAMES, VERB_NAMES)]
    cache = model(imgs, encode_and_save=True, text=text, targets=targets)
    out = model(imgs, encode_and_save=False, memory_cache=cache, text=text, targets=targets)
    loss_dict = criterion(out, targets)
    wd = criterion.weight_dict
    total = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
    total.backward()
    c = lambda t: t.detach().numpy()
    tol = dict(rtol=1e-3, atol=2e-4)
    assert len(out["aux_outputs"]) == int(g["n_aux"])
    for k in ("pred_sub_logits", "pred_obj_logits", "pred_verb_logits", "pred_sub_boxes", "pred_obj_boxes"):
        np.testing.assert_allclose(c(out[k]), g["out_" + k], err_msg=k, **tol)
        np.testing.assert_allclose(c(out["aux_outputs"][0][k]), g["aux0_" + k], err_msg="aux " + k, **tol)
    for b, (i, j) in enumerate(criterion.matcher({k: v for k, v in out.items() if k != "aux_outputs"}, targets)):
        np.testing.assert_array_equal(i.numpy(), g[f"match_{b}_i"])
        np.testing.assert_array_equal(j.numpy(), g[f"match_{b}_j"])
    assert sorted(loss_dict.keys()) == sorted(k[len("loss_"):] for k in g if k.startswith("loss_"))
    for k, v in loss_dict.items():
        np.testing.assert_allclose(float(v), float(g["loss_" + k]), rtol=1e-3, atol=1e-4, err_msg=k)
    np.testing.assert_allclose(float(total), float(g["total_loss"]), rtol=1e-3)
    params = model.state_dict(keep_vars=True)
    for k in [k[len("gradnorm_"):] for k in g if k.startswith("gradnorm_")]:
        np.testing.assert_allclose(float(params[k].grad.norm()), float(g["gradnorm_" + k]), rtol=3e-3, err_msg=k)
