"""Pins the functional CPU oracle of the train step (oracle/parseda_oracle.py) to the reference:
tests/golden/parseda_step.npz was produced by the reference's own RLIP_ParSeDA + SetCriterionHOI +
HungarianMatcherHOI (oracle/gen_golden_model.py).  CPU only."""
import os

import numpy as np
import torch

from oracle import parseda_oracle as po
from tests.golden_util import GOLDEN


def test_oracle_step_matches_reference():
    from oracle.gen_golden_model import GRAD_KEYS, OBJ_NAMES, VERB_NAMES
    from rlipv2_b200.nested import nested_tensor_from_tensor_list
    from rlipv2_b200.text_encoder import HashTokenizer
    with np.load(os.path.join(GOLDEN, "parseda_step.npz")) as z:
        g = {k: z[k] for k in z.files}
    W = po.synthetic_weights(16, seed=3)
    grad_keys = [k for k in GRAD_KEYS if not k.startswith("backbone.") and "text_encoder" not in k]
    for k in grad_keys:
        W[k].requires_grad_(True)
    third = po.build_third_party(W)
    third[1].eval()
    nt = nested_tensor_from_tensor_list([torch.from_numpy(g["img0"]), torch.from_numpy(g["img1"])])
    targets = [{k: torch.from_numpy(g[f"tgt{i}_{k}"]) for k in ("obj_labels", "sub_labels", "verb_labels", "sub_boxes", "obj_boxes")}
               for i in range(2)]
    out, extra = po.forward_step(W, third, nt.tensors, nt.mask, [(OBJ_NAMES, VERB_NAMES)], HashTokenizer())
    tol = dict(rtol=1e-3, atol=2e-4)
    np.testing.assert_array_equal(extra["text_attention_mask"].numpy(), g["text_attention_mask"])
    np.testing.assert_allclose(extra["img_memory"][:, ::3, ::8].detach().numpy(), g["img_memory_slice"], **tol)
    np.testing.assert_allclose(extra["text_memory_resized"].detach().numpy(), g["text_memory_resized"], **tol)
    for k in ("pred_sub_logits", "pred_obj_logits", "pred_verb_logits", "pred_sub_boxes", "pred_obj_boxes"):
        np.testing.assert_allclose(out[k].detach().numpy(), g["out_" + k], **tol)
        for i, a in enumerate(out["aux_outputs"]):
            np.testing.assert_allclose(a[k].detach().numpy(), g[f"aux{i}_" + k], **tol)
    losses, total, matches = po.criterion(out, targets)
    for li, ind in enumerate(matches):
        for b, (i, j) in enumerate(ind):
            np.testing.assert_array_equal(i.numpy(), g[f"match{li}_{b}_i"])
            np.testing.assert_array_equal(j.numpy(), g[f"match{li}_{b}_j"])
    gold = sorted(k[len("loss_"):] for k in g if k.startswith("loss_"))
    assert sorted(losses.keys()) == gold
    for k, v in losses.items():
        np.testing.assert_allclose(float(v), float(g["loss_" + k]), rtol=1e-3, atol=1e-4, err_msg=k)
    np.testing.assert_allclose(float(total), float(g["total_loss"]), rtol=1e-3)
    total.backward()
    for k in grad_keys:
        ref_norm = float(g["gradnorm_" + k])
        assert abs(float(W[k].grad.norm()) - ref_norm) <= 2e-3 * ref_norm + 1e-7, k
