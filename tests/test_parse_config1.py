"""BASELINE config 1: RLIP-ParSe (plain DETR) R50, forward + matcher on two synthetic 480x640
images + 32 label strings, CPU, world size 1 - the family's plumbing case.

Checker: tests/golden/parse_config1.npz, produced by the reference's OWN `RLIP_ParSe` + `SetCriterionHOI` +
`HungarianMatcherHOI` (oracle/gen_golden_parse.py, same name-keyed weights from oracle/detfill.py, eval mode, fp32).
Tolerance: 1e-3 relative (north_star); matcher indices bit-exact."""
import argparse
import json
import os

import numpy as np
import torch

from oracle.detfill import det_fill_
from tests.golden_util import GOLDEN

OBJ_NAMES = [f"object kind {i}" for i in range(15)] + ["no objects"]
VERB_NAMES = [f"relation {i} with" for i in range(16)]


def _args():
    """the flags of config 1 over main.py's defaults (main.py:38-491)"""
    from rlipv2_b200 import models
    d = vars(models.default_args(device="cpu", synthetic_text_encoder=True))
    d.update(RLIP_ParSe=True, RLIP_ParSeDA_v2=False, num_queries=100, dropout=0.1, pre_norm=False, pass_pos_and_query=True,
             giou_verb_label=False, use_no_obj_token=True, subject_class=True, masks=False, lr_backbone=1e-5,
             set_cost_bbox=2.5, set_cost_giou=1, bbox_loss_coef=2.5, giou_loss_coef=1, num_obj_classes=80,
             num_verb_classes=117, eos_coef=0.1)
    return argparse.Namespace(**d)


def _inputs():
    g = torch.Generator().manual_seed(21)                     # oracle/gen_golden_parse.py::make_inputs
    imgs = [torch.randn(3, 480, 640, generator=g), torch.randn(3, 480, 640, generator=g)]
    targets = []
    for k in (3, 2):
        box = lambda: torch.cat([torch.rand(k, 2, generator=g) * 0.4 + 0.3, torch.rand(k, 2, generator=g) * 0.2 + 0.1], 1)
        verbs = torch.zeros(k, len(VERB_NAMES))
        verbs[torch.arange(k), torch.randint(0, len(VERB_NAMES), (k,), generator=g)] = 1
        targets.append({"obj_labels": torch.randint(0, len(OBJ_NAMES) - 1, (k,), generator=g),
                        "sub_labels": torch.zeros(k, dtype=torch.long), "verb_labels": verbs,
                        "sub_boxes": box(), "obj_boxes": box()})
    return imgs, targets, [(OBJ_NAMES, VERB_NAMES)]


def test_parse_state_dict_keys_match_reference():
    from rlipv2_b200 import models
    model, _, _ = models.build_model(_args())
    ref = json.load(open(os.path.join(GOLDEN, "parse_state_dict_keys.json")))
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert mine == ref and len(mine) == 671


def test_config1_forward_matcher_losses_match_reference():
    from rlipv2_b200 import dense, models
    from rlipv2_b200.nested import nested_tensor_from_tensor_list
    dense.set_matmul_precision("fp32")
    with np.load(os.path.join(GOLDEN, "parse_config1.npz")) as z:
        g = {k: z[k] for k in z.files}
    model, criterion, _ = models.build_model(_args())
    det_fill_(model, seed=1)
    model.eval()
    criterion.eval()
    imgs, targets, text = _inputs()
    np.testing.assert_allclose([float(imgs[0].double().sum()), float(imgs[1].double().abs().sum())], g["img_checksum"],
                               rtol=1e-12)
    for t_i, t in enumerate(targets):
        for k, v in t.items():
            np.testing.assert_array_equal(v.numpy(), g[f"tgt{t_i}_{k}"])
    samples = nested_tensor_from_tensor_list(imgs)
    with torch.no_grad():
        cache = model(samples, encode_and_save=True, text=text, targets=targets)
        out = model(samples, encode_and_save=False, memory_cache=cache, text=text, targets=targets)
        losses = criterion(out, targets)
        indices = criterion.matcher({k: v for k, v in out.items() if k != "aux_outputs"}, targets)
        aux_indices = [criterion.matcher(a, targets) for a in out["aux_outputs"]]
    # phase A
    np.testing.assert_array_equal(cache["text_attention_mask"].numpy(), g["text_attention_mask"])
    assert 0 < int(g["text_attention_mask"].sum()) < g["text_attention_mask"].size       # quirk 4 is exercised both ways
    np.testing.assert_array_equal(cache["mask"].numpy(), g["mask"])
    close = lambda a, b, name: np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-4 * float(np.abs(b).max()), err_msg=name)
    close(cache["text_memory_resized"].numpy(), g["text_memory_resized"], "text_memory_resized")
    close(cache["text_memory"].numpy(), g["text_memory"], "text_memory")
    close(cache["img_memory"][::7, :, ::8].numpy(), g["img_memory_slice"], "img_memory")
    # phase B: every decoder level
    for k in ("pred_sub_logits", "pred_obj_logits", "pred_verb_logits", "pred_sub_boxes", "pred_obj_boxes"):
        close(out[k].numpy(), g["out_" + k], k)
        for i, a in enumerate(out["aux_outputs"]):
            close(a[k].numpy(), g[f"aux{i}_" + k], f"aux{i} {k}")
    assert out["pred_obj_logits"].shape == (2, 100, 16) and out["pred_verb_logits"].shape == (2, 100, 16)
    # matcher: bit-exact indices, all three decoder levels
    for li, ind in enumerate([indices] + aux_indices):
        for b, (i, j) in enumerate(ind):
            np.testing.assert_array_equal(i.numpy(), g[f"match{li}_{b}_i"])
            np.testing.assert_array_equal(j.numpy(), g[f"match{li}_{b}_j"])
    # every loss / meter the reference reports
    ref_keys = sorted(k[len("loss_"):] for k in g if k.startswith("loss_"))
    assert sorted(losses.keys()) == ref_keys
    for k in ref_keys:
        np.testing.assert_allclose(float(losses[k]), float(g["loss_" + k]), rtol=1e-3, atol=1e-4, err_msg=k)
