"""TEST INFRASTRUCTURE - generator of tests/golden/parseda_pretrain_step.npz (BASELINE config 3's model flags).

Runs HERE (build container).  The relational pre-training scripts
(/root/reference/scripts/RLIP_ParSeDA/train_RLIP_ParSeDA_v2_mixed_vgcocoo365_resnet.sh) build the same RLIP_ParSeDA as
the fine-tune step but with `--cross_modal_pretrain --pseudo_verb` instead of `--hoi`: the model additionally emits
`target_verb_sim` (pseudo relation labels from distances between the pre-fusion relation text embeddings,
models/hoi.py:2197-2239) and the verb loss adds them to the matched targets (hoi.py:3925-4028).  This fixture pins that
path to the reference's own modules: name-keyed weights (oracle/detfill.py), eval mode, fp32, 2 images, 16 queries.

    python oracle/gen_golden_pretrain.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from oracle.detfill import det_fill_  # noqa: E402
from oracle.gen_golden_model import make_step_inputs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
PRETRAIN_FLAGS = [f for f in ref_import.PARSEDA_FLAGS if f != "--hoi"] + ["--cross_modal_pretrain", "--pseudo_verb"]


def main():
    ref_import.install()
    args = ref_import.parse_args(PRETRAIN_FLAGS + ["--num_queries", "16"])
    with ref_import.chdir(ref_import.REF):
        from models import build_model
        model, criterion, _ = build_model(args)
    det_fill_(model, seed=3)
    model.eval()
    criterion.eval()
    imgs, targets, text = make_step_inputs()
    from util.misc import nested_tensor_from_tensor_list
    samples = nested_tensor_from_tensor_list(imgs)
    for p in model.parameters():
        p.requires_grad_(True)
    cache = model(samples, encode_and_save=True, text=text, targets=targets)
    out = model(samples, encode_and_save=False, memory_cache=cache, text=text, targets=targets)
    loss_dict = criterion(out, targets)
    wd = criterion.weight_dict
    total = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
    total.backward()
    indices = criterion.matcher({k: v for k, v in out.items() if k != "aux_outputs"}, targets)
    save = {"target_verb_sim": out["target_verb_sim"].detach().numpy(), "total_loss": total.detach().numpy()}
    for k in ("pred_sub_logits", "pred_obj_logits", "pred_verb_logits", "pred_sub_boxes", "pred_obj_boxes"):
        save["out_" + k] = out[k].detach().numpy()
    for k, v in loss_dict.items():
        save["loss_" + k] = np.asarray(float(v))
    for b, (i, j) in enumerate(indices):
        save[f"match_{b}_i"], save[f"match_{b}_j"] = i.numpy(), j.numpy()
    params = model.state_dict(keep_vars=True)
    for k in ("projection_text.bias", "bias_pred_a", "transformer.verb_decoder.layers.2.norm3.weight",
              "transformer.encoder.roberta_layers.2.output.LayerNorm.weight", "transformer.resizer.fc.bias"):
        save["gradnorm_" + k] = np.asarray(float(params[k].grad.norm()))
    np.savez_compressed(os.path.join(OUT, "parseda_pretrain_step.npz"), **save)
    print("pretrain fixture written; total", float(total), "nonzero pseudo labels", int((out["target_verb_sim"] > 0).sum()),
          "of", out["target_verb_sim"].numel())


if __name__ == "__main__":
    main()
