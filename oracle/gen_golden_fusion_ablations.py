"""TEST INFRASTRUCTURE - generator of tests/golden/parseda_mdetr_step.npz and parseda_nofusion_step.npz.

Runs HERE (build container).  Two published ParSeDA scripts (scripts/RLIP_ParSeDA/fine_tune_RLIP_ParSeDA_v2_hico_MDETR.sh,
train_RLIP_ParSeDA_v2_mixed_vgcoco_resnet_MDETR.sh) train the paper's late-fusion ablation `--fusion_type MDETR_attn`
(/root/reference/models/dab_deformable/deformable_transformer.py:252-256, 278-291, 552-562, 703-733); `no_fusion` is
main.py's default (main.py:203).  These fixtures pin both to the reference's own modules: name-keyed weights
(oracle/detfill.py), eval mode, fp32, the inputs of parseda_step.npz.

    python oracle/gen_golden_fusion_ablations.py
"""
import json
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
GRAD_KEYS = {
    "MDETR_attn": ["transformer.obj_fusion.layers.0.self_attn.in_proj_weight", "transformer.verb_fusion.layers.2.linear2.bias",
                   "transformer.obj_fusion.norm.weight", "transformer.encoder.layers.3.linear1.bias",
                   "transformer.resizer.fc.bias", "projection_text.bias", "transformer.text_encoder.pooler.dense.bias"],
    "no_fusion": ["transformer.encoder.layers.3.linear1.bias", "transformer.resizer.fc.bias", "projection_text.bias",
                  "transformer.verb_decoder.layers.2.norm3.weight"],
}


def flags(fusion_type):
    f = list(ref_import.PARSEDA_FLAGS)
    f[f.index("--fusion_type") + 1] = fusion_type
    return f + ["--num_queries", "16"]


def main():
    ref_import.install()
    keys = {}
    for fusion_type, name in (("MDETR_attn", "mdetr"), ("no_fusion", "nofusion")):
        args = ref_import.parse_args(flags(fusion_type))
        with ref_import.chdir(ref_import.REF):
            from models import build_model
            from util.misc import nested_tensor_from_tensor_list
            model, criterion, _ = build_model(args)
        keys[fusion_type] = {k: list(v.shape) for k, v in model.state_dict().items()}
        det_fill_(model, seed=3)
        model.eval()
        criterion.eval()
        imgs, targets, text = make_step_inputs()
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
        save = {"total_loss": total.detach().numpy(), "n_aux": np.asarray(len(out["aux_outputs"]))}
        for k in ("pred_sub_logits", "pred_obj_logits", "pred_verb_logits", "pred_sub_boxes", "pred_obj_boxes"):
            save["out_" + k] = out[k].detach().numpy()
            save["aux0_" + k] = out["aux_outputs"][0][k].detach().numpy()
        for k, v in loss_dict.items():
            save["loss_" + k] = np.asarray(float(v))
        for b, (i, j) in enumerate(indices):
            save[f"match_{b}_i"], save[f"match_{b}_j"] = i.numpy(), j.numpy()
        params = model.state_dict(keep_vars=True)
        for k in GRAD_KEYS[fusion_type]:
            save["gradnorm_" + k] = np.asarray(float(params[k].grad.norm()))
        np.savez_compressed(os.path.join(OUT, f"parseda_{name}_step.npz"), **save)
        print(fusion_type, "fixture written; total", float(total), "keys", len(keys[fusion_type]))
    json.dump(keys, open(os.path.join(OUT, "parseda_fusion_ablation_keys.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
