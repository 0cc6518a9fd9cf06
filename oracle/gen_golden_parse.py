"""TEST INFRASTRUCTURE - generator of the BASELINE config 1 fixture tests/golden/parse_config1.npz.

Runs HERE (build container): builds the reference's own RLIP_ParSe (models/hoi.py:2259, flags of SURVEY.md section 8d
"Config 1") through the shim set of oracle/ref_import.py, fills it with the deterministic name-keyed weights of
oracle/detfill.py, and runs phase A + phase B in eval() + criterion.matcher + the losses on CPU:
2 images randn(3, 480, 640) (no padding), 15 object strings + 'no objects' + 16 verb strings = 32 labels,
3 and 2 target triplets, 100 queries.

    python oracle/gen_golden_parse.py
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

OUT = os.path.join(ROOT, "tests", "golden")
CONFIG1_FLAGS = ["--RLIP_ParSe", "--hoi", "--backbone", "resnet50", "--enc_layers", "6", "--dec_layers", "3",
                 "--num_queries", "100", "--subject_class", "--use_no_obj_token", "--obj_loss_type", "cross_entropy",
                 "--verb_loss_type", "focal"]
OBJ_NAMES = [f"object kind {i}" for i in range(15)] + ["no objects"]
VERB_NAMES = [f"relation {i} with" for i in range(16)]


def make_inputs():
    g = torch.Generator().manual_seed(21)
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


def main():
    ref_import.install()
    args = ref_import.parse_args(CONFIG1_FLAGS)
    with ref_import.chdir(ref_import.REF):
        from models import build_model
        model, criterion, _ = build_model(args)
    json.dump({k: list(v.shape) for k, v in model.state_dict().items()},
              open(os.path.join(OUT, "parse_state_dict_keys.json"), "w"), indent=0)
    det_fill_(model, seed=1)          # seed 1: 11 of the 32 labels pool to a positive sum (quirk 4), margin 0.34
    model.eval()
    criterion.eval()
    imgs, targets, text = make_inputs()
    from util.misc import nested_tensor_from_tensor_list
    samples = nested_tensor_from_tensor_list(imgs)
    with torch.no_grad():
        cache = model(samples, encode_and_save=True, text=text, targets=targets)
        out = model(samples, encode_and_save=False, memory_cache=cache, text=text, targets=targets)
        loss_dict = criterion(out, targets)
        indices = criterion.matcher({k: v for k, v in out.items() if k != "aux_outputs"}, targets)
        aux_indices = [criterion.matcher(a, targets) for a in out["aux_outputs"]]
    print("labels treated as padding:", int(cache["text_attention_mask"].sum()), "of", cache["text_attention_mask"].numel())
    # the images are regenerated from the seed by the test (make_inputs); only a checksum is stored
    save = {"img_checksum": np.asarray([float(imgs[0].double().sum()), float(imgs[1].double().abs().sum())]),
            "img_memory_slice": cache["img_memory"][::7, :, ::8].numpy(),
            "text_memory": cache["text_memory"].numpy(),
            "text_memory_resized": cache["text_memory_resized"].numpy(),
            "text_attention_mask": cache["text_attention_mask"].numpy(),
            "mask": cache["mask"].numpy()}
    for k in ("pred_sub_logits", "pred_obj_logits", "pred_verb_logits", "pred_sub_boxes", "pred_obj_boxes"):
        save["out_" + k] = out[k].numpy()
        for i, a in enumerate(out["aux_outputs"]):
            save[f"aux{i}_" + k] = a[k].numpy()
    for k, v in loss_dict.items():
        save["loss_" + k] = np.asarray(float(v))
    for li, ind in enumerate([indices] + aux_indices):
        for b, (i, j) in enumerate(ind):
            save[f"match{li}_{b}_i"] = i.numpy()
            save[f"match{li}_{b}_j"] = j.numpy()
    for t_i, t in enumerate(targets):
        for k, v in t.items():
            save[f"tgt{t_i}_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "parse_config1.npz"), **save)
    print("config-1 fixture written;", {k: round(float(v), 5) for k, v in loss_dict.items() if k.startswith("loss_")})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    main()
