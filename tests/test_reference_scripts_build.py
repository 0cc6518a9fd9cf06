"""Drop-in at the level a user meets first: every published ParSeDA launch script of the reference
(/root/reference/scripts/RLIP_ParSeDA/*.sh - HICO-DET / V-COCO / OI-SGG fine-tuning, relational pre-training on VG / COCO /
Objects365 mixes, zero-shot, few-shot, UC splits; ResNet-50, Swin-T, Swin-L; ALIF and the MDETR-style late-fusion
ablation) is parsed with the reference's own argparse (main.py:38-491) and handed to this repo's `build_model`.
Scripts that shape the model identically are built once.  Only runs where /root/reference exists."""
import glob
import os
import shlex

import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")

SHAPING = ("backbone", "num_queries", "fusion_type", "hoi", "sgg", "cross_modal_pretrain", "pseudo_verb", "subject_class",
           "use_no_obj_token", "verb_query_tgt_type", "fusion_interval", "fusion_last_vis", "lang_aux_loss", "giou_verb_label",
           "gating_mechanism", "enc_layers", "dec_layers", "dim_feedforward", "num_feature_levels", "with_box_refine",
           "obj_loss_type", "verb_loss_type", "zero_shot_eval", "verb_curing", "dropout")


def _script_flags(path):
    lines = [l for l in open(path).read().splitlines() if not l.strip().startswith("#")]
    body = " ".join(l.rstrip("\\").strip() for l in lines)
    i = body.find(" main.py")
    return None if i < 0 else shlex.split(body[i + len(" main.py"):])


def test_every_published_parseda_script_builds_here():
    from rlipv2_b200 import models
    scripts = sorted(glob.glob(os.path.join(ref_import.REF, "scripts", "RLIP_ParSeDA", "*.sh")))
    assert len(scripts) >= 35
    seen, built, skipped = {}, 0, []
    for path in scripts:
        name = os.path.basename(path)
        flags = _script_flags(path)
        if flags is None:                                   # test_vcoco_official.sh runs generate_vcoco_official.py
            skipped.append(name)
            continue
        args = ref_import.parse_args(flags)
        assert args.RLIP_ParSeDA_v2, name
        args.synthetic_text_encoder = True
        sig = tuple(str(getattr(args, k, None)) for k in SHAPING)
        if sig in seen:
            continue
        args.device = "meta"                               # shapes and structure only: no parameter initialisation
        with torch.device("meta"):
            model, criterion, post = models.build_model(args)
        seen[sig] = name
        built += 1
        names = [n for n, _ in model.named_parameters()]
        assert any("backbone" in n for n in names) and any("text_encoder" in n for n in names), name
        assert criterion.weight_dict and hasattr(criterion, "matcher"), name
        if args.hoi:
            assert "hoi" in post, name
        if args.sgg:
            assert "sgg" in post, name
        if "swin" in args.backbone:
            assert model.backbone.num_channels[-1] == {"swin_tiny": 768, "swin_large": 1536}[args.backbone], name
        if args.fusion_type == "MDETR_attn":
            assert hasattr(model.transformer, "obj_fusion") and hasattr(model.transformer, "verb_fusion"), name
    assert skipped == ["test_vcoco_official.sh"] and built >= 6, built
