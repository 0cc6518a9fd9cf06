"""build_model(args) -> (model, criterion, postprocessors): the `--RLIP_ParSeDA_v2` branch (the hot path) and the
`--RLIP_ParSe` branch (BASELINE config 1, the CPU plumbing case) of
/root/reference/models/detr.py:320-701 (reached through models/__init__.py:15-16).

`args` is the reference's own argparse namespace (main.py:38-491), so main.py / engine.py can call
this in place of `models.build_model` unchanged.  Flags that select other model families raise
NotImplementedError: they are outside the ParSeDA hot path.
"""
import math

import torch

from .backbone import build_backbone
from .criterion import SetCriterionHOI
from .matcher import build_matcher
from .parseda import RLIP_ParSeDA
from .parseda_transformer import build_parseda_transformer
from .postprocess import build_postprocessors


def build_weight_dict(args):
    """detr.py:571-620 - every key the reference registers (unused ones included, engine.py:108
    only sums keys that appear in the loss dict)."""
    w = {
        "loss_obj_ce": args.obj_loss_coef, "loss_verb_ce": args.verb_loss_coef,
        "loss_sub_bbox": args.bbox_loss_coef, "loss_obj_bbox": args.bbox_loss_coef,
        "loss_sub_giou": args.giou_loss_coef, "loss_obj_giou": args.giou_loss_coef,
        "loss_entropy_bound": args.entropy_bound_coef, "loss_kl_divergence": args.kl_divergence_coef,
        "loss_verb_gt_recon": args.verb_gt_recon_coef, "loss_ranking_verbs": args.ranking_verb_coef,
        "loss_verb_hm": args.verb_hm_coef, "loss_semantic_similar": args.semantic_similar_coef,
        "loss_verb_threshold": args.verb_threshold_coef,
        "loss_sub_matching": args.obj_loss_coef, "loss_obj_matching": args.obj_loss_coef,
        "loss_verb_matching": args.verb_loss_coef, "loss_masked_recon": args.masked_loss_coef,
        "loss_masked_ce": args.masked_loss_coef, "loss_obj_ce_recon": args.obj_loss_coef,
        "loss_sub_bbox_recon": args.bbox_loss_coef, "loss_obj_bbox_recon": args.bbox_loss_coef,
        "loss_sub_giou_recon": args.giou_loss_coef, "loss_obj_giou_recon": args.giou_loss_coef,
    }
    exponential = ["loss_sub_bbox", "loss_obj_bbox", "loss_sub_giou", "loss_obj_giou", "loss_obj_ce", "loss_verb_ce"]
    if args.aux_loss:
        aux = {}
        for i in range(args.dec_layers - 1):
            for k, v in w.items():
                if args.exponential_loss and k in exponential:
                    v = math.pow(args.exponential_hyper, args.dec_layers - 1 - i) * v
                aux[k + f"_{i}"] = v
        w.update(aux)
    return w


def _build_parse(args):
    """detr.py:330-332 (vanilla single-level backbone, models/backbone.py:174-181) + :402-414"""
    from .backbone import Backbone, Joiner, PositionEmbeddingSine
    from .parse_detr import RLIP_ParSe, build_parse_transformer
    if getattr(args, "position_embedding", "sine") not in ("v2", "sine"):
        raise NotImplementedError("the RLIP scripts use the sine position embedding")
    backbone = Joiner(Backbone(args.backbone, args.lr_backbone > 0, bool(args.masks), args.dilation,
                               getattr(args, "backbone_weights", None)),
                      PositionEmbeddingSine(args.hidden_dim // 2, normalize=True))
    cross_modal = args.verb_loss_type == "cross_modal_matching" and args.obj_loss_type == "cross_modal_matching"
    return RLIP_ParSe(backbone, build_parse_transformer(args), num_queries=args.num_queries,
                      contrastive_align_loss=cross_modal, contrastive_hdim=64, aux_loss=args.aux_loss,
                      subject_class=args.subject_class, use_no_verb_token=getattr(args, "use_no_verb_token", False), args=args)


def build_model(args):
    parse = bool(getattr(args, "RLIP_ParSe", False))
    if not (getattr(args, "RLIP_ParSeDA_v2", False) or parse):
        raise NotImplementedError("rlipv2_b200 implements the --RLIP_ParSeDA_v2 and --RLIP_ParSe models only")
    if not (args.hoi or args.sgg or getattr(args, "cross_modal_pretrain", False)):
        raise NotImplementedError("the RLIP models run with --hoi, --sgg or --cross_modal_pretrain")
    device = torch.device(args.device)
    matcher = build_matcher(args)
    if parse:
        model = _build_parse(args)
    else:
        backbone = build_backbone(args)
        transformer = build_parseda_transformer(args)
        model = RLIP_ParSeDA(backbone, transformer, num_queries=args.num_queries,
                             num_feature_levels=args.num_feature_levels, aux_loss=args.aux_loss,
                             with_box_refine=args.with_box_refine, two_stage=args.two_stage, use_dab=True,
                             num_patterns=args.num_patterns, random_refpoints_xy=args.random_refpoints_xy,
                             subject_class=args.subject_class, pseudo_verb=getattr(args, "pseudo_verb", False), args=args)
    losses = ["obj_labels", "verb_labels", "sub_obj_boxes", "obj_cardinality"]
    for flag in ("entropy_bound", "kl_divergence", "verb_gt_recon", "ranking_verb", "no_verb_bce_focal", "verb_hm",
                 "semantic_similar", "verb_threshold", "masked_entity_modeling", "verb_tagger"):
        if getattr(args, flag, False):
            raise NotImplementedError(f"--{flag} is not used by the ParSeDA scripts (out of scope)")
    criterion = SetCriterionHOI(
        args.num_obj_classes, args.num_queries, args.num_verb_classes, matcher=matcher,
        weight_dict=build_weight_dict(args), eos_coef=args.eos_coef, losses=losses,
        verb_loss_type=args.verb_loss_type, obj_loss_type=args.obj_loss_type,
        matching_symmetric=getattr(args, "matching_symmetric", True), RLIP_ParSe=getattr(args, "RLIP_ParSe", False),
        subject_class=args.subject_class, use_no_verb_token=getattr(args, "use_no_verb_token", False),
        giou_verb_label=getattr(args, "giou_verb_label", False), verb_curing=getattr(args, "verb_curing", False),
        pseudo_verb=getattr(args, "pseudo_verb", False), triplet_filtering=getattr(args, "triplet_filtering", False),
        naive_obj_smooth=getattr(args, "naive_obj_smooth", 0), naive_verb_smooth=getattr(args, "naive_verb_smooth", 0),
        args=args)
    criterion.to(device)
    postprocessors = build_postprocessors(args)      # detr.py:683-691
    return model, criterion, postprocessors


def default_args(**overrides):
    """Namespace with the defaults of main.py:38-491 that the ParSeDA branch reads, set to the values
    of scripts/RLIP_ParSeDA/fine_tune_RLIP_ParSeDA_v2_hico.sh:17-59.  For programmatic use (bench,
    tests) when the reference's argparse is not importable."""
    import argparse
    d = dict(
        device="cuda", hoi=True, sgg=False, cross_modal_pretrain=False, RLIP_ParSeDA_v2=True,
        backbone="resnet50", dilation=False, position_embedding="sine", lr_backbone=1.41e-5, masks=False,
        hidden_dim=256, nheads=8, enc_layers=6, dec_layers=3, dim_feedforward=2048, dropout=0.0,
        num_feature_levels=4, dec_n_points=4, enc_n_points=4, two_stage=False, num_queries=128,
        with_box_refine=True, num_patterns=0, random_refpoints_xy=False, aux_loss=True,
        fusion_type="GLIP_attn", fusion_interval=2, fusion_last_vis=True, lang_aux_loss=True,
        gating_mechanism="VXAc", verb_query_tgt_type="vanilla_MBF", subject_class=True,
        use_no_obj_token=True, giou_verb_label=True, pseudo_verb=False,
        stable_softmax_2d=False, clamp_min_for_underflow=False, clamp_max_for_overflow=False,
        separate_bidirectional=False, do_lang_proj_outside_checkpoint=False, use_checkpoint_fusion=False,
        text_encoder_type="roberta-base", freeze_text_encoder=False, synthetic_text_encoder=None,
        set_cost_obj_class=1, set_cost_verb_class=1, set_cost_bbox=2.5, set_cost_giou=1,
        obj_loss_coef=1, verb_loss_coef=1, bbox_loss_coef=2.5, giou_loss_coef=1, eos_coef=0.1,
        entropy_bound_coef=0.01, kl_divergence_coef=0.01, verb_gt_recon_coef=1, ranking_verb_coef=1,
        verb_hm_coef=1, semantic_similar_coef=1, verb_threshold_coef=1, masked_loss_coef=1,
        exponential_loss=False, exponential_hyper=0.8, num_obj_classes=80, num_verb_classes=117,
        obj_loss_type="cross_entropy", verb_loss_type="focal", use_no_verb_token=False,
        verb_curing=False, triplet_filtering=False, naive_obj_smooth=0, naive_verb_smooth=0, verb_tagger=False,
        matching_symmetric=True, RLIP_ParSe=False,
    )
    d.update(overrides)
    return argparse.Namespace(**d)
