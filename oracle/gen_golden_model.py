"""TEST INFRASTRUCTURE - generator of the model-level golden fixtures tests/golden/parseda_*.

Runs HERE (build container): imports the reference's own modules from /root/reference through the
shim set of oracle/ref_import.py, fills them with the deterministic name-keyed weights of
oracle/detfill.py, runs them on CPU in eval() mode (fp32) and stores inputs + outputs:

  parseda_state_dict_keys.json   key -> shape of the reference RLIP_ParSeDA state_dict (1078 entries)
  parseda_alif.npz               RLIPv2_VLFuse block (models/fuse_helper.py:983)
  parseda_roberta.npz            RobertaLayer (models/modeling_roberta.py:340)
  parseda_step.npz               full RLIP_ParSeDA two-phase forward + SetCriterionHOI + matcher
                                 + backward (selected gradients), 2 images, 16 queries, 11 labels

    python oracle/gen_golden_model.py
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

OBJ_NAMES = ["person", "bicycle", "dining table", "cup", "umbrella", "no objects"]
VERB_NAMES = ["hold", "ride", "sit at", "carry", "look at"]
GRAD_KEYS = [
    "transformer.level_embed", "bias_obj_a", "bias_pred_a", "input_proj.0.0.bias", "input_proj.3.1.weight",
    "transformer.encoder.layers.0.self_attn.sampling_offsets.bias",
    "transformer.encoder.layers.5.self_attn.attention_weights.bias",
    "transformer.encoder.layers.2.linear1.bias", "transformer.encoder.VLFuse_layers.0.b_attn.gamma_v",
    "transformer.encoder.VLFuse_layers.2.b_attn.gamma_l", "transformer.encoder.VLFuse_layers.1.b_attn.attn.out_l_proj.bias",
    "transformer.encoder.roberta_layers.0.attention.self.query.bias", "transformer.encoder.roberta_layers.2.output.LayerNorm.weight",
    "transformer.resizer.fc.bias", "transformer.ho_decoder.layers.0.cross_attn.sampling_offsets.bias",
    "transformer.ho_decoder.ref_point_head.layers.1.bias", "transformer.verb_decoder.query_scale.layers.0.bias",
    "transformer.verb_tgt_generator.fc_3.5.bias", "transformer.verb_decoder.layers.2.norm3.weight",
    "sub_bbox_embed.0.layers.2.bias", "obj_bbox_embed.2.layers.2.bias", "projection_text.bias",
    "tgt_embed.weight", "verb_tgt_embed.weight", "refpoint_embed.weight",
    "transformer.text_encoder.pooler.dense.bias", "backbone.0.body.layer4.2.conv3.weight",
]


def make_step_inputs():
    g = torch.Generator().manual_seed(7)
    imgs = [torch.randn(3, 96, 128, generator=g), torch.randn(3, 80, 112, generator=g)]
    targets = []
    for k in (3, 2):
        c = torch.rand(k, 2, generator=g) * 0.4 + 0.3
        wh = torch.rand(k, 2, generator=g) * 0.2 + 0.1
        c2 = torch.rand(k, 2, generator=g) * 0.4 + 0.3
        wh2 = torch.rand(k, 2, generator=g) * 0.2 + 0.1
        verbs = torch.zeros(k, len(VERB_NAMES))
        verbs[torch.arange(k), torch.randint(0, len(VERB_NAMES), (k,), generator=g)] = 1
        targets.append({
            "obj_labels": torch.randint(0, len(OBJ_NAMES) - 1, (k,), generator=g),
            "sub_labels": torch.zeros(k, dtype=torch.long),
            "verb_labels": verbs,
            "sub_boxes": torch.cat([c, wh], 1), "obj_boxes": torch.cat([c2, wh2], 1),
        })
    text = [(OBJ_NAMES, VERB_NAMES)]
    return imgs, targets, text


def gen_alif_and_roberta():
    ref_import.install()
    import models.fuse_helper as fh
    import models.modeling_roberta as mr
    from rlipv2_b200.text_encoder import roberta_base_config
    args = ref_import.parse_args(ref_import.PARSEDA_FLAGS + ["--num_queries", "16"])
    g = torch.Generator().manual_seed(11)
    bs, Tv, Tl = 2, 37, 19
    v = torch.randn(bs, Tv, 256, generator=g)
    pos = torch.randn(bs, Tv, 256, generator=g)
    l = torch.randn(bs, Tl, 768, generator=g)
    mask_v = torch.ones(bs, Tv, dtype=torch.bool)
    mask_v[1, 30:] = False
    mask_l = torch.ones(bs, Tl, dtype=torch.bool)
    mask_l[:, 15:] = False
    fuse = det_fill_(fh.RLIPv2_VLFuse(args), seed=1).eval()
    with torch.no_grad():
        out = fuse({"visual": {"src": v.clone(), "padding_mask": mask_v, "pos": pos},
                    "lang": {"hidden": l.clone(), "masks": mask_l}})
    np.savez_compressed(os.path.join(OUT, "parseda_alif.npz"), v=v.numpy(), pos=pos.numpy(), l=l.numpy(),
                        mask_v=mask_v.numpy(), mask_l=mask_l.numpy(),
                        out_v=out["visual"]["src"].numpy(), out_l=out["lang"]["hidden"].numpy())
    layer = det_fill_(mr.RobertaLayer(roberta_base_config()), seed=2).eval()
    with torch.no_grad():
        y = layer(hidden_states=l, attention_mask=mask_l)
    np.savez_compressed(os.path.join(OUT, "parseda_roberta.npz"), x=l.numpy(), mask=mask_l.numpy(), y=y.numpy())
    print("alif/roberta fixtures written")


def gen_step():
    model, criterion, _, args = ref_import.build_reference_model(num_queries=16)
    json.dump({k: list(v.shape) for k, v in model.state_dict().items()},
              open(os.path.join(OUT, "parseda_state_dict_keys.json"), "w"), indent=0)
    det_fill_(model, seed=3)
    model.eval()
    criterion.eval()
    imgs, targets, text = make_step_inputs()
    from util.misc import nested_tensor_from_tensor_list
    samples = nested_tensor_from_tensor_list(imgs)
    for p in model.parameters():
        p.requires_grad_(True)
    memory_cache = model(samples, encode_and_save=True, text=text, targets=targets)
    outputs = model(samples, encode_and_save=False, memory_cache=memory_cache, text=text, targets=targets)
    # margin check of SURVEY quirk 4 (label treated as padding iff pooled vector sums <= 0)
    sums = memory_cache["text_memory_bf_resize"].sum(-1)
    print("pooled text sums (min |.|):", float(sums.abs().min()))
    assert float(sums.abs().min()) > 1e-2, "pick another seed: a text-mask decision is too close to 0"
    loss_dict = criterion(outputs, targets)
    wd = criterion.weight_dict
    total = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
    total.backward()
    indices = criterion.matcher({k: v for k, v in outputs.items() if k != "aux_outputs"}, targets)
    aux_indices = [criterion.matcher(a, targets) for a in outputs["aux_outputs"]]
    params = model.state_dict(keep_vars=True)      # includes the aliased bbox-head names
    save = {
        "img0": imgs[0].numpy(), "img1": imgs[1].numpy(),
        "img_memory_slice": memory_cache["img_memory"][:, ::3, ::8].detach().numpy(),
        "text_memory_resized": memory_cache["text_memory_resized"].detach().numpy(),
        "text_attention_mask": memory_cache["text_attention_mask"].numpy(),
        "valid_ratios": memory_cache["valid_ratios"].numpy(),
        "total_loss": total.detach().numpy(),
    }
    for k in ("pred_sub_logits", "pred_obj_logits", "pred_verb_logits", "pred_sub_boxes", "pred_obj_boxes"):
        save["out_" + k] = outputs[k].detach().numpy()
        for i, a in enumerate(outputs["aux_outputs"]):
            save[f"aux{i}_" + k] = a[k].detach().numpy()
    for k, v in loss_dict.items():
        save["loss_" + k] = np.asarray(float(v))
    for li, ind in enumerate([indices] + aux_indices):
        for b, (i, j) in enumerate(ind):
            save[f"match{li}_{b}_i"] = i.numpy()
            save[f"match{li}_{b}_j"] = j.numpy()
    for t_i, t in enumerate(targets):
        for k, v in t.items():
            save[f"tgt{t_i}_{k}"] = v.numpy()
    for k in GRAD_KEYS:
        gk = params[k].grad
        assert gk is not None, k
        flat = gk.detach().reshape(-1)
        save["grad_" + k] = flat[:: max(1, flat.numel() // 512)].numpy()
        save["gradnorm_" + k] = np.asarray(float(gk.norm()))
    np.savez_compressed(os.path.join(OUT, "parseda_step.npz"), **save)
    print("step fixture written; total loss", float(total), {k: round(float(v), 5) for k, v in loss_dict.items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_alif_and_roberta()
    gen_step()
