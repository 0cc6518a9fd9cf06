"""TEST INFRASTRUCTURE - generator of the Swin fixtures (BASELINE config 4: `--backbone swin_large`).

Runs HERE (build container): the reference's own Swin backbone (/root/reference/models/swin/backbone.py:194-205 ->
swin_transformer.py:585-763, timm stubbed by oracle/ref_import.py) and the reference's own RLIP_ParSeDA built on it
(models/detr.py:326-327), name-keyed weights from oracle/detfill.py, eval mode, fp32, CPU.

  swin_backbone.npz            swin_tiny Joiner on two images of different, window-unfriendly sizes (patch padding, window
                               padding, odd patch-merging sizes, padding masks, the Swin copy of the sine embedding):
                               every level's features (strided sample + full-tensor moments), masks, position embedding
                               samples, gradient norms of a fixed linear functional of the features
  swin_large_keys.json         key -> shape of the reference swin_large backbone state_dict + requires_grad of parameters
  parseda_swin_large_step.npz  RLIP_ParSeDA + swin_large (config 4's model flags, drop_path_rate 0.5, 16 queries):
                               two-phase forward + SetCriterionHOI + matcher on the inputs of parseda_step.npz

    python oracle/gen_golden_swin.py
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
GRAD_KEYS = ["0.body.patch_embed.proj.weight", "0.body.layers.0.blocks.1.attn.qkv.weight",
             "0.body.layers.1.blocks.0.mlp.fc2.bias", "0.body.layers.2.blocks.5.attn.proj.weight",
             "0.body.layers.2.downsample.reduction.weight", "0.body.layers.3.blocks.1.mlp.fc1.weight"]


def swin_inputs():
    g = torch.Generator().manual_seed(17)
    return [torch.randn(3, 150, 203, generator=g), torch.randn(3, 131, 180, generator=g)]


def functional_weights(shape, level):
    g = torch.Generator().manual_seed(100 + level)
    return torch.randn(shape, generator=g)


def gen_backbone():
    ref_import.install()
    args = ref_import.parse_args(ref_import.PARSEDA_FLAGS + ["--num_queries", "16", "--backbone", "swin_tiny",
                                                             "--drop_path_rate", "0.2"])
    with ref_import.chdir(ref_import.REF):
        from models.swin.backbone import build_backbone
        from util.misc import nested_tensor_from_tensor_list
        joiner = build_backbone(args)
    det_fill_(joiner, seed=5)
    joiner.eval()
    samples = nested_tensor_from_tensor_list(swin_inputs())
    feats, pos = joiner(samples)
    save = {"strides": np.asarray(joiner.strides), "num_channels": np.asarray(joiner.num_channels)}
    total = 0
    for l, (f, p) in enumerate(zip(feats, pos)):
        x = f.tensors
        save[f"shape_{l}"] = np.asarray(x.shape)
        save[f"feat_{l}"] = x[:, ::7].detach().numpy()
        save[f"moments_{l}"] = np.asarray([float(x.mean()), float(x.abs().mean()), float(x.pow(2).mean())])
        save[f"mask_{l}"] = f.mask.numpy()
        save[f"pos_{l}"] = p[:, ::16].detach().numpy()
        total = total + (x * functional_weights(x.shape, l)).sum()
    total.backward()
    params = dict(joiner.named_parameters())
    for k in GRAD_KEYS:
        save["gradnorm_" + k] = np.asarray(float(params[k].grad.norm()))
    save["frozen"] = np.asarray(sorted(k for k, p in params.items() if not p.requires_grad))
    np.savez_compressed(os.path.join(OUT, "swin_backbone.npz"), **save)
    print("swin_tiny backbone fixture:", [tuple(f.tensors.shape) for f in feats], "functional", float(total))


def gen_large():
    ref_import.install()
    flags = ref_import.PARSEDA_FLAGS + ["--num_queries", "16", "--backbone", "swin_large", "--drop_path_rate", "0.5"]
    args = ref_import.parse_args(flags)
    with ref_import.chdir(ref_import.REF):
        from models import build_model
        from util.misc import nested_tensor_from_tensor_list
        model, criterion, _ = build_model(args)
    body = model.backbone
    keys = {k: list(v.shape) for k, v in body.state_dict().items()}
    grads = {k: bool(p.requires_grad) for k, p in body.named_parameters()}
    json.dump({"shapes": keys, "requires_grad": grads}, open(os.path.join(OUT, "swin_large_keys.json"), "w"), indent=0)
    det_fill_(model, seed=3)
    model.eval()
    criterion.eval()
    imgs, targets, text = make_step_inputs()
    samples = nested_tensor_from_tensor_list(imgs)
    with torch.no_grad():
        cache = model(samples, encode_and_save=True, text=text, targets=targets)
        out = model(samples, encode_and_save=False, memory_cache=cache, text=text, targets=targets)
        loss_dict = criterion(out, targets)
    wd = criterion.weight_dict
    total = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
    indices = criterion.matcher({k: v for k, v in out.items() if k != "aux_outputs"}, targets)
    save = {"total_loss": total.numpy(), "n_params": np.asarray(sum(p.numel() for p in model.parameters()))}
    for k in ("pred_sub_logits", "pred_obj_logits", "pred_verb_logits", "pred_sub_boxes", "pred_obj_boxes"):
        save["out_" + k] = out[k].numpy()
    for k, v in loss_dict.items():
        save["loss_" + k] = np.asarray(float(v))
    for b, (i, j) in enumerate(indices):
        save[f"match_{b}_i"], save[f"match_{b}_j"] = i.numpy(), j.numpy()
    for k in ("input_proj.0.0.weight", "input_proj.2.0.weight", "input_proj.3.0.weight"):
        save["shape_" + k] = np.asarray(model.state_dict()[k].shape)
    np.savez_compressed(os.path.join(OUT, "parseda_swin_large_step.npz"), **save)
    print("swin_large step fixture: total", float(total), "params", int(save["n_params"]), "backbone keys", len(keys))


if __name__ == "__main__":
    gen_backbone()
    gen_large()
