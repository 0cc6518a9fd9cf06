"""BASELINE config 4 (`--backbone swin_large`): the timm-free Swin backbone (rlipv2_b200/swin.py) against fixtures produced
by the reference's OWN Swin wrapper and RLIP_ParSeDA build (oracle/gen_golden_swin.py; /root/reference/models/swin/
backbone.py:194-205, swin_transformer.py:585-763), same name-keyed weights (oracle/detfill.py), eval mode, fp32.
Tolerance: north_star's 1e-3 relative; masks, key / shape lists and matcher indices exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.detfill import det_fill_
from tests.golden_util import GOLDEN

DEVICES = ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)]


def _load(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def _args(device, **kw):
    from rlipv2_b200 import models
    return models.default_args(device=device, num_queries=16, synthetic_text_encoder=True, **kw)


def _swin_inputs():
    g = torch.Generator().manual_seed(17)                        # oracle/gen_golden_swin.py::swin_inputs
    return [torch.randn(3, 150, 203, generator=g), torch.randn(3, 131, 180, generator=g)]


@pytest.mark.parametrize("device", DEVICES)
def test_swin_tiny_joiner_matches_reference(device):
    """features of the three output stages on two differently sized images (patch padding, window padding, odd
    patch-merging sizes), padding masks, the Swin wrapper's sine embedding, gradient norms of a linear functional"""
    from rlipv2_b200 import dense
    from rlipv2_b200.backbone import build_backbone
    from rlipv2_b200.nested import nested_tensor_from_tensor_list
    dense.set_matmul_precision("fp32")
    g = _load("swin_backbone.npz")
    joiner = build_backbone(_args(device, backbone="swin_tiny", drop_path_rate=0.2, pretrained_swin="",
                                  use_checkpoint=False))
    det_fill_(joiner, seed=5)
    joiner.to(device).eval()
    assert list(joiner.strides) == list(g["strides"]) and list(joiner.num_channels) == list(g["num_channels"])
    feats, pos = joiner(nested_tensor_from_tensor_list([t.to(device) for t in _swin_inputs()]))
    total = 0
    for l, (f, p) in enumerate(zip(feats, pos)):
        x = f.tensors
        assert list(x.shape) == list(g[f"shape_{l}"])
        scale = float(np.abs(g[f"feat_{l}"]).max())
        np.testing.assert_allclose(x[:, ::7].detach().cpu().numpy(), g[f"feat_{l}"], rtol=1e-3, atol=1e-4 * scale)
        mom = [float(x.detach().mean()), float(x.detach().abs().mean()), float(x.detach().pow(2).mean())]
        np.testing.assert_allclose(mom, g[f"moments_{l}"], rtol=1e-3, atol=1e-5)
        np.testing.assert_array_equal(f.mask.cpu().numpy(), g[f"mask_{l}"])
        np.testing.assert_allclose(p[:, ::16].cpu().numpy(), g[f"pos_{l}"], rtol=1e-4, atol=1e-5)
        w = torch.randn(tuple(x.shape), generator=torch.Generator().manual_seed(100 + l)).to(device)
        total = total + (x * w).sum()
    total.backward()
    params = dict(joiner.named_parameters())
    assert sorted(k for k, p in params.items() if not p.requires_grad) == list(g["frozen"])
    for k in [k[len("gradnorm_"):] for k in g if k.startswith("gradnorm_")]:
        np.testing.assert_allclose(float(params[k].grad.norm()), float(g["gradnorm_" + k]), rtol=2e-3, err_msg=k)


def test_swin_large_keys_and_frozen_set_match_reference():
    from rlipv2_b200.backbone import build_backbone
    ref = json.load(open(os.path.join(GOLDEN, "swin_large_keys.json")))
    with torch.device("meta"):
        joiner = build_backbone(_args("cpu", backbone="swin_large", drop_path_rate=0.5))
    assert {k: list(v.shape) for k, v in joiner.state_dict().items()} == ref["shapes"]
    assert {k: bool(p.requires_grad) for k, p in joiner.named_parameters()} == ref["requires_grad"]
    assert joiner.num_channels == [384, 768, 1536] and joiner.strides == [8, 16, 32]
    # stochastic-depth rates rise linearly to drop_path_rate over the 24 blocks (swin_transformer.py:644)
    blocks = [b for layer in joiner[0].body.layers for b in layer.blocks]
    rates = [getattr(b.drop_path, "drop_prob", 0.0) for b in blocks]
    np.testing.assert_allclose(rates, np.linspace(0, 0.5, 24), atol=1e-6)


@pytest.mark.parametrize("device", DEVICES)
def test_parseda_swin_large_step_matches_reference(device, request):
    """config 4's model: RLIP_ParSeDA on swin_large (input projections 384 / 768 / 1536 -> 256), two-phase forward,
    criterion, matcher - on the inputs of the R50 fixture"""
    if device == "cpu":
        request.getfixturevalue("msda_cpu_stub")
    from oracle.gen_golden_model import OBJ_NAMES, VERB_NAMES
    from rlipv2_b200 import dense, models
    dense.set_matmul_precision("fp32")
    g, gs = _load("parseda_swin_large_step.npz"), _load("parseda_step.npz")
    model, criterion, _ = models.build_model(_args(device, backbone="swin_large", drop_path_rate=0.5))
    assert sum(p.numel() for p in model.parameters()) == int(g["n_params"])
    for k in ("input_proj.0.0.weight", "input_proj.2.0.weight", "input_proj.3.0.weight"):
        assert list(model.state_dict()[k].shape) == list(g["shape_" + k])
    det_fill_(model, seed=3)
    model.to(device).eval()
    criterion.to(device).eval()
    imgs = [torch.from_numpy(gs["img0"]).to(device), torch.from_numpy(gs["img1"]).to(device)]
    targets = [{k: torch.from_numpy(gs[f"tgt{i}_{k}"]).to(device)
                for k in ("obj_labels", "sub_labels", "verb_labels", "sub_boxes", "obj_boxes")} for i in range(2)]
    text = [(OBJ_NAMES, VERB_NAMES)]
    with torch.no_grad():
        cache = model(imgs, encode_and_save=True, text=text, targets=targets)
        out = model(imgs, encode_and_save=False, memory_cache=cache, text=text, targets=targets)
        loss_dict = criterion(out, targets)
    c = lambda t: t.detach().cpu().numpy()
    for k in ("pred_sub_logits", "pred_obj_logits", "pred_verb_logits", "pred_sub_boxes", "pred_obj_boxes"):
        np.testing.assert_allclose(c(out[k]), g["out_" + k], rtol=1e-3, atol=2e-4, err_msg=k)
    for b, (i, j) in enumerate(criterion.matcher({k: v for k, v in out.items() if k != "aux_outputs"}, targets)):
        np.testing.assert_array_equal(i.numpy(), g[f"match_{b}_i"])
        np.testing.assert_array_equal(j.numpy(), g[f"match_{b}_j"])
    assert sorted(loss_dict.keys()) == sorted(k[len("loss_"):] for k in g if k.startswith("loss_"))
    for k, v in loss_dict.items():
        np.testing.assert_allclose(float(v), float(g["loss_" + k]), rtol=1e-3, atol=1e-4, err_msg=k)


def test_drop_path_is_per_sample_and_unbiased():
    from rlipv2_b200.swin import DropPath
    torch.manual_seed(0)
    dp = DropPath(0.25).train()
    x = torch.ones(4096, 3, 5)
    y = dp(x)
    per_sample = y.flatten(1)
    assert bool(((per_sample == 0).all(1) | (per_sample == 1 / 0.75).all(1)).all())          # whole samples dropped / scaled
    assert abs(float(y.mean()) - 1.0) < 0.05
    assert dp.eval()(x) is x


def test_shifted_window_attention_equals_explicit_formulation():
    """WindowAttention's single fused-attention call (position bias + region mask merged) against the textbook sequence:
    scale, scores, + bias, + mask per window, softmax, weighted sum"""
    from rlipv2_b200.swin import WindowAttention, shifted_window_mask, to_windows
    torch.manual_seed(3)
    ws, heads, dim, B, Hp, Wp = 7, 3, 24, 2, 14, 21
    attn = WindowAttention(dim, (ws, ws), heads).eval()
    torch.nn.init.normal_(attn.relative_position_bias_table, std=0.5)
    x = to_windows(torch.randn(B, Hp, Wp, dim), ws)
    mask = shifted_window_mask(Hp, Wp, ws, ws // 2, x.device)
    nW, T = mask.shape[0], ws * ws
    qkv = attn.qkv(x).view(-1, T, 3, heads, dim // heads).permute(2, 0, 3, 1, 4)
    s = (qkv[0] * attn.scale) @ qkv[1].transpose(-2, -1) + attn.position_bias()[None]
    s = (s.view(B, nW, heads, T, T) + mask[None, :, None]).view(-1, heads, T, T).softmax(-1)
    want = attn.proj((s @ qkv[2]).transpose(1, 2).reshape(-1, T, dim))
    torch.testing.assert_close(attn(x, mask), want, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(attn(x), attn(x, torch.zeros_like(mask)), rtol=1e-4, atol=1e-5)
    # the region mask separates exactly the cells that wrapped around: the last window row / column only
    assert float(mask[0].abs().sum()) == 0 and float(mask[-1].abs().sum()) > 0


def test_swin_train_mode_step_has_gradients_where_the_reference_trains(msda_cpu_stub):
    """train() mode (stochastic depth active), forward + criterion + backward through the Swin backbone: finite loss,
    gradients on the trainable backbone parameters, none on the frozen position tables / LayerNorms
    (models/swin/backbone.py:68-70); optimizer groups of main.py:525-537 still select by 'backbone'"""
    from oracle.gen_golden_model import OBJ_NAMES, VERB_NAMES, make_step_inputs
    from rlipv2_b200 import dense, models
    dense.set_matmul_precision("fp32")
    torch.manual_seed(0)
    model, criterion, _ = models.build_model(_args("cpu", backbone="swin_tiny", drop_path_rate=0.2))
    model.train()
    criterion.train()
    imgs, targets, text = make_step_inputs()
    cache = model(imgs, encode_and_save=True, text=text, targets=targets)
    out = model(imgs, encode_and_save=False, memory_cache=cache, text=text, targets=targets)
    loss_dict = criterion(out, targets)
    total = sum(loss_dict[k] * criterion.weight_dict[k] for k in loss_dict if k in criterion.weight_dict)
    assert bool(torch.isfinite(total))
    total.backward()
    body = dict(model.backbone[0].body.named_parameters())
    for k, p in body.items():
        frozen = "relative_position_bias_table" in k or "norm" in k
        assert (p.grad is None) == frozen, k
    assert float(body["layers.3.blocks.1.mlp.fc2.weight"].grad.abs().sum()) > 0
    assert float(body["patch_embed.proj.weight"].grad.abs().sum()) > 0
