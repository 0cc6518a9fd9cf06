"""Module-level parity of the ParSeDA hot path against golden fixtures produced by the reference's
own modules (oracle/gen_golden_model.py: RLIPv2_VLFuse, RobertaLayer, full RLIP_ParSeDA two-phase
forward + SetCriterionHOI + HungarianMatcherHOI + backward), same name-keyed weights
(oracle/detfill.py), eval() mode, fp32.

Each test runs twice: on CPU (host logic; the CUDA op replaced by the golden-pinned torch oracle via
the `msda_cpu_stub` fixture) and, marked gpu, on cuda:0 through the real sm_100a kernels.
Tolerance: north_star's 1e-3 relative fp32 (matcher indices bit-exact)."""
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


def _fp32():
    from rlipv2_b200 import dense
    dense.set_matmul_precision("fp32")


def test_state_dict_keys_match_reference():
    from rlipv2_b200 import models
    model, _, _ = models.build_model(_args("cpu"))
    ref = json.load(open(os.path.join(GOLDEN, "parseda_state_dict_keys.json")))
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert mine == ref
    # the aliased box heads share storage (hoi.py:1980-1990)
    sd = model.state_dict()
    assert sd["sub_bbox_embed.0.layers.0.weight"].data_ptr() == \
        sd["transformer.ho_decoder.sub_bbox_embed.0.layers.0.weight"].data_ptr()
    assert sd["obj_bbox_embed.3.layers.0.weight"].data_ptr() == \
        sd["transformer.verb_decoder.obj_bbox_embed.0.layers.0.weight"].data_ptr()
    # optimizer groups of main.py:525-537 are selected by these substrings
    names = [n for n, _ in model.named_parameters()]
    assert any("backbone" in n for n in names) and any("text_encoder" in n for n in names)


@pytest.mark.parametrize("device", DEVICES)
def test_alif_block_golden(device):
    _fp32()
    from rlipv2_b200.alif import RLIPv2_VLFuse
    g = _load("parseda_alif.npz")
    fuse = det_fill_(RLIPv2_VLFuse(_args(device)), seed=1).eval().to(device)
    t = lambda k: torch.from_numpy(g[k]).to(device)
    with torch.no_grad():
        out = fuse({"visual": {"src": t("v"), "padding_mask": t("mask_v"), "pos": t("pos")},
                    "lang": {"hidden": t("l"), "masks": t("mask_l")}})
    np.testing.assert_allclose(out["visual"]["src"].cpu().numpy(), g["out_v"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(out["lang"]["hidden"].cpu().numpy(), g["out_l"], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("device", DEVICES)
def test_roberta_layer_golden(device):
    _fp32()
    from rlipv2_b200.roberta_layer import RobertaLayer
    from rlipv2_b200.text_encoder import roberta_base_config
    g = _load("parseda_roberta.npz")
    layer = det_fill_(RobertaLayer(roberta_base_config()), seed=2).eval().to(device)
    with torch.no_grad():
        y = layer(torch.from_numpy(g["x"]).to(device), attention_mask=torch.from_numpy(g["mask"]).to(device))
    np.testing.assert_allclose(y.cpu().numpy(), g["y"], rtol=1e-3, atol=1e-4)


def _run_step(device):
    from oracle.gen_golden_model import GRAD_KEYS, OBJ_NAMES, VERB_NAMES
    from rlipv2_b200 import models
    g = _load("parseda_step.npz")
    model, criterion, _ = models.build_model(_args(device))
    det_fill_(model, seed=3)
    model.to(device).eval()
    criterion.to(device).eval()
    imgs = [torch.from_numpy(g["img0"]).to(device), torch.from_numpy(g["img1"]).to(device)]
    targets = [{k: torch.from_numpy(g[f"tgt{i}_{k}"]).to(device)
                for k in ("obj_labels", "sub_labels", "verb_labels", "sub_boxes", "obj_boxes")} for i in range(2)]
    text = [(OBJ_NAMES, VERB_NAMES)]
    cache = model(imgs, encode_and_save=True, text=text, targets=targets)
    out = model(imgs, encode_and_save=False, memory_cache=cache, text=text, targets=targets)
    loss_dict = criterion(out, targets)
    wd = criterion.weight_dict
    total = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
    total.backward()
    return g, model, criterion, cache, out, loss_dict, total, targets, GRAD_KEYS


def _assert_step(res, tol, loss_rtol, grad_tol, exact_indices=True):
    g, model, criterion, cache, out, loss_dict, total, targets, GRAD_KEYS = res
    c = lambda t: t.detach().cpu().numpy()
    # phase A
    np.testing.assert_array_equal(c(cache["text_attention_mask"]), g["text_attention_mask"])
    np.testing.assert_allclose(c(cache["valid_ratios"]), g["valid_ratios"], rtol=1e-6)
    np.testing.assert_allclose(c(cache["img_memory"][:, ::3, ::8]), g["img_memory_slice"], **tol)
    np.testing.assert_allclose(c(cache["text_memory_resized"]), g["text_memory_resized"], **tol)
    # phase B outputs, all decoder layers
    for k in ("pred_sub_logits", "pred_obj_logits", "pred_verb_logits", "pred_sub_boxes", "pred_obj_boxes"):
        np.testing.assert_allclose(c(out[k]), g["out_" + k], **tol)
        for i, a in enumerate(out["aux_outputs"]):
            np.testing.assert_allclose(c(a[k]), g[f"aux{i}_" + k], **tol)
    # matcher indices: bit-exact (IEEE-class arithmetic), or - under TF32 products - optimal up to the perturbation of the costs
    layers = [{k: v for k, v in out.items() if k != "aux_outputs"}] + list(out["aux_outputs"])
    for li, o in enumerate(layers):
        C, _ = criterion.matcher.compute_costs(o, targets)
        sizes = [len(t["obj_labels"]) for t in targets]
        for b, ((i, j), cb) in enumerate(zip(criterion.matcher(o, targets), C.cpu().split(sizes, -1))):
            assert i.dtype == torch.int64 and j.dtype == torch.int64 and i.device.type == "cpu"
            gi, gj = g[f"match{li}_{b}_i"], g[f"match{li}_{b}_j"]
            if exact_indices:
                np.testing.assert_array_equal(i.numpy(), gi)
                np.testing.assert_array_equal(j.numpy(), gj)
            else:
                # the reference's assignment, priced with THIS run's costs, may beat this run's optimum only by the cost
                # perturbation (2 k delta): indices agree wherever the reference's margin exceeds the TF32 error
                mine, theirs = float(cb[b][i, j].sum()), float(cb[b][torch.from_numpy(gi), torch.from_numpy(gj)].sum())
                assert mine <= theirs + 1e-6
                assert theirs - mine <= 2 * len(gi) * 2e-2, (li, b, mine, theirs)
                assert sorted(j.tolist()) == sorted(gj.tolist())
    # losses: same keys, same values
    gold_keys = sorted(k[len("loss_"):] for k in g if k.startswith("loss_"))
    assert sorted(loss_dict.keys()) == gold_keys
    for k, v in loss_dict.items():
        if not exact_indices and ("class_error" in k or "cardinality_error" in k):
            continue            # logging metrics built on arg-max counts (hoi.py:3723, 3909-3923): steps of 1 / #targets
        np.testing.assert_allclose(float(v), float(g["loss_" + k]), rtol=loss_rtol, atol=loss_rtol / 10, err_msg=k)
    np.testing.assert_allclose(float(total), float(g["total_loss"]), rtol=loss_rtol)
    # backward: gradient norms and strided samples of selected parameters (vector-relative error: the
    # samples contain near-zero entries, where an element-wise relative test is meaningless)
    params = model.state_dict(keep_vars=True)
    worst = {}
    for k in GRAD_KEYS:
        gk = params[k].grad
        assert gk is not None, k
        ref_norm = float(g["gradnorm_" + k])
        flat = gk.detach().reshape(-1)
        sample = c(flat[:: max(1, flat.numel() // 512)]).astype(np.float64)
        ref = g["grad_" + k].astype(np.float64)
        worst[k] = (abs(float(gk.norm()) - ref_norm) / (ref_norm + 1e-12),
                    float(np.linalg.norm(sample - ref) / (np.linalg.norm(ref) + 1e-12)))
    # ALIF's VXAc gate uses gamma[0] only (fuse_helper.py:720-721): that parameter's gradient is ONE sum of ~4e5 signed terms;
    # under TF32 products it moves by up to 10 % from run to run (measured r02d-r02h: 2 % / 5.7 % / 10.3 %), under fp32 / 3xTF32
    # it sits within the common tolerance
    loose = lambda k: 3.0 if (not exact_indices and ".gamma_" in k) else 1.0
    bad = {k: v for k, v in worst.items() if v[0] > grad_tol[0] * loose(k) or v[1] > grad_tol[1] * loose(k)}
    assert not bad, f"gradient mismatch (norm rel err, sample rel err): {bad}"


@pytest.mark.parametrize("device", DEVICES)
def test_full_step_golden(device, request):
    _fp32()
    if device == "cpu":
        request.getfixturevalue("msda_cpu_stub")
    _assert_step(_run_step(device), dict(rtol=1e-3, atol=2e-4), 1e-3, (2e-3, 5e-3))


@pytest.mark.gpu
def test_full_step_golden_3xtf32():
    """The same fixture at the same 1e-3 tolerances with every supported linear / weight gradient / input gradient on the
    tcgen05 kernels, operands split into TF32 hi + lo parts (dense.py '3xtf32'; SURVEY.md section 7's error-compensated
    switch): model-level parity that DOES run through the hand-written contraction kernels."""
    from rlipv2_b200 import dense, dense_abi
    try:
        dense.set_matmul_precision("3xtf32")
        n0 = dense_abi.launch_count()
        res = _run_step("cuda")
        assert dense_abi.launch_count() - n0 > 100          # the step really went through the tcgen05 kernels
        _assert_step(res, dict(rtol=1e-3, atol=2e-4), 1e-3, (2e-3, 5e-3))
    finally:
        _fp32()


@pytest.mark.gpu
def test_full_step_golden_tf32():
    """The benchmark's arithmetic (TF32 tensor-core products: tcgen05 linears, fused attention cores, cuBLAS / cuDNN TF32)
    against the fp32 fixture: every output of every decoder level, every loss, gradient norms and samples, and matcher
    indices wherever the assignment is decided by more than the TF32 perturbation of the costs."""
    from rlipv2_b200 import attn_abi, dense, dense_abi
    try:
        dense.set_matmul_precision("tf32")
        n0, a0 = dense_abi.launch_count(), attn_abi.launch_count()
        res = _run_step("cuda")
        assert dense_abi.launch_count() - n0 > 100 and attn_abi.launch_count() - a0 >= 3 * 3 + 6
        # gradients through ~40 TF32 contractions in series: measured on B200 (r02d, r02e) norm errors <= 2.0e-2 (5.7e-2 for
        # ALIF's scalar gate gamma_l[0], a single sum with cancellation) and sample errors <= 5.5e-2 on the 27 tracked parameters (the reference's own TF32 default sits in the same band); the 3xTF32 run of
        # the same kernels above holds them to 2e-3 / 5e-3
        _assert_step(res, dict(rtol=2e-2, atol=2e-2), 5e-3, (8e-2, 8e-2), exact_indices=False)
    finally:
        _fp32()


@pytest.mark.parametrize("device", DEVICES)
def test_pretrain_step_golden(device, request):
    """BASELINE config 3's model flags (`--cross_modal_pretrain --pseudo_verb` instead of `--hoi`): pseudo relation labels
    (hoi.py:2197-2239) and their use in the verb loss (hoi.py:3925-4028) against the reference's own modules
    (oracle/gen_golden_pretrain.py), same inputs / weights as the fine-tune fixture"""
    _fp32()
    if device == "cpu":
        request.getfixturevalue("msda_cpu_stub")
    from oracle.gen_golden_model import OBJ_NAMES, VERB_NAMES
    from rlipv2_b200 import models
    g, gs = _load("parseda_pretrain_step.npz"), _load("parseda_step.npz")
    model, criterion, _ = models.build_model(_args(device, hoi=False, cross_modal_pretrain=True, pseudo_verb=True))
    det_fill_(model, seed=3)
    model.to(device).eval()
    criterion.to(device).eval()
    imgs = [torch.from_numpy(gs["img0"]).to(device), torch.from_numpy(gs["img1"]).to(device)]
    targets = [{k: torch.from_numpy(gs[f"tgt{i}_{k}"]).to(device)
                for k in ("obj_labels", "sub_labels", "verb_labels", "sub_boxes", "obj_boxes")} for i in range(2)]
    text = [(OBJ_NAMES, VERB_NAMES)]
    cache = model(imgs, encode_and_save=True, text=text, targets=targets)
    out = model(imgs, encode_and_save=False, memory_cache=cache, text=text, targets=targets)
    loss_dict = criterion(out, targets)
    wd = criterion.weight_dict
    total = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
    total.backward()
    c = lambda t: t.detach().cpu().numpy()
    tol = dict(rtol=1e-3, atol=2e-4)
    assert int((g["target_verb_sim"] > 0).sum()) > 0                      # the fixture does carry pseudo labels
    np.testing.assert_allclose(c(out["target_verb_sim"]), g["target_verb_sim"], **tol)
    for aux in out["aux_outputs"]:
        np.testing.assert_allclose(c(aux["target_verb_sim"]), g["target_verb_sim"], **tol)
    for k in ("pred_sub_logits", "pred_obj_logits", "pred_verb_logits", "pred_sub_boxes", "pred_obj_boxes"):
        np.testing.assert_allclose(c(out[k]), g["out_" + k], **tol)
    for b, (i, j) in enumerate(criterion.matcher({k: v for k, v in out.items() if k != "aux_outputs"}, targets)):
        np.testing.assert_array_equal(i.numpy(), g[f"match_{b}_i"])
        np.testing.assert_array_equal(j.numpy(), g[f"match_{b}_j"])
    gold_keys = sorted(k[len("loss_"):] for k in g if k.startswith("loss_"))
    assert sorted(loss_dict.keys()) == gold_keys
    for k, v in loss_dict.items():
        np.testing.assert_allclose(float(v), float(g["loss_" + k]), rtol=1e-3, atol=1e-4, err_msg=k)
    np.testing.assert_allclose(float(total), float(g["total_loss"]), rtol=1e-3)
    params = model.state_dict(keep_vars=True)
    for k in [k[len("gradnorm_"):] for k in g if k.startswith("gradnorm_")]:
        np.testing.assert_allclose(float(params[k].grad.norm()), float(g["gradnorm_" + k]), rtol=3e-3, err_msg=k)
