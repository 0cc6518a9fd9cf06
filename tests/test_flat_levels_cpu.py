"""Plumbing of the token-major input-projection path (parseda.py FlatLevels + parseda_transformer._encode) on CPU.

The fused op itself (dense.group_norm_tokens -> csrc/fused_ops.cu gn_tok_*) is CUDA only and parity-checked in
tests/test_fused_gpu.py.  Here it is replaced by a torch function of the same contract, so that the model-side wiring -
level order, the extra stride-2 level, masks / position embeddings, the views handed to the transformer and the shortcut
around flatten + cat - can be compared with the reference-shaped path on the same weights."""
import torch
import torch.nn.functional as F

from rlipv2_b200 import dense, models, train_step
from rlipv2_b200.nested import NestedTensor
from tests.conftest import msda_cpu_stub  # noqa: F401  (fixture)


def _torch_group_norm_tokens(xs, norms):
    return torch.cat([F.group_norm(x, n.num_groups, n.weight, n.bias, n.eps).flatten(2).transpose(1, 2)
                      for x, n in zip(xs, norms)], 1)


def test_flat_levels_path_equals_per_level_path(monkeypatch, msda_cpu_stub):
    args = models.default_args(device="cpu", num_queries=16, synthetic_text_encoder=True)
    torch.manual_seed(0)
    model, _, _ = models.build_model(args)
    model.eval()
    text = train_step.synthetic_text(5, 3)
    images, targets = train_step.synthetic_batch(2, 96, 128, n_obj=5, n_verb=3, triplets=2, seed=0, pin=False)
    mask = torch.zeros(2, 96, 128, dtype=torch.bool)
    mask[1, :, 100:] = True                                     # a padded image: masks / valid ratios matter
    samples = NestedTensor(images, mask)
    with torch.no_grad():
        ref = model(samples, encode_and_save=True, text=text, targets=targets)
        monkeypatch.setattr(dense, "group_norm_tokens_supported", lambda *a: True)
        monkeypatch.setattr(dense, "group_norm_tokens", _torch_group_norm_tokens)
        seen = {}
        orig = model.transformer._encode

        def spy(srcs, *a, **k):
            seen["flat"] = getattr(srcs, "flat", None)
            seen["srcs"] = list(srcs)
            return orig(srcs, *a, **k)
        monkeypatch.setattr(model.transformer, "_encode", spy)
        new = model(samples, encode_and_save=True, text=text, targets=targets)
    assert seen["flat"] is not None and len(seen["srcs"]) == 4              # the shortcut was taken, 4 feature levels
    S = sum(s.shape[2] * s.shape[3] for s in seen["srcs"])
    assert seen["flat"].shape == (2, S, 256)
    for k in ("img_memory", "text_memory_resized", "masks", "pos_embed", "valid_ratios", "spatial_shapes",
              "level_start_index"):
        a, b = new[k], ref[k]
        if a.dtype.is_floating_point:
            torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5, msg=lambda m: f"{k}: {m}")
        else:
            assert torch.equal(a, b), k
    # the per-level views are the rows of the token buffer
    start = 0
    for s in seen["srcs"]:
        n, c, h, w = s.shape
        torch.testing.assert_close(s.flatten(2).transpose(1, 2), seen["flat"][:, start:start + h * w])
        start += h * w
