"""CPU: the gradient-readiness markers behind the overlapped all-reduce of the graphed step (rlipv2_b200/grad_ready.py,
flat_dp.EarlyReducer, train_step.early_reduce_entries; the flat-buffer counterpart of DDP's bucket hooks,
/root/reference/main.py:515-517).  What must hold for an early all-reduce to be correct: at the moment a range is
launched, every gradient in it is final.  Checked on the real model's autograd graph (CPU, CUDA op stubbed)."""
import pytest
import torch


@pytest.mark.parametrize("text_first", [False, True])
def test_gradients_are_final_when_their_markers_fire(msda_cpu_stub, monkeypatch, text_first):
    """text_first=True reproduces the autograd-node creation order of the GPU path on CPU: there the text tower is started
    on a side stream BEFORE the backbone (parseda.py `encode_text_async`), so its nodes are older than the backbone's and the
    'text' marker fires after the backbone's backward has been queued."""
    from oracle.gen_golden_model import OBJ_NAMES, VERB_NAMES, make_step_inputs
    from rlipv2_b200 import dense, grad_ready, models
    from rlipv2_b200.train_step import TEXT_TOWER_SPLIT
    dense.set_matmul_precision("fp32")
    torch.manual_seed(0)
    model, criterion, _ = models.build_model(models.default_args(device="cpu", num_queries=16, synthetic_text_encoder=True))
    model.train()
    criterion.train()
    named = dict(model.named_parameters())

    def layer_of(n):
        return int(n.split("encoder.layer.")[1].split(".")[0]) if "encoder.layer." in n else None

    ranges = {
        "rest": [n for n in named if "backbone" not in n and "text_encoder" not in n],
        "text_mid": [n for n in named if "text_encoder" in n and ((layer_of(n) is not None and layer_of(n) >= TEXT_TOWER_SPLIT)
                                                                  or "pooler" in n)],
        "text_emb": [n for n in named if "text_encoder" in n and layer_of(n) is not None and layer_of(n) < TEXT_TOWER_SPLIT],
    }
    needs = {"rest": {"image0", "image1", "image2", "text"}, "text_mid": {"text_mid"}, "text_emb": {"text_emb"}}
    fired, order, snaps = set(), [], {}

    def cb(tag):
        fired.add(tag)
        order.append(tag)
        for key, tags in needs.items():
            if key not in snaps and tags <= fired:
                snaps[key] = {n: None if named[n].grad is None else named[n].grad.clone() for n in ranges[key]}

    grad_ready.set_callback(cb)
    hooks = grad_ready.install_text_tower_markers(model.transformer.text_encoder, TEXT_TOWER_SPLIT)
    try:
        imgs, targets, text = make_step_inputs()
        if text_first:
            tr = type(model.transformer)
            monkeypatch.setattr(tr, "_join_text", staticmethod(lambda handle, device: handle["_async_text"]))
            text = {"_async_text": model.transformer.encode_text(text, torch.device("cpu")), "_stream": None}
        cache = model(imgs, encode_and_save=True, text=text, targets=targets)
        out = model(imgs, encode_and_save=False, memory_cache=cache, text=text, targets=targets)
        assert grad_ready.applied_tags() == {"image0", "image1", "image2", "text", "text_mid", "text_emb"}
        loss_dict = criterion(out, targets)
        sum(loss_dict[k] * criterion.weight_dict[k] for k in loss_dict if k in criterion.weight_dict).backward()
    finally:
        grad_ready.set_callback(None)
        for h in hooks:
            h.remove()
    assert set(order) == {"image0", "image1", "image2", "text", "text_mid", "text_emb"} and len(order) == 6
    assert order.index("text_mid") < order.index("text_emb")
    if text_first:
        assert order.index("text") > max(order.index(f"image{i}") for i in range(3))     # the GPU path's order
    assert set(snaps) == set(needs)
    n_checked = 0
    for key, snap in snaps.items():
        for n, g in snap.items():
            final = named[n].grad
            assert (g is None) == (final is None), (key, n)
            if g is not None:
                assert torch.equal(g, final), (key, n)            # nothing was added after the marker fired
                n_checked += 1
    assert n_checked > 400
    # and the marker is inert when no callback is installed
    x = torch.ones(3, requires_grad=True)
    assert grad_ready.mark(x, "t") is x


def test_early_reduce_entries_follow_the_flat_layout():
    from rlipv2_b200.train_step import early_reduce_entries
    names = ["input_proj.0.0.weight", "transformer.level_embed", "backbone.0.body.layer2.0.conv1.weight",
             "transformer.text_encoder.embeddings.word_embeddings.weight",
             "transformer.text_encoder.encoder.layer.0.attention.self.query.weight",
             "transformer.text_encoder.encoder.layer.5.output.dense.bias",
             "transformer.text_encoder.encoder.layer.6.attention.self.query.weight",
             "transformer.text_encoder.encoder.layer.11.output.LayerNorm.bias",
             "transformer.text_encoder.pooler.dense.weight"]
    offsets = [0, 10, 16, 28, 128, 138, 148, 158, 160]
    groups = [(0, 14, 1e-4), (16, 26, 1e-5), (28, 170, 1e-5)]
    e = early_reduce_entries(names, offsets, groups)
    assert e == [({"image0", "image1", "image2", "text"}, 0, 14), ({"text_mid"}, 148, 170), ({"text_emb"}, 128, 148)]
    # a tower without per-layer names (e.g. frozen): only the first entry
    assert len(early_reduce_entries(names[:3], offsets[:3], groups)) == 1


def test_early_reducer_remaining_ranges_and_inactive_single_process():
    from rlipv2_b200.flat_dp import EarlyReducer, FlatParams
    ps = [torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(9))]
    flat = FlatParams([(ps[:1], 1e-3), (ps[1:2], 1e-3), (ps[2:], 1e-3)], "cpu")
    n = flat.flat_grad.numel()
    r = EarlyReducer(flat, [({"a", "b"}, 0, 7), ({"c"}, 16, 25)], force=True)
    r.begin()
    r.on_tag("a")
    assert r.launched == [] and r.remaining() == [(0, n)]
    r.on_tag("c")
    assert r.launched == [(16, 25)] and r.remaining() == [(0, 16)] + ([(25, n)] if n > 25 else [])
    r.on_tag("b")
    r.on_tag("b")                                       # firing twice launches once
    assert sorted(r.launched) == [(0, 7), (16, 25)]
    r.finish()
    r.begin()
    assert r.launched == []
    with pytest.raises(AssertionError, match="overlap"):
        EarlyReducer(flat, [({"a"}, 0, 8), ({"b"}, 7, 12)])
    assert not EarlyReducer(flat, [({"a"}, 0, 7)]).active     # world size 1: nothing to do
