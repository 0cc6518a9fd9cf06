"""Drop-in at the module surface: the reference's OWN training loop (`engine.train_one_epoch`,
/root/reference/engine.py:45-201, with its `merge_batch_data` text merging and `utils.reduce_dict`
logging) drives this repo's `build_model` products unchanged - same call protocol
(`model(samples, encode_and_save=True/False, memory_cache=..., text=..., targets=...)`,
`model.module.transformer.text_encoder`, `criterion(outputs, targets)`, `criterion.weight_dict`, three
optimizer groups selected by parameter-name substrings, main.py:523-539).

Only runs where /root/reference exists (the build container); CPU, with the CUDA op replaced by the
golden-pinned torch oracle (msda_cpu_stub)."""
import os

import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")


class _Holder(torch.nn.Module):
    """engine.py:71 dereferences `model.module` (the reference always wraps in DDP)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


def _loader(n_batches=2):
    ref_import.install()
    from util.misc import nested_tensor_from_tensor_list      # the reference's own collate pieces
    objs, verbs = ["person", "cup", "bench"], ["hold", "sit on"]
    g = torch.Generator().manual_seed(0)
    batches = []
    for _ in range(n_batches):
        imgs = [torch.randn(3, 64, 96, generator=g), torch.randn(3, 56, 80, generator=g)]
        targets = []
        for k in (2, 1):
            v = torch.zeros(k, len(verbs))
            v[torch.arange(k), torch.randint(0, len(verbs), (k,), generator=g)] = 1
            box = lambda: torch.cat([torch.rand(k, 2, generator=g) * 0.4 + 0.3, torch.rand(k, 2, generator=g) * 0.2 + 0.1], 1)
            targets.append({"obj_labels": torch.randint(0, len(objs), (k,), generator=g), "sub_labels": torch.zeros(k, dtype=torch.long),
                            "verb_labels": v, "sub_boxes": box(), "obj_boxes": box(), "obj_classes": list(objs),
                            "verb_classes": list(verbs), "filename": "synthetic", "image_id": 0})
        batches.append((nested_tensor_from_tensor_list(imgs), targets))

    class Loader(list):
        dataset = None
    return Loader(batches)


def test_reference_train_one_epoch_drives_this_model(msda_cpu_stub):
    ref_import.install()
    args = ref_import.parse_args(ref_import.PARSEDA_FLAGS + ["--num_queries", "16", "--epochs", "1"])
    args.synthetic_text_encoder = True
    from rlipv2_b200 import models
    model, criterion, _ = models.build_model(args)          # the reference's own argparse namespace
    with ref_import.chdir(ref_import.REF):
        import engine as ref_engine
    holder = _Holder(model)
    named = list(model.named_parameters())
    groups = [                                               # main.py:523-539
        {"params": [p for n, p in named if "backbone" not in n and "text_encoder" not in n and p.requires_grad]},
        {"params": [p for n, p in named if "backbone" in n and p.requires_grad], "lr": args.lr_backbone},
        {"params": [p for n, p in named if "text_encoder" in n and p.requires_grad], "lr": args.text_encoder_lr},
    ]
    optimizer = torch.optim.AdamW(groups, lr=args.lr, weight_decay=args.weight_decay)
    before = model.transformer.level_embed.detach().clone()
    stats = ref_engine.train_one_epoch(holder, criterion, _loader(), optimizer, torch.device("cpu"), 0,
                                       args.clip_max_norm, args=args)
    assert stats["loss"] == stats["loss"] and stats["loss"] > 0
    for k in ("loss_obj_ce", "loss_verb_ce", "loss_sub_bbox", "loss_obj_giou", "obj_class_error", "sub_class_error", "lr_text_encoder"):
        assert k in stats, k
    assert not torch.equal(before, model.transformer.level_embed.detach())      # the optimizer stepped


def _eval_loader():
    ref_import.install()
    from util.misc import nested_tensor_from_tensor_list
    g = torch.Generator().manual_seed(5)
    batches = []
    for b, sizes in enumerate([((64, 96), (56, 80)), ((72, 72),)]):          # the last batch is short (engine.py:415-419)
        imgs = [torch.randn(3, h, w, generator=g) for h, w in sizes]
        targets = [{"orig_size": torch.tensor([h * 4, w * 4]), "size": torch.tensor([h, w]), "id": 10 * b + i,
                    "filename": f"synthetic_{b}_{i}.jpg"} for i, (h, w) in enumerate(sizes)]
        batches.append((nested_tensor_from_tensor_list(imgs), targets))

    class Dataset:
        object_text = ["person", "cup", "bench", "dining table"]
        verb_text = ["hold", "sit on", "look at"]
        rare_triplets, non_rare_triplets, correct_mat = [], [], None

    class Loader(list):
        dataset = Dataset()
    return Loader(batches), Dataset()


def test_reference_evaluate_hoi_with_text_drives_this_model(msda_cpu_stub, monkeypatch):
    """the reference's OWN evaluation loop (engine.py:360-468: label strings encoded once through
    `model.module.transformer.tokenizer / text_encoder`, pre-encoded `text` tuples, short last batch, post-processor,
    all_gather) runs this repo's model + PostProcessHOI, and the per-image predictions it hands to the evaluator equal the
    ones the reference's own model + post-processor produce from the same name-keyed weights."""
    from oracle.detfill import det_fill_
    from rlipv2_b200 import dense, models
    dense.set_matmul_precision("fp32")
    args = ref_import.parse_args(ref_import.PARSEDA_FLAGS + ["--num_queries", "16", "--batch_size", "2", "--eval"])
    args.synthetic_text_encoder = True
    with ref_import.chdir(ref_import.REF):
        import engine as ref_engine
        from models import build_model as ref_build_model
        ref_model, _, ref_post = ref_build_model(args)
    captured = []

    class _Evaluator:                                                    # datasets/hico_eval.py is outside the hot path
        def __init__(self, preds, gts, *a):
            captured.append((preds, gts))

        def evaluate(self):
            return {"mAP": 0.0}

    monkeypatch.setattr(ref_engine, "HICOEvaluator", _Evaluator)
    model, _, post = models.build_model(args)
    loader, dataset_val = _eval_loader()
    for m, p in ((ref_model, ref_post), (model, post)):
        det_fill_(m, seed=3)
        stats = ref_engine.evaluate_hoi_with_text("hico", _Holder(m), p, loader, dataset_val, 0, torch.device("cpu"), args)
        assert stats == {"mAP": 0.0}
    (ref_preds, ref_gts), (preds, gts) = captured
    assert len(preds) == len(ref_preds) == 3 and [t["id"] for t in gts] == [t["id"] for t in ref_gts]
    for mine, ref in zip(preds, ref_preds):
        assert sorted(mine.keys()) == sorted(ref.keys())
        for k in ref:
            a, b = torch.as_tensor(mine[k]), torch.as_tensor(ref[k])
            assert a.shape == b.shape and a.dtype == b.dtype, k
            if a.is_floating_point():
                torch.testing.assert_close(a, b, rtol=1e-3, atol=1e-3 if k == "boxes" else 1e-5, msg=k)
            else:
                assert torch.equal(a, b), k
