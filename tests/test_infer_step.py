"""Inference step (rlipv2_b200/infer_step.py): the per-batch body of /root/reference/engine.py:360-430 as one call.
CPU: equals running the text encoding of engine.py:367-391, the two model phases and PostProcessHOI by hand (the pre-encoded
`text` tuple with an all-False label mask - NOT the training-style label strings, whose mask follows SURVEY quirk 4).  The
reference's own loop is checked against the same model in tests/test_reference_engine_dropin.py; the CUDA-graph variant in
tests/test_zz3_infer_gpu.py."""
import torch

OBJ = ["person", "cup", "bench", "dining table"]
VERB = ["hold", "sit on", "look at"]


def _build(device):
    from oracle.detfill import det_fill_
    from rlipv2_b200 import dense, models
    dense.set_matmul_precision("fp32")
    model, _, post = models.build_model(models.default_args(device=device, num_queries=16, synthetic_text_encoder=True))
    det_fill_(model, seed=3)
    return model.to(device).eval(), post["hoi"]


def test_inference_step_equals_manual_phases(msda_cpu_stub):
    from rlipv2_b200.infer_step import ParSeDAInference
    model, post = _build("cpu")
    g = torch.Generator().manual_seed(9)
    imgs = [torch.randn(3, 64, 96, generator=g), torch.randn(3, 56, 80, generator=g)]
    sizes = torch.tensor([[256, 384], [224, 320]])
    infer = ParSeDAInference(model, post, OBJ, VERB, batch_size=2)
    got = infer(imgs, sizes)
    def by_hand(images, sz):
        with torch.no_grad():
            bs = len(images)
            tr = model.transformer
            tok = tr.tokenizer.batch_encode_plus(OBJ + ["no objects"] + VERB, padding="longest", return_tensors="pt")
            mem = tr.text_encoder(**tok).pooler_output.unsqueeze(1).repeat(1, bs, 1)
            text = (torch.zeros(mem.shape[:2], dtype=torch.bool), mem, torch.tensor([[len(OBJ) + 1, len(VERB)]]))
            cache = model(images, encode_and_save=True, text=text)
            return post(model(images, encode_and_save=False, memory_cache=cache, text=text), sz)

    want = by_hand(imgs, sizes)
    assert len(got) == len(want) == 2
    for a, b in zip(got, want):
        assert torch.equal(a["labels"], b["labels"]) and torch.equal(a["sub_ids"], b["sub_ids"])
        assert torch.equal(a["boxes"], b["boxes"]) and torch.equal(a["verb_scores"], b["verb_scores"])
    # short last batch (engine.py:415-419): one image through a batch-2 label set
    one = infer(imgs[:1], sizes[:1])
    assert len(one) == 1 and one[0]["boxes"].shape == (16, 4)
    want1 = by_hand(imgs[:1], sizes[:1])
    assert torch.equal(one[0]["verb_scores"], want1[0]["verb_scores"]) and torch.equal(one[0]["labels"], want1[0]["labels"])
