"""Host logic of the text-tower plumbing (rlipv2_b200/text_encoder.py) on CPU: the synthetic RoBERTa-base stand-in,
the graph-safe additive mask path of `pooled_text`, and the attention-interface hook - which must leave CPU results
exactly HF's (the short-sequence kernel only takes CUDA fp32 inputs of <= 8 tokens)."""
import torch

from rlipv2_b200.text_encoder import (_ATTN_KEY, build_text_encoder, hash_tokenize, pooled_text, route_through_dense_seam,
                                       use_short_attention)


def _tiny_encoder():
    from transformers import RobertaConfig, RobertaModel
    cfg = RobertaConfig(vocab_size=50265, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                        intermediate_size=256, max_position_embeddings=64, type_vocab_size=1, pad_token_id=1,
                        hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    return RobertaModel(cfg).eval()


def test_hash_tokenizer_is_deterministic_and_padded():
    ids, mask = hash_tokenize(["object kind 3", "no objects", "x"])
    ids2, _ = hash_tokenize(["object kind 3", "no objects", "x"])
    assert torch.equal(ids, ids2) and ids.shape == mask.shape == (3, 5)
    assert ids[:, 0].eq(0).all() and mask.sum(1).tolist() == [5, 4, 3]
    assert ids[2, 3:].eq(1).all()                                   # pad id


def test_pooled_text_fast_path_equals_hf_forward():
    enc = _tiny_encoder()
    ids, mask = hash_tokenize(["a b c d", "hello world", "x"])
    ref = enc(input_ids=ids, attention_mask=mask).pooler_output
    torch.testing.assert_close(pooled_text(enc, ids, mask), ref, rtol=1e-5, atol=1e-6)


def test_attention_hook_and_dense_seam_leave_cpu_results_unchanged():
    enc = _tiny_encoder()
    ids, mask = hash_tokenize(["a b c d", "hello world", "x"])
    ref = pooled_text(enc, ids, mask)
    route_through_dense_seam(enc)
    use_short_attention(enc)
    assert enc.config._attn_implementation == _ATTN_KEY
    salts = [m._rlipv2_salt for m in enc.modules() if hasattr(m, "_rlipv2_salt")]
    assert salts == list(range(len(salts))) and len(salts) == 2
    out = pooled_text(enc, ids, mask)
    assert torch.equal(out, ref)                                    # CPU tensors fall through to HF's eager attention
    # a caller that invokes the tower directly (engine.py:377, the evaluation loop) goes through HF's own mask
    # construction under the registered attention key: padded tokens must stay masked there as well
    direct = enc(input_ids=ids, attention_mask=mask).pooler_output
    torch.testing.assert_close(direct, ref, rtol=1e-5, atol=1e-6)
    (out.sum()).backward()                                          # and autograd still reaches the parameters
    assert enc.encoder.layer[0].attention.self.query.weight.grad is not None


def test_synthetic_encoder_has_roberta_base_shape():
    tok, enc = build_text_encoder(synthetic=True)
    assert enc.config.hidden_size == 768 and enc.config.num_hidden_layers == 12 and enc.config.num_attention_heads == 12
    be = tok.batch_encode_plus(["no objects"], padding="longest", return_tensors="pt")
    assert be["input_ids"].shape == (1, 4)
