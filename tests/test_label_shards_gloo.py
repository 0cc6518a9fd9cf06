"""SURVEY 8f rank 3 - the label set de-duplicated across ranks (rlipv2_b200/text_encoder.py::pooled_text_sharded): two gloo
ranks that hold the same label strings each encode half of them; the gathered embeddings equal the full encode on every
rank, and after the step's gradient averaging the text tower's parameter gradients equal those of plain data parallelism,
where every rank encodes everything (/root/reference/models/dab_deformable/deformable_transformer.py:489-502 under
main.py:514-519's DDP)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _tower():
    from transformers import RobertaConfig, RobertaModel
    torch.manual_seed(0)
    cfg = RobertaConfig(vocab_size=200, hidden_size=32, num_hidden_layers=2, num_attention_heads=4, intermediate_size=64,
                        max_position_embeddings=40, type_vocab_size=1, pad_token_id=1, bos_token_id=0, eos_token_id=2,
                        hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    cfg._attn_implementation = "eager"
    return RobertaModel(cfg).eval()


def _tokens(n):
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(3, 200, (n, 6), generator=g)
    ids[:, 0] = 0
    lens = torch.randint(3, 7, (n,), generator=g)
    mask = (torch.arange(6)[None] < lens[:, None]).long()
    ids = torch.where(mask.bool(), ids, torch.ones_like(ids))
    return ids, mask


def _worker(rank, port, n_labels, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from rlipv2_b200.text_encoder import pooled_text, pooled_text_sharded
        ids, mask = _tokens(n_labels)
        g = torch.Generator().manual_seed(10 + rank)
        w = torch.randn(n_labels, 32, generator=g)                  # this rank's loss: its own images see every label

        def grads(fn):
            tower = _tower()
            pooled = fn(tower, ids, mask)
            (pooled * w).sum().backward()
            out = {}
            for name, p in tower.named_parameters():
                gr = p.grad if p.grad is not None else torch.zeros_like(p)
                gr = gr.clone()
                dist.all_reduce(gr)                                  # the step's gradient averaging
                out[name] = gr / WORLD
            return pooled.detach(), out
        full, g_full = grads(pooled_text)
        shard, g_shard = grads(pooled_text_sharded)
        # numpy: pickled by value (torch tensors travel as shared-memory handles that die with the worker)
        npd = lambda d: {k: v.numpy() for k, v in d.items()}
        q.put((rank, full.numpy(), shard.numpy(), npd(g_full), npd(g_shard)))
    finally:
        dist.destroy_process_group()


def _run(n_labels):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, n_labels, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(WORLD)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_sharded_label_encoding_equals_full_encoding_on_two_ranks():
    for n_labels in (12, 7):                                         # even split and a ragged tail
        for rank, full, shard, g_full, g_shard in _run(n_labels):
            assert shard.shape == full.shape == (n_labels, 32)
            np.testing.assert_allclose(shard, full, rtol=1e-5, atol=1e-6)
            assert g_full.keys() == g_shard.keys()
            for name in g_full:
                np.testing.assert_allclose(g_shard[name], g_full[name], rtol=1e-4, atol=1e-6, err_msg=f"{name} rank {rank}")


def test_without_a_process_group_the_sharded_call_is_the_plain_one():
    from rlipv2_b200.text_encoder import pooled_text, pooled_text_sharded
    tower = _tower()
    ids, mask = _tokens(5)
    with torch.no_grad():
        torch.testing.assert_close(pooled_text_sharded(tower, ids, mask), pooled_text(tower, ids, mask), rtol=0, atol=0)


def test_transformer_routes_label_text_through_the_sharded_call(monkeypatch):
    """`transformer.shard_label_text()` is what switches `encode_text` over (parseda_transformer.py)"""
    import types
    import rlipv2_b200.parseda_transformer as pt
    calls = []
    real = pt.pooled_text_sharded
    monkeypatch.setattr(pt, "pooled_text_sharded", lambda te, ids, mask, group=None: calls.append(group) or real(te, ids, mask, group))
    ids, mask = _tokens(6)
    tok = {"input_ids": ids, "attention_mask": mask, "sums": [(4, 2)]}
    fake = types.SimpleNamespace(text_encoder=_tower(), shard_labels=False, label_shard_group=None)
    enc = pt.RLIP_ParSeDABDeformableTransformer_v2.encode_text
    with torch.no_grad():
        plain = enc(fake, tok, "cpu")
        assert calls == []
        pt.RLIP_ParSeDABDeformableTransformer_v2.shard_label_text(fake, True, None)
        assert fake.shard_labels is True
        sharded = enc(fake, tok, "cpu")
    assert calls == [None]
    torch.testing.assert_close(sharded[0], plain[0], rtol=0, atol=0)
    assert torch.equal(sharded[1], plain[1])
