"""Text-encoder plumbing of the ParSeDA transformer.

The reference encodes every label string with HF `RobertaTokenizerFast` + `RobertaModel`
(roberta-base) each step and keeps `pooler_output`
(/root/reference/models/dab_deformable/deformable_transformer.py:333-335, 497-502).  Those are
third-party components with downloaded weights; they stay third-party here.
Offline (no weights on disk) the synthetic path below is used: a random-init RoBERTa-base of the
same shape and a deterministic hash tokenizer, so that benchmarks and parity fixtures do the same
arithmetic as a real run.
"""
import os
import types
import zlib

import torch
from torch import nn


def roberta_base_config():
    """roberta-base architecture (hidden 768, 12 layers x 12 heads, FFN 3072, LN eps 1e-5)."""
    from transformers import RobertaConfig
    return RobertaConfig(vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                         intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
                         attention_probs_dropout_prob=0.1, max_position_embeddings=514, type_vocab_size=1,
                         layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=0, eos_token_id=2)


def hash_tokenize(texts):
    """Deterministic stand-in for the BPE tokenizer: <s> w1 w2 ... </s>, one id per whitespace word
    (crc32 of the lower-cased word), padded to the longest with pad id 1.
    -> (input_ids [n, T] int64, attention_mask [n, T] int64)"""
    rows = []
    for t in texts:
        words = str(t).lower().split()
        rows.append([0] + [3 + zlib.crc32(w.encode()) % 50000 for w in words] + [2])
    T = max(len(r) for r in rows)
    ids = torch.full((len(rows), T), 1, dtype=torch.long)
    mask = torch.zeros((len(rows), T), dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, :len(r)] = torch.tensor(r)
        mask[i, :len(r)] = 1
    return ids, mask


class HashTokenizer:
    """`batch_encode_plus(texts, padding="longest", return_tensors="pt")` like the HF fast tokenizer."""

    def batch_encode_plus(self, texts, padding="longest", return_tensors="pt"):
        from transformers import BatchEncoding
        ids, mask = hash_tokenize(texts)
        return BatchEncoding({"input_ids": ids, "attention_mask": mask})


def build_text_encoder(text_encoder_type="roberta-base", synthetic=None):
    """-> (tokenizer, text_encoder).  `synthetic=None` tries the local HF cache first and falls back
    to the synthetic pair only when the pretrained files are absent (offline box)."""
    from transformers import RobertaModel, RobertaTokenizerFast
    if not synthetic:
        try:
            tok = RobertaTokenizerFast.from_pretrained(text_encoder_type, local_files_only=True)
            enc = RobertaModel.from_pretrained(text_encoder_type, local_files_only=True, attn_implementation="eager")
            return tok, enc
        except Exception:
            if synthetic is False:
                raise
    cfg = roberta_base_config()
    cfg._attn_implementation = "eager"      # see pooled_text(): bmm+softmax beats flash kernels on 3-8 tokens
    return HashTokenizer(), RobertaModel(cfg)


def pooled_text(text_encoder, input_ids, attention_mask):
    """`text_encoder(input_ids, attention_mask).pooler_output` (what the reference keeps,
    dab_deformable/deformable_transformer.py:502).

    Label strings are 3-8 tokens long; on such sequences HF's plain bmm+softmax ("eager") attention is
    several times faster than the flash / memory-efficient SDPA kernels (3.3 ms -> ~0.5 ms per step), but
    HF's own mask construction for the eager path does `torch.tensor(0.0, device=...)`, a pageable H2D
    copy that a CUDA-graph capture rejects.  So for eager-attention models the additive mask is built
    here (same values: 0 / finfo.min) and the model's own embeddings -> encoder -> pooler are called
    directly; every arithmetic op is still HF's."""
    te = text_encoder
    if (getattr(te.config, "_attn_implementation", None) in ("eager", _ATTN_KEY) and hasattr(te, "embeddings")
            and hasattr(te, "encoder") and getattr(te, "pooler", None) is not None):
        if getattr(te.config, "_attn_implementation", None) == _ATTN_KEY and te.training and input_ids.is_cuda:
            _dropout_seed(input_ids.device).add_(1)                 # fresh dropout masks for this pass
        emb = te.embeddings(input_ids=input_ids)
        ext = (1.0 - attention_mask[:, None, None, :].to(emb.dtype)) * torch.finfo(emb.dtype).min
        return te.pooler(te.encoder(emb, attention_mask=ext).last_hidden_state)
    return te(input_ids=input_ids, attention_mask=attention_mask).pooler_output


# ---- one label set, many ranks: each rank encodes 1 / world of the strings -----------------------------------------
class _GatherRows(torch.autograd.Function):
    """x [rows, C] on every rank -> the ranks' blocks stacked [world * rows, C]; the backward hands every rank the SUM over
    ranks of the gradient of its own block (the label embeddings feed every rank's loss)."""

    @staticmethod
    def forward(ctx, x, group):
        import torch.distributed as dist
        ctx.group = group
        world = dist.get_world_size(group)
        x = x.contiguous()
        out = x.new_empty((world * x.shape[0],) + tuple(x.shape[1:]))
        if dist.get_backend(group) == "gloo":                      # CPU tests: no all_gather_into_tensor / reduce-scatter
            parts = list(out.chunk(world, 0))
            dist.all_gather(parts, x, group=group)
        else:
            dist.all_gather_into_tensor(out, x, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        import torch.distributed as dist
        world, rank = dist.get_world_size(ctx.group), dist.get_rank(ctx.group)
        g = g.contiguous()
        rows = g.shape[0] // world
        if dist.get_backend(ctx.group) == "gloo":
            g = g.clone()
            dist.all_reduce(g, group=ctx.group)
            return g[rank * rows:(rank + 1) * rows].clone(), None
        mine = g.new_empty((rows,) + tuple(g.shape[1:]))
        dist.reduce_scatter_tensor(mine, g, group=ctx.group)
        return mine, None


def pooled_text_sharded(text_encoder, input_ids, attention_mask, group=None):
    """`pooled_text` for a label set that is THE SAME on every rank of `group` (fine-tuning on a fixed vocabulary: every
    rank tokenises the same object + relation names each step, dab_deformable/deformable_transformer.py:489-502, and runs
    the 12-layer tower on all of them).  Rank r encodes rows [r * n, (r + 1) * n) of the (padded) token matrix, the pooled
    vectors are all-gathered; in the backward every rank receives the rank-sum of the gradient of its rows, so that after
    the step's gradient averaging the tower's parameter gradients equal the data-parallel ones (both are
    1 / world * sum over ranks and labels).  Dropout inside the tower is then drawn once per label instead of once per
    label and rank.  Falls back to `pooled_text` without an initialised process group or with one rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return pooled_text(text_encoder, input_ids, attention_mask)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = input_ids.shape[0]
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    ids, mask = input_ids[lo:lo + per], attention_mask[lo:lo + per]
    if ids.shape[0] < per:                                         # ragged tail: pad with copies of the first string
        fill = per - ids.shape[0]
        ids = torch.cat((ids, input_ids[:1].expand(fill, -1)), 0)
        mask = torch.cat((mask, attention_mask[:1].expand(fill, -1)), 0)
    mine = pooled_text(text_encoder, ids, mask)
    return _GatherRows.apply(mine, group)[:n]


# ---- short-sequence attention for the label strings (csrc/fused_ops.cu short_attn_*) --------------------------------
_SHORT_ATTN = os.environ.get("RLIPV2_SHORT_ATTN", "1") != "0"
_ATTN_KEY = "rlipv2_short"
_seeds = {}


def _dropout_seed(device):
    """device-resident int64 counter the dropout masks are hashed from (advanced once per text-tower forward)"""
    s = _seeds.get(device)
    if s is None:
        s = _seeds[device] = torch.zeros(1, dtype=torch.int64, device=device) + (torch.initial_seed() % (2 ** 62))
    return s


class _ShortAttention(torch.autograd.Function):
    """softmax(q k^T * scale + mask) (dropout) v for T <= 8 tokens, head dim 64; one kernel each way.
    q, k, v: [B, H, T, 64] views of [B, T, H * 64] projections (HF layout) -> [B, T, H, 64]"""

    @staticmethod
    def forward(ctx, q, k, v, mask, scale, dropout_p, seed, salt):
        from . import fused_abi
        qt, kt, vt = (t.transpose(1, 2) for t in (q, k, v))            # [B, T, H, 64]: contiguous for HF's views
        qt, kt, vt = (t if t.is_contiguous() else t.contiguous() for t in (qt, kt, vt))
        out, seed_used = fused_abi.short_attention_fwd(qt, kt, vt, mask, scale, dropout_p, seed if dropout_p > 0 else None,
                                                       salt)
        ctx.save_for_backward(qt, kt, vt, mask, seed_used)
        ctx.conf = (scale, dropout_p, salt)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        from . import fused_abi
        qt, kt, vt, mask, seed_used = ctx.saved_tensors
        scale, dropout_p, salt = ctx.conf
        g = grad_out if grad_out.is_contiguous() else grad_out.contiguous()
        dq, dk, dv = fused_abi.short_attention_bwd(qt, kt, vt, mask, g, scale, dropout_p, seed_used, salt)
        return dq.transpose(1, 2), dk.transpose(1, 2), dv.transpose(1, 2), None, None, None, None, None


def _short_attention_forward(module, query, key, value, attention_mask, dropout=0.0, scaling=None, **kwargs):
    """attention interface of HF's RobertaSelfAttention (same contract as modeling_roberta.eager_attention_forward):
    -> (attn_output [B, T, H, D], attn_weights)"""
    from transformers.models.roberta.modeling_roberta import eager_attention_forward
    from . import fused_abi
    B, H, T, D = query.shape
    ok = (query.is_cuda and query.dtype == torch.float32 and D == fused_abi.SHORT_ATTN_D and T <= fused_abi.SHORT_ATTN_MAX_T
          and key.shape == query.shape and value.shape == query.shape
          and (attention_mask is None or (attention_mask.dtype == torch.float32 and attention_mask.numel() == B * T)))
    if not ok:
        return eager_attention_forward(module, query, key, value, attention_mask, dropout=dropout, scaling=scaling, **kwargs)
    if scaling is None:
        scaling = D ** -0.5
    mask = attention_mask.reshape(B, T).contiguous() if attention_mask is not None else None
    p = float(dropout) if module.training else 0.0
    out = _ShortAttention.apply(query, key, value, mask, float(scaling), p, _dropout_seed(query.device),
                                int(getattr(module, "_rlipv2_salt", 0)))
    return out, None


def use_short_attention(text_encoder):
    """route the tower's self-attention through the short-sequence kernel (label strings are 3-8 tokens); longer inputs,
    CPU tensors and other dtypes fall through to HF's eager attention inside the interface function"""
    if not _SHORT_ATTN:
        return text_encoder
    try:
        from transformers import AttentionInterface
        from transformers.masking_utils import AttentionMaskInterface, eager_mask
        AttentionInterface.register(_ATTN_KEY, _short_attention_forward)
        # callers that invoke the tower directly (`model.transformer.text_encoder(**tokens)`, engine.py:377) go through
        # HF's own mask construction, which looks the mask builder up under the attention key: without this entry the
        # raw 0/1 padding mask would reach the attention function unexpanded
        AttentionMaskInterface.register(_ATTN_KEY, eager_mask)
    except Exception:                                   # older transformers: keep HF's eager attention
        return text_encoder
    salt = 0
    for m in text_encoder.modules():
        if m.__class__.__name__ == "RobertaSelfAttention":
            m._rlipv2_salt = salt
            salt += 1
    if salt:
        text_encoder.config._attn_implementation = _ATTN_KEY
    return text_encoder


def route_through_dense_seam(text_encoder):
    """Send the text tower's `nn.Linear` / `nn.LayerNorm` calls through dense.py ("the
    same tcgen05 RobertaLayer kernels" for the 12 RoBERTa layers that run every step).  Module structure, parameter
    names and the arithmetic class stay HF's (x W^T + b, LayerNorm in fp32); what changes on a GPU in 'tf32' mode is
    the execution: tcgen05 forward GEMMs with the bias in the epilogue, one-pass bias-gradient column sums, fused
    LayerNorm backward, and - in the graphed step - weight / bias / LayerNorm gradients added straight into the
    flat gradient buffer instead of ~200 AccumulateGrad kernels.  On CPU tensors and in 'fp32' mode dense.py calls
    F.linear / F.layer_norm exactly as the modules did.  `RLIPV2_TEXT_DENSE=0` leaves the tower untouched."""
    if os.environ.get("RLIPV2_TEXT_DENSE", "1") == "0":
        return text_encoder
    from . import dense

    def linear_forward(self, x):
        return dense.linear(x, self.weight, self.bias, library_small=True)

    def ln_forward(self, x):
        return dense.layer_norm(x, self.weight, self.bias, self.eps)

    for m in text_encoder.modules():
        if type(m) is nn.Linear:
            m.forward = types.MethodType(linear_forward, m)
        elif type(m) is nn.LayerNorm and len(m.normalized_shape) == 1 and m.elementwise_affine:
            m.forward = types.MethodType(ln_forward, m)
    return text_encoder
