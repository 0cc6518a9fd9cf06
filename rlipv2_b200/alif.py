"""ALIF - Asymmetric Language-Image Fusion block of RLIPv2.

Mirror of the reference's
  RLIPv2_VLFuse                         /root/reference/models/fuse_helper.py:983-1097
  RLIPv2_BiAttentionBlockForCheckpoint  models/fuse_helper.py:591-752
  RLIPv2_BiMultiHeadAttention           models/fuse_helper.py:314-466
  FeatureResizer                        models/fuse_helper.py:54-73 (= models/ParSetransformer.py:1909-1928)
with identical parameter names/shapes (state_dict compatible) and identical arithmetic:

  v' = LN_v(v), l' = LN_l(l)
  q = (W_q (v' + pos) + b_q) * 256^-0.5,  k = W_k l' + b_k,  vv = W_vv v' + b,  vl = W_vl l' + b
  S[b,h] = q k^T                                  [Tv, Tl], 8 heads x head_dim 256
  P_v = softmax_rows(S)  (vision attends to language),  P_l = softmax_rows(S^T)
  dv = W_ov (P_v vl) + b,  dl = W_ol (P_l vv) + b
  v_out = v' + gamma_v[0] * dv,   l_out = l' + gamma_l[0] * dl          ("VXAc" gate)

Reference quirks reproduced on purpose:
  * both attention masks are *bool* tensors and the reference applies
    `mask.masked_fill(mask == 0, -9e15)` to them (fuse_helper.py:410-412, 425-427): on a bool tensor
    that yields all-True, i.e. the constant +1.0 is added to every logit and nothing is masked.
    Padded image cells and padded label slots therefore DO take part in both softmaxes; we keep
    that behaviour (the +1 shift itself is a softmax no-op and is not materialised);
  * the residual is taken on the LayerNorm-ed stream, which replaces the input stream
    (fuse_helper.py:685-686, 720-721);
  * dropout p=0.1 on both attention maps whenever the module is in training mode, independent of
    `--dropout` (fuse_helper.py:438-439, 1011).
"""
import os

import torch
from torch import nn

from . import dense, streams

# The block is two chains that meet only in the attention core: LayerNorm + two projections of the image tokens, the same
# for the labels, then one attention direction + out-projection + gate each.  Every kernel in them is a few hundred rows -
# latency-bound, 16-112 of 148 SMs busy - so the label chain runs on its own stream beside the image chain,
# forward and (through autograd's stream bookkeeping) backward.  RLIPV2_ALIF_STREAMS=0/1 is the A/B switch.
_ALIF_STREAMS = os.environ.get("RLIPV2_ALIF_STREAMS", "1") != "0"


class FeatureResizer(nn.Module):
    """Linear C1->C2 + LayerNorm(eps 1e-12) + dropout (fuse_helper.py:54-73)."""

    def __init__(self, input_feat_size, output_feat_size, dropout, do_ln=True):
        super().__init__()
        self.do_ln = do_ln
        self.fc = nn.Linear(input_feat_size, output_feat_size, bias=True)
        self.layer_norm = nn.LayerNorm(output_feat_size, eps=1e-12)
        self.dropout = nn.Dropout(dropout)

    def forward(self, encoder_features):
        x = dense.linear(encoder_features, self.fc.weight, self.fc.bias)
        if self.do_ln:
            x = dense.layer_norm(x, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps)
        return self.dropout(x)


class RLIPv2_BiMultiHeadAttention(nn.Module):
    """Bidirectional cross-attention sharing one score matrix (fuse_helper.py:314-466)."""

    def __init__(self, v_dim, l_dim, embed_dim, num_heads, dropout=0.1, args=None):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        self.v_dim = v_dim
        self.l_dim = l_dim
        assert self.head_dim * self.num_heads == self.embed_dim
        self.scale = self.head_dim ** (-0.5)
        self.dropout = dropout
        self.v_proj = nn.Linear(v_dim, embed_dim)
        self.l_proj = nn.Linear(l_dim, embed_dim)
        self.values_v_proj = nn.Linear(v_dim, embed_dim)
        self.values_l_proj = nn.Linear(l_dim, embed_dim)
        self.out_v_proj = nn.Linear(embed_dim, v_dim)
        self.out_l_proj = nn.Linear(embed_dim, l_dim)
        self.stable_softmax_2d = bool(getattr(args, "stable_softmax_2d", False))
        self.clamp_min_for_underflow = bool(getattr(args, "clamp_min_for_underflow", False))
        self.clamp_max_for_overflow = bool(getattr(args, "clamp_max_for_overflow", False))
        if self.stable_softmax_2d or self.clamp_min_for_underflow or self.clamp_max_for_overflow:
            raise NotImplementedError("only the defaults of main.py:209-211 (all False) are on the hot path")
        self._reset_parameters()

    def _reset_parameters(self):
        for lin in (self.v_proj, self.l_proj, self.values_v_proj, self.values_l_proj,
                    self.out_v_proj, self.out_l_proj):
            nn.init.xavier_uniform_(lin.weight)
            lin.bias.data.fill_(0)

    def forward(self, v, l, v_pos=None, attention_mask_l=None, attention_mask_v=None):
        # the reference's bool-mask handling is a no-op (module docstring); masks are accepted and ignored
        for m in (attention_mask_l, attention_mask_v):
            if m is not None and m.dtype != torch.bool:
                raise NotImplementedError("non-bool ALIF masks are not on the reference's call path")
        q_in = v if v_pos is None else v + v_pos
        # (the reference scales q before the product, `v_proj(...) * scale`; scale = 256^-0.5 is a power of two, so
        # folding it into the score is bit-identical)
        q = dense.linear(q_in, self.v_proj.weight, self.v_proj.bias)
        k = dense.linear(l, self.l_proj.weight, self.l_proj.bias)
        vv = dense.linear(v, self.values_v_proj.weight, self.values_v_proj.bias)
        vl = dense.linear(l, self.values_l_proj.weight, self.values_l_proj.bias)
        # one score matrix S = scale * q k^T per head, softmax over its rows (vision -> language) and over its columns
        # (language -> vision), dropout on both maps, two probability x value products: dense.bi_attention
        out_v, out_l = dense.bi_attention(q, k, vv, vl, self.num_heads, self.scale, self.dropout, self.training,
                                          salt=getattr(self, "_rlipv2_salt", 0))
        out_v = dense.linear(out_v, self.out_v_proj.weight, self.out_v_proj.bias)
        out_l = dense.linear(out_l, self.out_l_proj.weight, self.out_l_proj.bias)
        return out_v, out_l


class RLIPv2_BiAttentionBlockForCheckpoint(nn.Module):
    """LN -> bidirectional attention -> scalar gate + residual (fuse_helper.py:591-752)."""

    SUPPORTED_GATES = ("VXAc", "XGating", "GLIP")

    def __init__(self, v_dim, l_dim, embed_dim, num_heads, hidden_dim=None, dropout=0.1,
                 drop_path=.0, init_values=1e-4, args=None):
        super().__init__()
        self.layer_norm_v = nn.LayerNorm(v_dim)
        self.layer_norm_l = nn.LayerNorm(l_dim)
        self.attn = RLIPv2_BiMultiHeadAttention(v_dim=v_dim, l_dim=l_dim, embed_dim=embed_dim,
                                                num_heads=num_heads, dropout=dropout, args=args)
        if drop_path > 0.:
            raise NotImplementedError("drop_path > 0 is never used by the ParSeDA scripts")
        self.gamma_v = nn.Parameter(init_values * torch.ones((v_dim)), requires_grad=True)
        self.gamma_l = nn.Parameter(init_values * torch.ones((l_dim)), requires_grad=True)
        self.gating_mechanism = getattr(args, "gating_mechanism", "VXAc")
        if self.gating_mechanism not in self.SUPPORTED_GATES:
            raise NotImplementedError(
                f"gating_mechanism={self.gating_mechanism}: every ParSeDA script uses VXAc "
                "(SURVEY.md section 3.2a); the other 12 variants are out of scope")
        if getattr(args, "separate_bidirectional", False):
            raise NotImplementedError("separate_bidirectional is not used by the ParSeDA scripts")

    def forward(self, q, l, q_pos=None, attention_mask_l=None, attention_mask_v=None, dummy_tensor=None):
        new_v, new_l = self.single_attention_call(q, l, q_pos, attention_mask_l, attention_mask_v)
        return new_v, new_l, None, None, None

    def _gate(self, x, gamma, delta):
        if self.gating_mechanism == "VXAc":
            return x + gamma[0] * delta
        if self.gating_mechanism == "GLIP":
            return x + gamma * delta
        return x + delta                                     # XGating

    def _two_stream_call(self, v, l, v_pos):
        """same arithmetic as single_attention_call, the label chain on the 'alif' side stream"""
        a = self.attn
        cur = torch.cuda.current_stream(v.device)
        side = streams.get(v.device, "alif")
        salt = getattr(a, "_rlipv2_salt", 0)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            l_n = dense.layer_norm(l, self.layer_norm_l.weight, self.layer_norm_l.bias, self.layer_norm_l.eps)
            k = dense.linear(l_n, a.l_proj.weight, a.l_proj.bias)
            vl = dense.linear(l_n, a.values_l_proj.weight, a.values_l_proj.bias)
        v_n = dense.layer_norm(v, self.layer_norm_v.weight, self.layer_norm_v.bias, self.layer_norm_v.eps)
        q = dense.linear(v_n if v_pos is None else v_n + v_pos, a.v_proj.weight, a.v_proj.bias)
        vv = dense.linear(v_n, a.values_v_proj.weight, a.values_v_proj.bias)
        # each direction needs the other chain's projections
        cur.wait_stream(side)
        side.wait_stream(cur)
        for t in (k, vl):
            t.record_stream(cur)
        for t in (q, vv):
            t.record_stream(side)
        with torch.cuda.stream(side):
            out_l = dense.attention(k, q, vv, a.num_heads, a.scale, None, a.dropout, a.training, 2 * salt + 1)
            l_out = self._gate(l_n, self.gamma_l, dense.linear(out_l, a.out_l_proj.weight, a.out_l_proj.bias))
        out_v = dense.attention(q, k, vl, a.num_heads, a.scale, None, a.dropout, a.training, 2 * salt)
        v_out = self._gate(v_n, self.gamma_v, dense.linear(out_v, a.out_v_proj.weight, a.out_v_proj.bias))
        cur.wait_stream(side)
        l_out.record_stream(cur)
        return v_out, l_out

    def single_attention_call(self, v, l, v_pos, attention_mask_l=None, attention_mask_v=None, dummy_tensor=None):
        if _ALIF_STREAMS and v.is_cuda:
            for m in (attention_mask_l, attention_mask_v):
                if m is not None and m.dtype != torch.bool:
                    raise NotImplementedError("non-bool ALIF masks are not on the reference's call path")
            return self._two_stream_call(v, l, v_pos)
        v = dense.layer_norm(v, self.layer_norm_v.weight, self.layer_norm_v.bias, self.layer_norm_v.eps)
        l = dense.layer_norm(l, self.layer_norm_l.weight, self.layer_norm_l.bias, self.layer_norm_l.eps)
        delta_v, delta_l = self.attn(v, l, v_pos, attention_mask_l=attention_mask_l,
                                     attention_mask_v=attention_mask_v)
        if self.gating_mechanism == "VXAc":
            v = v + self.gamma_v[0] * delta_v
            l = l + self.gamma_l[0] * delta_l
        elif self.gating_mechanism == "GLIP":
            v = v + self.gamma_v * delta_v
            l = l + self.gamma_l * delta_l
        else:  # XGating
            v = v + delta_v
            l = l + delta_l
        return v, l


class RLIPv2_VLFuse(nn.Module):
    """Dict-in / dict-out wrapper the encoder calls (fuse_helper.py:983-1097)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.lang_model = getattr(args, "text_encoder_type", "roberta-base")
        self.joint_embedding_size = 256
        self.n_head = 8
        self.embed_dim = 2048
        self.t2i_hidden_dim = 1024
        self.i2t_hidden_dim = 3072
        self.lang_dim = 768 if self.lang_model in ["bert-base-uncased", "roberta-base", "clip"] else 1024
        if getattr(args, "fusion_type", "GLIP_attn") != "GLIP_attn":
            raise NotImplementedError("only fusion_type=GLIP_attn is on the ParSeDA hot path")
        self.b_attn = RLIPv2_BiAttentionBlockForCheckpoint(
            v_dim=self.joint_embedding_size, l_dim=self.lang_dim, embed_dim=self.embed_dim,
            num_heads=self.n_head, hidden_dim=self.i2t_hidden_dim, dropout=0.1, drop_path=.0,
            init_values=1.0 / args.num_feature_levels, args=args)

    def forward(self, x):
        vis, lang = x["visual"], x["lang"]
        q, l0, _, _, _ = self.b_attn(q=vis["src"], l=lang["hidden"], q_pos=vis["pos"],
                                     attention_mask_l=lang["masks"], attention_mask_v=vis["padding_mask"])
        vis["src"] = q
        lang["hidden"] = l0
        return {"visual": vis, "lang": lang}
