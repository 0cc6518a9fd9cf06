"""ParSeDA deformable encoder / decoder with ALIF (language-image fusion) - module layer.

Mirrors, with identical parameter names (state_dict compatible) and arithmetic:
  RLIP_ParSeDABDeformableTransformer_v2   /root/reference/models/dab_deformable/deformable_transformer.py:234-744
  RLIPv2_DeformableTransformerEncoder     models/deformable_transformer.py:791-884
  DeformableTransformerEncoderLayer       dab_deformable/deformable_transformer.py:1261-1300
  DeformableTransformerDecoderLayer       dab_deformable/deformable_transformer.py:1346-1401
  DABDeformableTransformerDecoderHOI      dab_deformable/deformable_transformer.py:1404-1552
  MultiBranchFusion                       dab_deformable/deformable_transformer.py:1025-1068
  MLP / gen_sineembed_for_position        dab_deformable/deformable_transformer.py:1763-1802

B200-first differences that do not change results:
  * level shapes travel as a python list next to the device tensor, so none of the reference's
    per-call device->host syncs remain (`assert ... .sum() == Len_in`, ms_deform_attn.py:96;
    iterating a CUDA `spatial_shapes`, models/deformable_transformer.py:805; `level_start_index[-1]`
    as a slice bound, :829,845);
  * MultiBranchFusion runs its 16 branches as three stacked GEMMs instead of 48 tiny ones;
  * the decoder's self-attention keeps nn.MultiheadAttention's parameter names but is computed with
    one fused in-projection;
  * every dense contraction goes through `rlipv2_b200.dense` (the seam the tcgen05 kernels plug into).
"""
import copy
import math
import os

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import normal_
from torch.nn.utils.rnn import pad_sequence

from . import dense, grad_ready, streams
from .alif import FeatureResizer, RLIPv2_VLFuse
from .ms_deform_attn import MSDeformAttn
from .nested import inverse_sigmoid
from .roberta_layer import RobertaLayer
from .text_encoder import build_text_encoder, pooled_text, pooled_text_sharded


_LEVEL_CACHE = {}
_LANG_STREAM = os.environ.get("RLIPV2_LANG_STREAM", "1") != "0"       # A/B switch for measurements


def _level_tensors(shapes_host, device):
    """(spatial_shapes [L,2], level_start_index [L]) int64 device tensors, cached per shape set: they are
    constants of the image size, and creating them per call costs two pageable H2D copies."""
    key = (shapes_host, str(device))
    if key not in _LEVEL_CACHE:
        starts = [0]
        for h, w in shapes_host[:-1]:
            starts.append(starts[-1] + h * w)
        _LEVEL_CACHE[key] = (torch.as_tensor(shapes_host, dtype=torch.long, device=device),
                             torch.as_tensor(starts, dtype=torch.long, device=device))
    return _LEVEL_CACHE[key]


_VALUE_STREAM = os.environ.get("RLIPV2_VALUE_STREAM", "1") != "0"   # decoder value projections on a side stream


def _get_clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


class MLP(nn.Module):
    """Linear -> ReLU -> ... -> Linear (dab_deformable/deformable_transformer.py:1763-1775)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            if i < self.num_layers - 1:
                x = dense.linear_relu(x, layer.weight, layer.bias)
            else:
                x = dense.linear(x, layer.weight, layer.bias)
        return x


def gen_sineembed_for_position(pos_tensor):
    """[bs, nq, 2|4] normalised (x, y[, w, h]) -> [bs, nq, 256|512] sine embedding in the order
    (y, x[, w, h]); 128 features each, temperature 10000 (deformable_transformer.py:1777-1802).
    The reference embeds one coordinate at a time (6 kernels each); here all coordinates go through the
    same elementwise ops at once - the values are identical, the launch count drops from ~28 to 7."""
    return dense.sine_embed(pos_tensor)


class MultiBranchFusion(nn.Module):
    """relu(sum_c fc_3[c](relu(fc_1[c](a) * fc_2[c](b)))) with `cardinality` branches
    (deformable_transformer.py:1025-1068).  Parameters stay per-branch (checkpoint layout); the
    forward stacks them into three GEMMs."""

    def __init__(self, appearance_size, spatial_size, representation_size, cardinality):
        super().__init__()
        self.cardinality = cardinality
        sub = int(representation_size / cardinality)
        assert sub * cardinality == representation_size
        self.fc_1 = nn.ModuleList([nn.Linear(appearance_size, sub) for _ in range(cardinality)])
        self.fc_2 = nn.ModuleList([nn.Linear(spatial_size, sub) for _ in range(cardinality)])
        self.fc_3 = nn.ModuleList([nn.Linear(sub, representation_size) for _ in range(cardinality)])

    def forward(self, appearance, spatial):
        w1, b1 = _StackedBranchParams.apply(0, *[t for m in self.fc_1 for t in (m.weight, m.bias)])
        w2, b2 = _StackedBranchParams.apply(0, *[t for m in self.fc_2 for t in (m.weight, m.bias)])
        w3, b3 = _StackedBranchParams.apply(1, *[t for m in self.fc_3 for t in (m.weight, m.bias)])
        h = F.relu(dense.linear(appearance, w1, b1) * dense.linear(spatial, w2, b2))
        return dense.linear_relu(h, w3, b3)


class _StackedBranchParams(torch.autograd.Function):
    """(weight, bias) of `cardinality` per-branch nn.Linear modules -> the operands of one stacked GEMM.
    mode 0: branches side by side along the output features: W [n*o, i] = cat(w_c, 0), b [n*o] = cat(b_c)
    mode 1: branches summed: W [o, n*i] = cat(w_c, 1), b [o] = sum_c b_c
    Backward: autograd would hand 2n slices to 2n AccumulateGrad nodes (96 tiny `+=` kernels for the verb-query
    generator).  When the parameters' .grad are adjacent views of the step's flat gradient buffer (train_step marks
    them `_fuse_grad`; layout w_0 | b_0 | w_1 | b_1 | ...), the same sums are two strided adds."""

    @staticmethod
    def forward(ctx, mode, *params):
        ws, bs = params[0::2], params[1::2]
        ctx.mode, ctx.n = mode, len(ws)
        ctx.params = params
        if mode == 0:
            return torch.cat(ws, 0), torch.cat(bs, 0)
        return torch.cat(ws, 1), torch.stack(bs, 0).sum(0)

    @staticmethod
    def _adjacent_grad_views(params):
        """[n, numel(w) + numel(b)] strided view over the parameters' flat gradient range, or None"""
        ws, bs = params[0::2], params[1::2]
        if not all(getattr(p, "_fuse_grad", False) and p.grad is not None and p.grad.is_contiguous() for p in params):
            return None
        nw, nb = ws[0].numel(), bs[0].numel()
        if any(w.numel() != nw for w in ws) or any(b.numel() != nb for b in bs):
            return None
        base = ws[0].grad
        ptr = base.data_ptr()
        for p in params:
            if p.grad.data_ptr() != ptr or p.grad.dtype != base.dtype:
                return None
            ptr += p.numel() * base.element_size()
        return base.as_strided((len(ws), nw + nb), (nw + nb, 1))

    @staticmethod
    def backward(ctx, gw, gb):
        n, params = ctx.n, ctx.params
        w0, b0 = params[0], params[1]
        o, i = w0.shape
        if ctx.mode == 0:
            gws = gw.reshape(n, o * i)                              # branch c = rows c*o .. (c+1)*o
            gbs = gb.reshape(n, o)
        else:
            gws = gw.reshape(o, n, i).transpose(0, 1).reshape(n, o * i)
            gbs = gb.reshape(1, o).expand(n, o)
        region = _StackedBranchParams._adjacent_grad_views(params)
        if region is not None:
            region[:, :o * i].add_(gws)
            region[:, o * i:].add_(gbs)
            return (None,) + (None,) * len(params)
        grads = []
        for c in range(n):
            grads += [gws[c].reshape(o, i), gbs[c].reshape(o)]
        return (None,) + tuple(grads)


class DeformableTransformerEncoderLayer(nn.Module):
    """MSDeformAttn self-attention + add&LN + FFN + add&LN (deformable_transformer.py:1261-1300)."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if activation != "relu":
            raise NotImplementedError("ParSeDA scripts use relu")
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    def forward_ffn(self, src):
        if self.dropout2.p == 0 or not self.training:
            src2 = dense.ffn_relu(src, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias)
        else:
            h = self.dropout2(dense.linear_relu(src, self.linear1.weight, self.linear1.bias))
            src2 = dense.linear(h, self.linear2.weight, self.linear2.bias)
        src2 = self.dropout3(src2)
        return dense.add_layer_norm(src2, src, self.norm2.weight, self.norm2.bias, self.norm2.eps)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None,
                spatial_shapes_host=None):
        q = src if pos is None else src + pos
        src2 = self.self_attn(q, reference_points, src, spatial_shapes, level_start_index, padding_mask,
                              spatial_shapes_host=spatial_shapes_host)
        src = dense.add_layer_norm(self.dropout1(src2), src, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        return self.forward_ffn(src)


class RLIPv2_DeformableTransformerEncoder(nn.Module):
    """6 deformable layers; before layers 0, 2, 4 the coarsest level's tokens and the label
    embeddings are fused by ALIF and the labels pass one RobertaLayer
    (models/deformable_transformer.py:791-884)."""

    def __init__(self, encoder_layer, roberta_layer, VLFuse_layer, num_layers, fusion_interval=2,
                 fusion_last_vis=False, lang_aux_loss=False):
        super().__init__()
        self.layers = _get_clones(encoder_layer, num_layers)
        self.num_layers = num_layers
        self.fusion_interval = fusion_interval
        self.roberta_layers = _get_clones(roberta_layer, num_layers // fusion_interval)
        self.VLFuse_layers = _get_clones(VLFuse_layer, num_layers // fusion_interval)
        self.fusion_last_vis = fusion_last_vis
        self.lang_aux_loss = lang_aux_loss
        # call-site ids of the fused attention kernels' hashed dropout (one stream of masks per attention map)
        for i, m in enumerate(self.VLFuse_layers):
            m.b_attn.attn._rlipv2_salt = 0x100 + i
        for i, m in enumerate(self.roberta_layers):
            m.attention.self._rlipv2_salt = 0x400 + i

    @staticmethod
    def get_reference_points(spatial_shapes_host, valid_ratios, device):
        """cell centres / valid extent, then scaled to every level (:803-815) -> [bs, S, L, 2]"""
        refs = []
        for lvl, (H_, W_) in enumerate(spatial_shapes_host):
            ref_y, ref_x = torch.meshgrid(
                torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device),
                torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device), indexing="ij")
            ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
            ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
            refs.append(torch.stack((ref_x, ref_y), -1))
        reference_points = torch.cat(refs, 1)
        return reference_points[:, :, None] * valid_ratios[:, None]

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None,
                lang_hidden=None, lang_masks=None, spatial_shapes_host=None):
        if spatial_shapes_host is None:     # reference-compatible call: one sync to learn the shapes
            spatial_shapes_host = [tuple(int(v) for v in hw) for hw in spatial_shapes.tolist()]
        last_start = sum(h * w for h, w in spatial_shapes_host[:-1])
        reference_points = self.get_reference_points(spatial_shapes_host, valid_ratios, src.device)
        inv_padding_mask = ~padding_mask
        inv_lang_masks = ~lang_masks
        if self.fusion_last_vis:
            vis = {"src": src, "padding_mask": inv_padding_mask[:, last_start:], "pos": pos[:, last_start:]}
        else:
            vis = {"src": src, "padding_mask": inv_padding_mask, "pos": pos}
        lang = {"hidden": lang_hidden, "masks": inv_lang_masks}
        multi_lay_lang = []
        if self.training and src.is_cuda:
            dense.advance_dropout_seed(src.device)          # fresh masks for the fused attention kernels' hashed dropout
        side, pending = None, False
        if _LANG_STREAM and src.is_cuda:
            if getattr(self, "_lang_stream", None) is None:
                self._lang_stream = streams.get(src.device, "lang")
            side = self._lang_stream
        for idx, layer in enumerate(self.layers):
            if idx % self.fusion_interval == 0:
                k = idx // self.fusion_interval
                if self.fusion_last_vis:
                    full_src = vis["src"]
                    vis["src"] = full_src[:, last_start:]
                fused = self.VLFuse_layers[k]({"visual": vis, "lang": lang})
                vis, lang = fused["visual"], fused["lang"]
                if self.fusion_last_vis:
                    # write the fused coarsest level back (the reference does it in place, :856-859)
                    vis["src"] = torch.cat((full_src[:, :last_start], vis["src"]), dim=1)
                if side is not None:
                    # the label stream's RobertaLayer does not touch the image tokens: run it (and, through
                    # autograd's stream bookkeeping, its backward) beside the next `fusion_interval`
                    # deformable layers instead of in front of them
                    cur = torch.cuda.current_stream(src.device)
                    side.wait_stream(cur)
                    lang_in = lang["hidden"]
                    with torch.cuda.stream(side):
                        lang["hidden"] = self.roberta_layers[k](lang_in, attention_mask=lang["masks"])
                    # the fused label stream was allocated on `cur` and is read on `side`: without autograd holding it
                    # (torch.no_grad inference) it is freed right here and `cur`'s next allocation would reuse the block
                    # while the RobertaLayer still reads it (round 1's graphed-inference test failure)
                    lang_in.record_stream(side)
                    lang["masks"].record_stream(side)
                    pending = True
                else:
                    lang["hidden"] = self.roberta_layers[k](lang["hidden"], attention_mask=lang["masks"])
                multi_lay_lang.append(lang["hidden"])
            vis["src"] = layer(vis["src"], pos, reference_points, spatial_shapes, level_start_index,
                               padding_mask, spatial_shapes_host=spatial_shapes_host)
            if pending and ((idx + 1) % self.fusion_interval == 0 or idx + 1 == self.num_layers):
                cur = torch.cuda.current_stream(src.device)           # join before the labels are used again
                cur.wait_stream(side)
                lang["hidden"].record_stream(cur)
                pending = False
        if self.lang_aux_loss:
            if self.fusion_interval == 2:
                multi_lay_lang = torch.stack(multi_lay_lang, dim=0)
            elif self.fusion_interval == 1:
                multi_lay_lang = torch.stack(multi_lay_lang[::2], dim=0)
        else:
            multi_lay_lang = multi_lay_lang[-1]
        return vis["src"], multi_lay_lang


class DeformableTransformerEncoder(nn.Module):
    """the plain deformable encoder (no early fusion) of the `--fusion_type MDETR_attn / no_fusion` ablations
    (dab_deformable/deformable_transformer.py:1303-1343)"""

    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = _get_clones(encoder_layer, num_layers)
        self.num_layers = num_layers

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None,
                spatial_shapes_host=None):
        if spatial_shapes_host is None:
            spatial_shapes_host = [tuple(int(v) for v in hw) for hw in spatial_shapes.tolist()]
        reference_points = RLIPv2_DeformableTransformerEncoder.get_reference_points(spatial_shapes_host, valid_ratios,
                                                                                    src.device)
        for layer in self.layers:
            src = layer(src, pos, reference_points, spatial_shapes, level_start_index, padding_mask,
                        spatial_shapes_host=spatial_shapes_host)
        return src


class QuerySelfAttention(nn.Module):
    """8x32 self-attention among the queries with nn.MultiheadAttention's parameter layout
    (`in_proj_weight [3C, C]`, `in_proj_bias`, `out_proj`), batch-first."""

    def __init__(self, embed_dim, num_heads, dropout=0.0):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.)

    def forward(self, qk_input, v_input):
        b, t, c = qk_input.shape
        h, d = self.num_heads, c // self.num_heads
        qk = dense.linear(qk_input, self.in_proj_weight[:2 * c], self.in_proj_bias[:2 * c])
        v = dense.linear(v_input, self.in_proj_weight[2 * c:], self.in_proj_bias[2 * c:])
        q, k = qk[..., :c], qk[..., c:]                       # column slices of the fused projection, used in place
        # softmax(q k^T / sqrt(d)) -> dropout -> . v, 8 heads x 32 (nn.MultiheadAttention, :1383-1390): dense.attention
        o = dense.attention(q, k, v, h, d ** -0.5, None, self.dropout, self.training, salt=getattr(self, "_rlipv2_salt", 0))
        return dense.linear(o, self.out_proj.weight, self.out_proj.bias)


class DeformableTransformerDecoderLayer(nn.Module):
    """query self-attention + LN, MSDeformAttn cross-attention + LN, FFN + LN (:1346-1401)."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8,
                 n_points=4, do_self_attn=True):
        super().__init__()
        if activation != "relu":
            raise NotImplementedError("ParSeDA scripts use relu")
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.do_self_attn = do_self_attn
        if do_self_attn:
            self.self_attn = QuerySelfAttention(d_model, n_heads, dropout=dropout)
            self.dropout2 = nn.Dropout(dropout)
            self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)

    def forward(self, tgt, query_pos, reference_points, src, src_spatial_shapes, level_start_index,
                src_padding_mask=None, spatial_shapes_host=None, value=None):
        if self.do_self_attn:
            qk = tgt if query_pos is None else tgt + query_pos
            tgt2 = self.self_attn(qk, tgt)
            tgt = dense.add_layer_norm(self.dropout2(tgt2), tgt, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        q = tgt if query_pos is None else tgt + query_pos
        tgt2 = self.cross_attn(q, reference_points, src, src_spatial_shapes, level_start_index,
                               src_padding_mask, spatial_shapes_host=spatial_shapes_host, value=value)
        tgt = dense.add_layer_norm(self.dropout1(tgt2), tgt, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        h = self.dropout3(dense.linear_relu(tgt, self.linear1.weight, self.linear1.bias))
        tgt2 = self.dropout4(dense.linear(h, self.linear2.weight, self.linear2.bias))
        return dense.add_layer_norm(tgt2, tgt, self.norm3.weight, self.norm3.bias, self.norm3.eps)


class DABDeformableTransformerDecoderHOI(nn.Module):
    """DAB (dynamic anchor box) decoder loop over subject/object anchor boxes (:1404-1552).
    ParSe=True: pair decoder, queries = [subjects ; objects], each half refines its own boxes.
    ParSe=False: verb decoder, one query per pair, anchored at the mean of its two boxes."""

    def __init__(self, decoder_layer, num_layers, return_intermediate=False, use_dab=False, d_model=256,
                 high_dim_query_update=False, no_sine_embed=False, ParSe=False):
        super().__init__()
        assert use_dab and not high_dim_query_update and not no_sine_embed, \
            "ParSeDA builds both decoders with use_dab=True only (transformer.py:1344-1362)"
        self.layers = _get_clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.return_intermediate = return_intermediate
        self.sub_bbox_embed = None
        self.obj_bbox_embed = None
        self.class_embed = None
        self.use_dab = use_dab
        self.d_model = d_model
        self.query_scale = MLP(d_model, d_model, d_model, 2)
        self.ref_point_head = MLP(2 * d_model, d_model, d_model, 2)
        self.ParSe = ParSe

    def forward(self, tgt, reference_points, src, src_spatial_shapes, src_level_start_index, src_valid_ratios,
                query_pos=None, src_padding_mask=None, spatial_shapes_host=None):
        assert query_pos is None
        output = tgt
        bs = src.shape[0]
        sub_ref, obj_ref = reference_points
        if self.ParSe:
            sub_ref = sub_ref[None].repeat(bs, 1, 1)
            obj_ref = obj_ref[None].repeat(bs, 1, 1)
        assert sub_ref.shape[-1] == 4 and obj_ref.shape[-1] == 4
        pair_num = obj_ref.shape[1]
        vr4 = torch.cat([src_valid_ratios, src_valid_ratios], -1)[:, None]       # [bs, 1, L, 4]
        inter, inter_sub, inter_obj = [], [], []
        # Every layer's cross-attention projects the same encoder memory with its own value_proj (44k rows: the only
        # large GEMM of the decoder, ~30 us forward and ~80 us backward per layer).  None of them depends on the
        # queries, so they run on a side stream beside the chain of small per-query kernels (forward here, and -
        # autograd replays nodes on the stream of their forward - the backward too).
        values = [None] * len(self.layers)
        if _VALUE_STREAM and src.is_cuda:
            cur = torch.cuda.current_stream(src.device)
            if getattr(self, "_value_stream", None) is None:
                self._value_stream = streams.get(src.device, "value_pair" if self.ParSe else "value_verb")
            side = self._value_stream
            side.wait_stream(cur)
            values, value_ready = [], []
            with torch.cuda.stream(side):
                for layer in self.layers:
                    values.append(layer.cross_attn.project_value(src, src_padding_mask))
                    ev = torch.cuda.Event()
                    ev.record(side)
                    value_ready.append(ev)
            for v in values:
                v.record_stream(cur)
        # refined boxes with their autograd history: the pair decoder's are exactly the model's box
        # predictions (hoi.py:2122-2141 recomputes the same MLP on the same inputs), so the head reuses them
        refined = self.refined_boxes = []
        for lid, layer in enumerate(self.layers):
            if self.ParSe:
                ref_input = torch.cat((sub_ref[:, :, None] * vr4, obj_ref[:, :, None] * vr4), dim=1)
            else:
                ref_input = 0.5 * (sub_ref + obj_ref)[:, :, None] * vr4
            raw_query_pos = self.ref_point_head(gen_sineembed_for_position(ref_input[:, :, 0, :]))
            query_pos_l = raw_query_pos if lid == 0 else self.query_scale(output) * raw_query_pos
            if values[lid] is not None:
                torch.cuda.current_stream(src.device).wait_event(value_ready[lid])
            output = layer(output, query_pos_l, ref_input, src, src_spatial_shapes, src_level_start_index,
                           src_padding_mask, spatial_shapes_host=spatial_shapes_host, value=values[lid])
            # iterative box refinement; the refined anchors are detached (:1511-1541)
            if self.sub_bbox_embed is not None:
                sub_in = output[:, :pair_num] if self.ParSe else output
                sub_box = dense.box_refine(self.sub_bbox_embed[lid](sub_in), sub_ref)
                sub_ref = sub_box.detach()
            if self.obj_bbox_embed is not None:
                obj_in = output[:, pair_num:] if self.ParSe else output
                obj_box = dense.box_refine(self.obj_bbox_embed[lid](obj_in), obj_ref)
                obj_ref = obj_box.detach()
            if self.sub_bbox_embed is not None and self.obj_bbox_embed is not None:
                refined.append((sub_box, obj_box))
            if self.return_intermediate:
                inter.append(output)
                inter_sub.append(sub_ref)
                inter_obj.append(obj_ref)
        if self.return_intermediate:
            refs = torch.stack((torch.stack(inter_sub), torch.stack(inter_obj)), dim=0).transpose(0, 1)
            return torch.stack(inter), refs
        return output, reference_points


class RLIP_ParSeDABDeformableTransformer_v2(nn.Module):
    """Two-phase transformer of RLIPv2-ParSeDA (deformable_transformer.py:234-744).
    Phase A (`encode_and_save=True`): flatten levels, encode label strings, ALIF encoder ->
    memory_cache dict.  Phase B: pair decoder -> verb queries (MBF) -> verb decoder."""

    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=1024,
                 dropout=0.1, activation="relu", return_intermediate_dec=False, num_feature_levels=4,
                 dec_n_points=4, enc_n_points=4, two_stage=False, two_stage_num_proposals=300, use_dab=False,
                 high_dim_query_update=False, no_sine_embed=False, pass_pos_and_query=True,
                 text_encoder_type="roberta-base", freeze_text_encoder=False, args=None):
        super().__init__()
        if two_stage or not use_dab:
            raise NotImplementedError("ParSeDA is built with use_dab=True, two_stage=False (transformer.py:1344-1362)")
        self.d_model, self.nhead = d_model, nhead
        self.two_stage = two_stage
        self.two_stage_num_proposals = two_stage_num_proposals
        self.use_dab = use_dab
        self.fusion_type = args.fusion_type
        if self.fusion_type not in ("GLIP_attn", "MDETR_attn", "no_fusion"):
            raise ValueError(f"unknown --fusion_type {self.fusion_type}")
        if self.fusion_type != "GLIP_attn":
            # ablations of the paper (scripts/RLIP_ParSeDA/*_MDETR.sh): plain deformable encoder (:252-256); module creation
            # order follows the reference so that seeded initialisations agree
            self.encoder = DeformableTransformerEncoder(
                DeformableTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels, nhead,
                                                  enc_n_points), num_encoder_layers)

        ho_layer = DeformableTransformerDecoderLayer(d_model, dim_feedforward, dropout, activation,
                                                     num_feature_levels, nhead, dec_n_points)
        self.ho_decoder = DABDeformableTransformerDecoderHOI(ho_layer, num_decoder_layers, return_intermediate_dec,
                                                             use_dab=use_dab, d_model=d_model, ParSe=True)
        verb_layer = DeformableTransformerDecoderLayer(d_model, dim_feedforward, dropout, activation,
                                                       num_feature_levels, nhead, dec_n_points, do_self_attn=True)
        self.verb_decoder = DABDeformableTransformerDecoderHOI(verb_layer, num_decoder_layers, return_intermediate_dec,
                                                               use_dab=use_dab, d_model=d_model, ParSe=False)
        self.verb_tgt_generator = MultiBranchFusion(256, 256, 256, 16)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))

        from .text_encoder import roberta_base_config
        if self.fusion_type == "MDETR_attn":
            # late fusion: two pre-norm encoder stacks over (decoder states ; label embeddings) (:278-291)
            from .parse_detr import CrossModelTransformerEncoder, TransformerEncoderLayer
            self.obj_fusion = CrossModelTransformerEncoder(
                TransformerEncoderLayer(d_model, 8, dim_feedforward, dropout, activation, normalize_before=True),
                num_decoder_layers, nn.LayerNorm(d_model), return_intermediate=True)
            self.verb_fusion = CrossModelTransformerEncoder(
                TransformerEncoderLayer(d_model, 8, dim_feedforward, dropout, activation, normalize_before=True),
                num_decoder_layers, nn.LayerNorm(d_model), return_intermediate=True)
        elif self.fusion_type == "GLIP_attn":
            enc_layer = DeformableTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation,
                                                          num_feature_levels, nhead, enc_n_points)
            self.encoder = RLIPv2_DeformableTransformerEncoder(
                enc_layer, RobertaLayer(roberta_base_config()), RLIPv2_VLFuse(args), num_encoder_layers,
                fusion_interval=args.fusion_interval, fusion_last_vis=args.fusion_last_vis,
                lang_aux_loss=args.lang_aux_loss)

        self._reset_parameters()

        self.pass_pos_and_query = pass_pos_and_query
        self.tokenizer, self.text_encoder = build_text_encoder(
            text_encoder_type, synthetic=getattr(args, "synthetic_text_encoder", None))
        from .text_encoder import route_through_dense_seam, use_short_attention
        route_through_dense_seam(self.text_encoder)
        use_short_attention(self.text_encoder)
        if freeze_text_encoder:
            for p in self.text_encoder.parameters():
                p.requires_grad_(False)
        self.expander_dropout = 0.1
        self.resizer = FeatureResizer(input_feat_size=self.text_encoder.config.hidden_size,
                                      output_feat_size=d_model, dropout=self.expander_dropout)
        self.verb_query_tgt_type = args.verb_query_tgt_type
        if "MBF" in self.verb_query_tgt_type:
            self.verb_tgt_generator = MultiBranchFusion(256, 256, 256, 16)
        self.shard_labels, self.label_shard_group = False, None

    def shard_label_text(self, enabled=True, group=None):
        """SURVEY 8f rank 3 (label set de-duplicated across ranks): when EVERY rank passes the same label strings to every
        step - fine-tuning on a dataset's fixed object / relation vocabulary - each rank runs the text tower on 1 / world of
        them and the pooled vectors are all-gathered (text_encoder.pooled_text_sharded).  The caller owns that promise; label
        sets that differ per rank (relational pre-training with per-batch negatives, engine.py:92-98) must leave this off."""
        self.shard_labels, self.label_shard_group = bool(enabled), group
        return self

    def _reset_parameters(self):
        # xavier on every matrix built so far - including the RobertaLayers / ALIF blocks, but not the
        # text encoder and resizer, which are attached afterwards (:364-374)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        normal_(self.level_embed)

    @staticmethod
    def get_valid_ratio(mask):
        _, H, W = mask.shape
        valid_H = torch.sum(~mask[:, :, 0], 1)
        valid_W = torch.sum(~mask[:, 0, :], 1)
        return torch.stack([valid_W.float() / W, valid_H.float() / H], -1)

    # ---- text ---------------------------------------------------------------------------------
    def tokenize(self, text, device):
        """Host part of the text path: label strings -> token ids on the device (:489-497).  Returns a
        dict that `forward(text=...)` accepts in place of the raw strings, so a caller can tokenise
        once outside a CUDA graph."""
        sums, flat = [], []
        for obj_names, pred_names in text:
            sums.append((len(obj_names), len(pred_names)))
            flat += list(obj_names) + list(pred_names)
        tok = self.tokenizer.batch_encode_plus(flat, padding="longest", return_tensors="pt")
        return {"input_ids": tok["input_ids"].to(device), "attention_mask": tok["attention_mask"].to(device),
                "sums": sums}

    def encode_text(self, text, device):
        """label strings -> pooled RoBERTa vectors, padded per image (:489-522).
        -> text_memory [n_text, n_tuples, 768], text_attention_mask [n_text, n_tuples] (True = pad),
           obj_pred_names_sums [n_tuples, 2]"""
        tok = text if isinstance(text, dict) else self.tokenize(text, device)
        sums = tok["sums"]
        obj_pred_names_sums = torch.tensor(sums)
        if getattr(self, "shard_labels", False):
            # every rank was promised the same label set: encode 1 / world of it here (text_encoder.pooled_text_sharded)
            pooled = pooled_text_sharded(self.text_encoder, tok["input_ids"], tok["attention_mask"], self.label_shard_group)
        else:
            pooled = pooled_text(self.text_encoder, tok["input_ids"], tok["attention_mask"])
        pooled = grad_ready.mark(pooled, "text")         # no-op unless the data-parallel step installed a callback
        i, objs, preds = 0, [], []
        for n_obj, n_pred in sums:
            objs.append(pooled[i:i + n_obj])
            preds.append(pooled[i + n_obj:i + n_obj + n_pred])
            i += n_obj + n_pred
        text_memory = torch.cat([pad_sequence(objs), pad_sequence(preds)], dim=0)
        text_attention_mask = ~(text_memory.sum(dim=-1) > 0)
        return text_memory, text_attention_mask, obj_pred_names_sums

    @staticmethod
    def _is_label_text(text):
        return isinstance(text, dict) or (isinstance(text, list) and len(text) > 0 and isinstance(text[0], tuple))

    def encode_text_async(self, text, device):
        """Start `encode_text` on a side stream and return a handle that `forward(text=handle)` joins where
        the label embeddings are first needed.  The text tower (12 RoBERTa layers on 3-8 token strings) is
        ~350 launch-bound kernels forward and ~700 backward that do not depend on the image: forked before
        the backbone they run under the convolutions instead of after them - and autograd replays the
        backward of these nodes on the same side stream, i.e. under the backbone's backward.  Works eagerly
        and inside a CUDA-graph capture (the fork/join become graph edges)."""
        cur = torch.cuda.current_stream(device)
        if getattr(self, "_text_stream", None) is None:
            self._text_stream = streams.get(device, "text")
        side = self._text_stream
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            text_memory, text_attention_mask, sums = self.encode_text(text, device)
        return {"_async_text": (text_memory, text_attention_mask, sums), "_stream": side}

    @staticmethod
    def _join_text(handle, device):
        cur = torch.cuda.current_stream(device)
        cur.wait_stream(handle["_stream"])
        text_memory, text_attention_mask, sums = handle["_async_text"]
        for t in (text_memory, text_attention_mask):
            t.record_stream(cur)
        return text_memory, text_attention_mask, sums

    # ---- forward ------------------------------------------------------------------------------
    def forward(self, srcs=None, masks=None, pos_embeds=None, query_embed=None, text=None, encode_and_save=True,
                text_memory=None, img_memory=None, text_attention_mask=None, obj_pred_names_sums=None,
                spatial_shapes=None, level_start_index=None, valid_ratios=None, spatial_shapes_host=None):
        assert query_embed is not None
        if encode_and_save:
            return self._encode(srcs, masks, pos_embeds, query_embed, text)
        return self._decode(masks, query_embed, text_memory, img_memory, spatial_shapes, level_start_index,
                            valid_ratios, spatial_shapes_host, text_attention_mask, obj_pred_names_sums)

    def _encode(self, srcs, masks, pos_embeds, query_embed, text):
        src_flatten, mask_flatten, lvl_pos_flatten, shapes_host = [], [], [], []
        for lvl, (src, mask, pos_embed) in enumerate(zip(srcs, masks, pos_embeds)):
            bs, c, h, w = src.shape
            shapes_host.append((h, w))
            src_flatten.append(src.flatten(2).transpose(1, 2))
            mask_flatten.append(mask.flatten(1))
            lvl_pos_flatten.append(pos_embed.flatten(2).transpose(1, 2) + self.level_embed[lvl].view(1, 1, -1))
        # levels that are already rows of one token buffer (parseda.py FlatLevels) need no concatenation
        pre_flat = getattr(srcs, "flat", None)
        src_flatten = pre_flat if pre_flat is not None else torch.cat(src_flatten, 1)
        mask_flatten = torch.cat(mask_flatten, 1)
        lvl_pos_flatten = torch.cat(lvl_pos_flatten, 1)
        device = src_flatten.device
        spatial_shapes, level_start_index = _level_tensors(tuple(shapes_host), device)
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)

        if isinstance(text, dict) and "_async_text" in text:            # training, text tower already in flight
            text_memory, text_attention_mask, obj_pred_names_sums = self._join_text(text, device)
            text = None
        elif self._is_label_text(text):                                     # training: label strings
            text_memory, text_attention_mask, obj_pred_names_sums = self.encode_text(text, device)
            text = None
        if self.fusion_type != "GLIP_attn":
            # no early fusion (:552-562, :588-595): image tokens through the plain encoder, labels resized as they are
            img_memory = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, lvl_pos_flatten,
                                      mask_flatten, spatial_shapes_host=shapes_host)
            if text is None:
                text_memory_resized = self.resizer(text_memory)
                if text_memory_resized.shape[1] != bs:
                    text_memory_resized = text_memory_resized.repeat(1, bs, 1)
                    text_attention_mask = text_attention_mask.repeat(1, bs)
            else:                                                           # eval: already resized by the caller
                text_attention_mask, text_memory_resized, obj_pred_names_sums = text
                text_memory = text_memory_resized
            return self._memory_cache(text_memory, text_memory_resized, img_memory, mask_flatten, text_attention_mask,
                                      lvl_pos_flatten, query_embed, obj_pred_names_sums, spatial_shapes,
                                      level_start_index, valid_ratios, shapes_host)
        if text is None:
            lang = text_memory
            if lang.shape[1] != bs:
                lang = lang.repeat(1, bs, 1)
                text_attention_mask = text_attention_mask.repeat(1, bs)
        else:                                                               # eval: pre-encoded text
            text_attention_mask, text_memory, obj_pred_names_sums = text
            lang = text_memory
        img_memory, lang_out = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios,
                                            lvl_pos_flatten, mask_flatten, lang_hidden=lang.transpose(0, 1),
                                            lang_masks=text_attention_mask.transpose(0, 1),
                                            spatial_shapes_host=shapes_host)
        if lang_out.dim() == 3:
            text_memory_resized = self.resizer(lang_out.transpose(0, 1))
        else:                                                               # [3, bs, Tl, 768] with lang_aux_loss
            text_memory_resized = self.resizer(lang_out.transpose(1, 2))
        return self._memory_cache(text_memory, text_memory_resized, img_memory, mask_flatten, text_attention_mask,
                                  lvl_pos_flatten, query_embed, obj_pred_names_sums, spatial_shapes, level_start_index,
                                  valid_ratios, shapes_host)

    @staticmethod
    def _memory_cache(text_memory, text_memory_resized, img_memory, mask_flatten, text_attention_mask, lvl_pos_flatten,
                      query_embed, obj_pred_names_sums, spatial_shapes, level_start_index, valid_ratios, shapes_host):
        """the phase-A dictionary (:598-613)"""
        return {
            "text_memory_bf_resize": text_memory,
            "text_memory_resized": text_memory_resized,
            "text_memory": text_memory_resized,
            "img_memory": img_memory,
            "masks": mask_flatten,
            "text_attention_mask": text_attention_mask,
            "pos_embed": lvl_pos_flatten,
            "ho_query_embed": query_embed,
            "obj_pred_names_sums": obj_pred_names_sums,
            "spatial_shapes": spatial_shapes,
            "level_start_index": level_start_index,
            "valid_ratios": valid_ratios,
            "spatial_shapes_host": shapes_host,        # extra key: lets phase B skip the shape syncs
        }

    def _late_fusion(self, fusion, hs_last, text, text_mask):
        """`--fusion_type MDETR_attn` (:703-733): the last decoder level's states and the label embeddings as ONE sequence
        through a pre-norm encoder stack (labels padded per `text_mask`); every layer's output is split back into
        (states [layers, bs, nq, C], labels [layers, n_text, bs, C])"""
        n_text = text.shape[0]
        seq = torch.cat((hs_last.permute(1, 0, 2), text), dim=0)
        pad = torch.cat((torch.zeros(hs_last.shape[:2], dtype=torch.bool, device=hs_last.device), text_mask.permute(1, 0)),
                        dim=1)
        out = fusion(seq, src_key_padding_mask=pad)                      # [layers, nq + n_text, bs, C]
        nq = out.shape[1] - n_text
        return out[:, :nq].permute(0, 2, 1, 3), out[:, nq:]

    def _decode(self, mask_flatten, query_embed, text_memory, img_memory, spatial_shapes, level_start_index,
                valid_ratios, spatial_shapes_host, text_attention_mask=None, obj_pred_names_sums=None):
        bs = img_memory.shape[0]
        c = self.d_model
        nq = query_embed.shape[0]
        reference_points = query_embed[..., 2 * c:].sigmoid()
        ref_sub, ref_obj = reference_points[:nq // 2], reference_points[nq // 2:]
        tgt = query_embed[..., :c].unsqueeze(0).expand(bs, -1, -1)
        verb_tgt = query_embed[..., c:2 * c].unsqueeze(0).expand(bs, -1, -1)
        init_reference_out = (ref_sub, ref_obj)
        hs_ho, inter_refs = self.ho_decoder(tgt, init_reference_out, img_memory, spatial_shapes, level_start_index,
                                            valid_ratios, query_pos=None, src_padding_mask=mask_flatten,
                                            spatial_shapes_host=spatial_shapes_host)
        if self.verb_query_tgt_type == "vanilla":
            merge_verb_tgt = verb_tgt[:, :nq // 2] + verb_tgt[:, nq // 2:]
        elif self.verb_query_tgt_type == "MBF":
            merge_verb_tgt = self.verb_tgt_generator(hs_ho[-1][:, :nq // 2], hs_ho[-1][:, nq // 2:])
        elif self.verb_query_tgt_type == "vanilla_MBF":
            merge_verb_tgt = self.verb_tgt_generator(hs_ho[-1][:, :nq // 2], hs_ho[-1][:, nq // 2:]) \
                + verb_tgt[:, :nq // 2] + verb_tgt[:, nq // 2:]
        else:
            raise ValueError(self.verb_query_tgt_type)
        hs_verb, _ = self.verb_decoder(merge_verb_tgt, inter_refs[-1], img_memory, spatial_shapes,
                                       level_start_index, valid_ratios, query_pos=None,
                                       src_padding_mask=mask_flatten, spatial_shapes_host=spatial_shapes_host)
        if self.fusion_type == "MDETR_attn":
            n_obj, n_pred = int(obj_pred_names_sums[:, 0].max()), int(obj_pred_names_sums[:, 1].max())
            assert n_obj + n_pred == text_memory.shape[0] == text_attention_mask.shape[0]
            hs_ho_dec, obj_text_dec = self._late_fusion(self.obj_fusion, hs_ho[-1], text_memory[:n_obj],
                                                        text_attention_mask[:n_obj])
            hs_verb_dec, pred_text_dec = self._late_fusion(self.verb_fusion, hs_verb[-1], text_memory[n_obj:],
                                                           text_attention_mask[n_obj:])
            text_dec = torch.cat((obj_text_dec, pred_text_dec), dim=1)
            return hs_ho_dec, hs_verb_dec, text_dec, init_reference_out, inter_refs, hs_ho, hs_verb, None, None
        hs_layer = hs_ho.shape[0]
        if text_memory.dim() == 4 and text_memory.shape[0] == hs_layer:
            text_dec = text_memory
        else:
            text_dec = text_memory.unsqueeze(0).repeat(hs_layer, 1, 1, 1)
        return hs_ho, hs_verb, text_dec, init_reference_out, inter_refs, hs_ho, hs_verb, None, None


def build_parseda_transformer(args):
    """The RLIP_ParSeDA_v2 branch of build_transformer (models/transformer.py:1344-1362)."""
    return RLIP_ParSeDABDeformableTransformer_v2(
        d_model=args.hidden_dim, nhead=args.nheads, num_encoder_layers=args.enc_layers,
        num_decoder_layers=args.dec_layers, dim_feedforward=args.dim_feedforward, dropout=args.dropout,
        activation="relu", return_intermediate_dec=True, num_feature_levels=args.num_feature_levels,
        dec_n_points=args.dec_n_points, enc_n_points=args.enc_n_points, two_stage=args.two_stage,
        two_stage_num_proposals=args.num_queries, use_dab=True, args=args)
    # NB: like the reference, `--text_encoder_type` / `--freeze_text_encoder` are NOT forwarded here
    # (transformer.py:1346-1362), so ParSeDA always trains a roberta-base text encoder.
