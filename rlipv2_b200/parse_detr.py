"""RLIP-ParSe: the plain-DETR member of the family (BASELINE config 1, SURVEY.md section 8d "plumbing, CPU").

Mirror of the reference's
  RLIP_ParSe                    /root/reference/models/hoi.py:2259-2512
  ParSeTransformer              models/ParSetransformer.py:963-1204
  CrossModelTransformerEncoder  models/ParSetransformer.py:1503-1533
  TransformerDecoder            models/ParSetransformer.py:1637-1687
  TransformerEncoderLayer       models/ParSetransformer.py:1690-1751
  TransformerDecoderLayer       models/ParSetransformer.py:1754-1906
with identical sub-module / parameter names (the 671 `state_dict` keys of the reference model load unchanged) and the
same two-phase `encode_and_save` protocol as RLIP_ParSeDA.

What it computes.  Phase A: ResNet C5 -> 1x1 projection -> image tokens [HW, N, 256]; the label strings go through the
text tower (pooled vectors, resized 768 -> 256) and are APPENDED to the image tokens; six post-norm encoder layers run
plain multi-head self-attention over the joint sequence (position embedding added to q / k of the image tokens, zeros
for the labels) and every layer's output is kept.  Phase B: a pair decoder (2 x num_queries queries: subjects then
objects) and a relation decoder whose query positions are the sums of each pair's final features; three DETR decoder
layers each (self-attention, cross-attention to the last encoder layer's joint memory, FFN).  Heads as in ParSeDA: box
MLPs, and class logits = <feature + bias_a, projected normalised label feature / 2> + bias_c with the label features of
the matching encoder layer.

This configuration is the family's CPU-runnable plumbing case: everything is torch (nn.MultiheadAttention, Linear,
LayerNorm); none of the sm_100a kernels is on this path.
"""
import copy
import math

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.utils.rnn import pad_sequence

from .alif import FeatureResizer
from .nested import NestedTensor, nested_tensor_from_tensor_list
from .parseda_transformer import MLP
from .text_encoder import build_text_encoder, pooled_text


def _clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


def _activation(name):
    try:
        return {"relu": F.relu, "gelu": F.gelu, "glu": F.glu}[name]
    except KeyError:
        raise RuntimeError(f"activation should be relu/gelu, not {name}.")


def _add(x, pos):
    return x if pos is None else x + pos


class TransformerEncoderLayer(nn.Module):
    """self-attention + FFN, post-norm (default) or pre-norm (ParSetransformer.py:1690-1751)"""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.activation = _activation(activation)
        self.normalize_before = normalize_before

    def _ffn(self, x):
        return self.linear2(self.dropout(self.activation(self.linear1(x))))

    def forward(self, src, src_mask=None, src_key_padding_mask=None, pos=None):
        if self.normalize_before:
            y = self.norm1(src)
            qk = _add(y, pos)
            src = src + self.dropout1(self.self_attn(qk, qk, value=y, attn_mask=src_mask,
                                                     key_padding_mask=src_key_padding_mask)[0])
            return src + self.dropout2(self._ffn(self.norm2(src)))
        qk = _add(src, pos)
        src = self.norm1(src + self.dropout1(self.self_attn(qk, qk, value=src, attn_mask=src_mask,
                                                            key_padding_mask=src_key_padding_mask)[0]))
        return self.norm2(src + self.dropout2(self._ffn(src)))


class TransformerDecoderLayer(nn.Module):
    """self-attention among the queries, cross-attention to the joint image + label memory, FFN
    (ParSetransformer.py:1754-1906; the separate text cross-attention of MDETR is commented out there, so
    `text_memory` is accepted and unused; norm2 / dropout2 do not exist)"""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.cross_attn_image = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.norm4 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)
        self.dropout4 = nn.Dropout(dropout)
        self.activation = _activation(activation)
        self.normalize_before = normalize_before

    def _ffn(self, x):
        return self.linear2(self.dropout(self.activation(self.linear1(x))))

    def forward(self, tgt, memory, text_memory=None, tgt_mask=None, memory_mask=None,
                text_memory_key_padding_mask=None, tgt_key_padding_mask=None, memory_key_padding_mask=None,
                pos=None, query_pos=None):
        key = _add(memory, pos)
        if self.normalize_before:
            y = self.norm1(tgt)
            qk = _add(y, query_pos)
            tgt = tgt + self.dropout1(self.self_attn(qk, qk, value=y, attn_mask=tgt_mask,
                                                     key_padding_mask=tgt_key_padding_mask)[0])
            y = self.norm3(tgt)
            tgt = tgt + self.dropout3(self.cross_attn_image(query=_add(y, query_pos), key=key, value=memory,
                                                            attn_mask=memory_mask,
                                                            key_padding_mask=memory_key_padding_mask)[0])
            return tgt + self.dropout4(self._ffn(self.norm4(tgt)))
        qk = _add(tgt, query_pos)
        tgt = self.norm1(tgt + self.dropout1(self.self_attn(qk, qk, value=tgt, attn_mask=tgt_mask,
                                                            key_padding_mask=tgt_key_padding_mask)[0]))
        tgt = self.norm3(tgt + self.dropout3(self.cross_attn_image(query=_add(tgt, query_pos), key=key, value=memory,
                                                                   attn_mask=memory_mask,
                                                                   key_padding_mask=memory_key_padding_mask)[0]))
        return self.norm4(tgt + self.dropout4(self._ffn(tgt)))


class CrossModelTransformerEncoder(nn.Module):
    """the encoder stack over the joint (image ; label) sequence; returns every layer's output when
    `return_intermediate` (ParSetransformer.py:1503-1533)"""

    def __init__(self, encoder_layer, num_layers, norm=None, return_intermediate=False):
        super().__init__()
        self.layers = _clones(encoder_layer, num_layers)
        self.num_layers = num_layers
        self.norm = norm
        self.return_intermediate = return_intermediate

    def forward(self, src, mask=None, src_key_padding_mask=None, pos=None):
        out, kept = src, []
        for layer in self.layers:
            out = layer(out, src_mask=mask, src_key_padding_mask=src_key_padding_mask, pos=pos)
            kept.append(out if self.norm is None else self.norm(out))
        return torch.stack(kept) if self.return_intermediate else kept[-1]


class TransformerDecoder(nn.Module):
    """DETR decoder stack; with `return_intermediate` the (normed) output of every layer (:1637-1687)"""

    def __init__(self, decoder_layer, num_layers, norm=None, return_intermediate=False):
        super().__init__()
        self.layers = _clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.norm = norm
        self.return_intermediate = return_intermediate

    def forward(self, tgt, memory, text_memory, tgt_mask=None, memory_mask=None, text_memory_key_padding_mask=None,
                tgt_key_padding_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None):
        out, kept = tgt, []
        for layer in self.layers:
            out = layer(out, memory, text_memory=text_memory, tgt_mask=tgt_mask, memory_mask=memory_mask,
                        text_memory_key_padding_mask=text_memory_key_padding_mask,
                        tgt_key_padding_mask=tgt_key_padding_mask, memory_key_padding_mask=memory_key_padding_mask,
                        pos=pos, query_pos=query_pos)
            if self.return_intermediate:
                kept.append(self.norm(out))
        if self.norm is not None:
            out = self.norm(out)           # same value as kept[-1]: the last layer's output, normed
        return torch.stack(kept) if self.return_intermediate else out


class ParSeTransformer(nn.Module):
    """joint image-label encoder + pair decoder + relation decoder (ParSetransformer.py:963-1204)"""

    def __init__(self, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=2048,
                 dropout=0.1, activation="relu", normalize_before=False, return_intermediate_dec=False,
                 pass_pos_and_query=True, text_encoder_type="roberta-base", freeze_text_encoder=False,
                 synthetic_text_encoder=None):
        super().__init__()
        self.pass_pos_and_query = pass_pos_and_query
        enc_layer = TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        self.encoder = CrossModelTransformerEncoder(enc_layer, num_encoder_layers,
                                                    nn.LayerNorm(d_model) if normalize_before else None,
                                                    return_intermediate=True)
        dec_layer = TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        self.ho_decoder = TransformerDecoder(dec_layer, num_decoder_layers, nn.LayerNorm(d_model),
                                             return_intermediate=return_intermediate_dec)
        dec_layer = TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        self.verb_decoder = TransformerDecoder(dec_layer, num_decoder_layers, nn.LayerNorm(d_model),
                                               return_intermediate=return_intermediate_dec)
        for p in self.parameters():            # :1030-1033, before the text tower and the resizer are attached
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        if "roberta" not in text_encoder_type:
            raise NotImplementedError("the RLIP scripts use roberta-base")
        self.tokenizer, self.text_encoder = build_text_encoder(text_encoder_type, synthetic=synthetic_text_encoder)
        if freeze_text_encoder:
            for p in self.text_encoder.parameters():
                p.requires_grad_(False)
        self.expander_dropout = 0.1
        self.resizer = FeatureResizer(input_feat_size=self.text_encoder.config.hidden_size, output_feat_size=d_model,
                                      dropout=self.expander_dropout)
        self.d_model = d_model
        self.nhead = nhead

    # ---- label strings -> pooled text vectors, padded per tuple (:1086-1135) ----------------------------------------
    def encode_text(self, text, device):
        sums, flat = [], []
        for objs, verbs in text:
            sums.append((len(objs), len(verbs)))
            flat += list(objs) + list(verbs)
        tok = self.tokenizer.batch_encode_plus(flat, padding="longest", return_tensors="pt").to(device)
        pooled = pooled_text(self.text_encoder, tok["input_ids"], tok["attention_mask"])
        obj_rows, verb_rows, at = [], [], 0
        for n_obj, n_verb in sums:
            obj_rows.append(pooled[at:at + n_obj])
            verb_rows.append(pooled[at + n_obj:at + n_obj + n_verb])
            at += n_obj + n_verb
        text_memory = torch.cat([pad_sequence(obj_rows), pad_sequence(verb_rows)], dim=0)     # [n_text, n_tuples, 768]
        text_attention_mask = ~(text_memory.sum(dim=-1) > 0)
        return text_memory, text_attention_mask, torch.tensor(sums)

    def forward(self, src=None, mask=None, query_embed=None, pos_embed=None, text=None, encode_and_save=True,
                text_memory=None, img_memory=None, text_attention_mask=None):
        if not encode_and_save:
            return self._decode(mask, query_embed, pos_embed, text_memory, img_memory, text_attention_mask)
        bs = src.shape[0]
        src = src.flatten(2).permute(2, 0, 1)                                                  # [HW, N, C]
        pos_embed = pos_embed.flatten(2).permute(2, 0, 1)
        ho_query_embed = query_embed.unsqueeze(1).repeat(1, bs, 1)
        mask = mask.flatten(1)
        if not self.pass_pos_and_query:
            # the reference's branch (:1067) sets pos_embed / query_embed to None and then concatenates / zeros_like()s
            # them: it cannot run; every script keeps the default
            raise NotImplementedError("pass_pos_and_query=False does not run in the reference either")
        if isinstance(text, list) and isinstance(text[0], tuple):                              # training: label strings
            raw, text_attention_mask, sums = self.encode_text(text, src.device)
            text_memory_resized = self.resizer(raw)
            if text_memory_resized.shape[1] != bs:
                text_memory_resized = text_memory_resized.repeat(1, bs, 1)
                text_attention_mask = text_attention_mask.repeat(1, bs)
        else:                                                                                   # eval: pre-encoded
            text_attention_mask, text_memory_resized, sums = text
        n_text = len(text_memory_resized)
        src = torch.cat([src, text_memory_resized], dim=0)
        text_attention_mask = text_attention_mask.transpose(0, 1)
        mask = torch.cat([mask, text_attention_mask], dim=1)
        # zeros for the label tokens: adding them is a no-op
        pos_embed = torch.cat([pos_embed, torch.zeros_like(text_memory_resized)], dim=0)
        layers_out = self.encoder(src, src_key_padding_mask=mask, pos=pos_embed)               # [L, HW + n_text, N, C]
        return {
            "text_memory_resized": text_memory_resized,
            "text_memory": layers_out[:, -n_text:],
            "img_memory": layers_out[-1],
            "mask": mask,
            "text_attention_mask": text_attention_mask,
            "pos_embed": pos_embed,
            "ho_query_embed": ho_query_embed,
            "obj_pred_names_sums": sums,
        }

    def _decode(self, mask, query_embed, pos_embed, text_memory, img_memory, text_attention_mask):
        ho = self.ho_decoder(torch.zeros_like(query_embed), img_memory, text_memory, memory_key_padding_mask=mask,
                             text_memory_key_padding_mask=text_attention_mask, pos=pos_embed, query_pos=query_embed)
        ho = ho.transpose(1, 2)                                                                 # [L, N, 2 nq, C]
        pairs = ho.shape[2] // 2
        h_out, o_out = ho[:, :, :pairs], ho[:, :, pairs:]
        verb_query = (h_out[-1] + o_out[-1]).permute(1, 0, 2)
        verb = self.verb_decoder(torch.zeros_like(verb_query), img_memory, text_memory, memory_key_padding_mask=mask,
                                 text_memory_key_padding_mask=text_attention_mask, pos=pos_embed, query_pos=verb_query)
        return h_out, o_out, verb.transpose(1, 2)


class RLIP_ParSe(nn.Module):
    """models/hoi.py:2259-2512 restricted to the classification variant every RLIP script uses
    (`contrastive_align_loss=False`: cross-entropy / focal losses on label-text logits with the bias trick)."""

    def __init__(self, backbone, transformer, num_queries, contrastive_align_loss=False, contrastive_hdim=64,
                 aux_loss=False, subject_class=False, use_no_verb_token=False, pseudo_verb=False, args=None):
        super().__init__()
        if contrastive_align_loss:
            raise NotImplementedError("cross_modal_matching losses are not used by the RLIP scripts")
        self.num_queries = num_queries
        self.transformer = transformer
        hidden_dim = transformer.d_model
        self.query_embed = nn.Embedding(num_queries * 2, hidden_dim)
        self.sub_bbox_embed = MLP(hidden_dim, hidden_dim, 4, 3)
        self.obj_bbox_embed = MLP(hidden_dim, hidden_dim, 4, 3)
        channels = backbone.num_channels
        self.input_proj = nn.Conv2d(channels[-1] if isinstance(channels, (list, tuple)) else channels, hidden_dim,
                                    kernel_size=1)
        self.backbone = backbone
        self.aux_loss = aux_loss
        self.contrastive_align_loss = False
        self.subject_class = subject_class
        self.use_no_verb_token = use_no_verb_token
        self.pseudo_verb = pseudo_verb
        self.projection_text = nn.Linear(hidden_dim, hidden_dim)
        prior_prob = 0.01
        self.bias_c = -math.log((1 - prior_prob) / prior_prob)
        self.bias_obj_a = nn.Parameter(torch.zeros((256,), dtype=torch.float32), requires_grad=True)
        self.bias_pred_a = nn.Parameter(torch.zeros((256,), dtype=torch.float32), requires_grad=True)
        for head in (self.sub_bbox_embed, self.obj_bbox_embed):
            nn.init.constant_(head.layers[-1].weight.data, 0)
            nn.init.constant_(head.layers[-1].bias.data, 0)
        nn.init.xavier_uniform_(self.input_proj.weight, gain=1)
        nn.init.constant_(self.input_proj.bias, 0)
        self.verb_tagger = getattr(args, "verb_tagger", False)

    def forward(self, samples, encode_and_save=True, memory_cache=None, **kwargs):
        if not isinstance(samples, NestedTensor):
            if hasattr(samples, "tensors") and hasattr(samples, "mask"):
                samples = NestedTensor(samples.tensors, samples.mask)
            else:
                samples = nested_tensor_from_tensor_list(samples)
        if encode_and_save:
            assert memory_cache is None
            features, pos = self.backbone(samples)
            src, mask = features[-1].decompose()
            assert mask is not None
            return self.transformer(src=self.input_proj(src), mask=mask, query_embed=self.query_embed.weight,
                                    pos_embed=pos[-1], text=kwargs["text"], encode_and_save=True)
        assert memory_cache is not None
        h_out, o_out, v_out = self.transformer(
            mask=memory_cache["mask"], query_embed=memory_cache["ho_query_embed"], pos_embed=memory_cache["pos_embed"],
            encode_and_save=False, text_memory=memory_cache["text_memory"][-1], img_memory=memory_cache["img_memory"],
            text_attention_mask=memory_cache["text_attention_mask"])
        sums = memory_cache["obj_pred_names_sums"]
        n_obj, n_verb = int(sums[:, 0].max()), int(sums[:, 1].max())
        sub_boxes = self.sub_bbox_embed(h_out).sigmoid()
        obj_boxes = self.obj_bbox_embed(o_out).sigmoid()
        n_dec = o_out.shape[0]
        obj_cls, verb_cls, sub_cls = [], [], []
        for i in range(-n_dec, 0):                       # the LAST n_dec encoder layers' label features (:2385-2400)
            labels = F.normalize(memory_cache["text_memory"][i].transpose(0, 1), p=2, dim=-1)
            proj = self.projection_text(labels / 2.0)
            assert n_obj + n_verb == proj.shape[1]
            obj_text, verb_text = proj[:, :n_obj].transpose(1, 2), proj[:, n_obj:n_obj + n_verb].transpose(1, 2)
            obj_cls.append(torch.matmul(o_out[i] + self.bias_obj_a, obj_text) + self.bias_c)
            verb_cls.append(torch.matmul(v_out[i] + self.bias_pred_a, verb_text) + self.bias_c)
            if self.subject_class:
                sub_cls.append(torch.matmul(h_out[i] + self.bias_obj_a, obj_text) + self.bias_c)
        out = {"pred_obj_logits": obj_cls[-1], "pred_verb_logits": verb_cls[-1], "pred_sub_boxes": sub_boxes[-1],
               "pred_obj_boxes": obj_boxes[-1]}
        if self.subject_class:
            out["pred_sub_logits"] = sub_cls[-1]
        if self.aux_loss:
            aux = []
            for i in range(n_dec - 1):
                d = {"pred_obj_logits": obj_cls[i], "pred_verb_logits": verb_cls[i], "pred_sub_boxes": sub_boxes[i],
                     "pred_obj_boxes": obj_boxes[i]}
                if self.subject_class:
                    d = {"pred_sub_logits": sub_cls[i], **d}
                aux.append(d)
            out["aux_outputs"] = aux
        if self.pseudo_verb:
            raise NotImplementedError("--pseudo_verb with RLIP_ParSe needs `text_memory_bf_resized`, which the reference's "
                                      "ParSeTransformer never stores (hoi.py:2446 would raise KeyError there too)")
        return out


def build_parse_transformer(args):
    """models/transformer.py:1188-1202"""
    return ParSeTransformer(
        d_model=args.hidden_dim, dropout=args.dropout, nhead=args.nheads, dim_feedforward=args.dim_feedforward,
        num_encoder_layers=args.enc_layers, num_decoder_layers=args.dec_layers, normalize_before=args.pre_norm,
        return_intermediate_dec=True, pass_pos_and_query=getattr(args, "pass_pos_and_query", True),
        text_encoder_type=args.text_encoder_type, freeze_text_encoder=args.freeze_text_encoder,
        synthetic_text_encoder=getattr(args, "synthetic_text_encoder", None))
