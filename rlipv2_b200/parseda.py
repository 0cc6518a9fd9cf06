"""RLIP_ParSeDA - the RLIPv2-ParSeDA model head layer (flag `--RLIP_ParSeDA_v2`).

Mirror of /root/reference/models/hoi.py:1871-2256 (+ MLP :3589-3601): input projections, the
two-phase `encode_and_save` protocol driven by engine.py:99-100, iterative box heads, label-text
projection and the similarity logits, auxiliary outputs, optional pseudo-verb similarity targets.
Parameter names and aliases are identical: the box heads are registered as `sub_bbox_embed.{0..5}`
/ `obj_bbox_embed.{0..5}` and shared with `transformer.ho_decoder.*_bbox_embed.{0..2}` and
`transformer.verb_decoder.*_bbox_embed.{0..2}` (hoi.py:1980-1990).
"""
import copy
import os
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import dense, grad_ready
from .nested import NestedTensor, inverse_sigmoid, nested_tensor_from_tensor_list
from .parseda_transformer import MLP


_TEXT_STREAM = os.environ.get("RLIPV2_TEXT_STREAM", "1") != "0"      # A/B switch for measurements
class FlatLevels(list):
    """per-level [N, C, H, W] views of one token buffer `flat` [N, sum HW, C] (the transformer takes `flat` as is)"""
    flat = None


_STACKED_HEADS = os.environ.get("RLIPV2_STACKED_HEADS", "1") != "0"  # label-text heads of all decoder levels in one pass


def _get_clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


class RLIP_ParSeDA(nn.Module):
    def __init__(self, backbone, transformer, num_queries, num_feature_levels, aux_loss=True, with_box_refine=True,
                 two_stage=False, use_dab=True, num_patterns=0, random_refpoints_xy=False, subject_class=False,
                 pseudo_verb=False, args=None):
        super().__init__()
        if two_stage or not use_dab or num_patterns != 0 or not with_box_refine:
            raise NotImplementedError("ParSeDA scripts: use_dab, with_box_refine, no two_stage, no patterns")
        self.num_queries = num_queries
        self.transformer = transformer
        hidden_dim = transformer.d_model
        self.num_feature_levels = num_feature_levels
        self.use_dab = use_dab
        self.num_patterns = num_patterns
        self.random_refpoints_xy = random_refpoints_xy

        sub_bbox_embed = MLP(hidden_dim, hidden_dim, 4, 3)
        obj_bbox_embed = MLP(hidden_dim, hidden_dim, 4, 3)
        self.projection_text = nn.Linear(hidden_dim, hidden_dim)
        prior_prob = 0.01
        self.bias_c = -math.log((1 - prior_prob) / prior_prob)
        self.bias_obj_a = nn.Parameter(torch.zeros((256,), dtype=torch.float32), requires_grad=True)
        self.bias_pred_a = nn.Parameter(torch.zeros((256,), dtype=torch.float32), requires_grad=True)

        self.tgt_embed = nn.Embedding(num_queries, hidden_dim)
        self.verb_tgt_embed = nn.Embedding(num_queries, hidden_dim)
        self.refpoint_embed = nn.Embedding(num_queries, 4)
        if random_refpoints_xy:
            self.refpoint_embed.weight.data[:, :2].uniform_(0, 1)
            self.refpoint_embed.weight.data[:, :2] = inverse_sigmoid(self.refpoint_embed.weight.data[:, :2])

        # 1x1 conv + GroupNorm per backbone level; extra levels: 3x3 stride-2 conv on the last one
        proj = []
        in_channels = None
        for i in range(len(backbone.strides)):
            in_channels = backbone.num_channels[i]
            proj.append(nn.Sequential(nn.Conv2d(in_channels, hidden_dim, kernel_size=1), nn.GroupNorm(32, hidden_dim)))
        for _ in range(num_feature_levels - len(backbone.strides)):
            proj.append(nn.Sequential(nn.Conv2d(in_channels, hidden_dim, kernel_size=3, stride=2, padding=1),
                                      nn.GroupNorm(32, hidden_dim)))
            in_channels = hidden_dim
        self.input_proj = nn.ModuleList(proj)
        self.backbone = backbone
        self.aux_loss = aux_loss
        self.with_box_refine = with_box_refine
        self.two_stage = two_stage

        for head in (sub_bbox_embed, obj_bbox_embed):
            nn.init.constant_(head.layers[-1].weight.data, 0)
            nn.init.constant_(head.layers[-1].bias.data, 0)
        for p in self.input_proj:
            nn.init.xavier_uniform_(p[0].weight, gain=1)
            nn.init.constant_(p[0].bias, 0)

        num_pred = transformer.ho_decoder.num_layers
        self.sub_bbox_embed = _get_clones(sub_bbox_embed, num_pred * 2)
        nn.init.constant_(self.sub_bbox_embed[0].layers[-1].bias.data[2:], -2.0)
        self.transformer.ho_decoder.sub_bbox_embed = self.sub_bbox_embed[:num_pred]
        self.transformer.verb_decoder.sub_bbox_embed = self.sub_bbox_embed[num_pred:]
        self.obj_bbox_embed = _get_clones(obj_bbox_embed, num_pred * 2)
        nn.init.constant_(self.obj_bbox_embed[0].layers[-1].bias.data[2:], -2.0)
        self.transformer.ho_decoder.obj_bbox_embed = self.obj_bbox_embed[:num_pred]
        self.transformer.verb_decoder.obj_bbox_embed = self.obj_bbox_embed[num_pred:]

        self.subject_class = subject_class
        self.pseudo_verb = pseudo_verb
        self.pseudo_verb_mode = "online"

    # ---- phase A: backbone + input projections + encoder -------------------------------------------
    def _encode(self, samples, text):
        if _TEXT_STREAM and samples.tensors.is_cuda and self.transformer._is_label_text(text):
            text = self.transformer.encode_text_async(text, samples.tensors.device)   # overlaps the backbone
        features, pos = self.backbone(samples)
        for i, feat in enumerate(features):              # no-ops unless the data-parallel step installed a callback
            feat.tensors = grad_ready.mark(feat.tensors, f"image{i}")
        srcs, masks = [], []
        norms = [p[1] for p in self.input_proj]
        if (self.num_feature_levels <= len(features) + 1
                and dense.group_norm_tokens_supported([f.tensors for f in features], [p[0] for p in self.input_proj],
                                                      norms)):
            # GPU path: the projections' convolutions (cuDNN, NHWC outputs), then ONE fused op that applies every
            # level's GroupNorm on the token-major data and writes the encoder's [N, sum HW, 256] token buffer directly
            # (no NCHW round trip, no flatten / transpose / cat copies); `srcs` are views of that buffer
            convs = [self.input_proj[l][0](feat.tensors) for l, feat in enumerate(features)]
            masks = [feat.mask for feat in features]
            for l in range(len(features), self.num_feature_levels):
                convs.append(self.input_proj[l][0](features[-1].tensors))
                mask = F.interpolate(samples.mask[None].float(), size=convs[-1].shape[-2:]).to(torch.bool)[0]
                pos.append(self.backbone[1](NestedTensor(convs[-1], mask)).to(convs[-1].dtype))
                masks.append(mask)
            flat = dense.group_norm_tokens(convs, norms[:len(convs)])
            srcs, start = FlatLevels(), 0
            for c in convs:
                n_, ch, h, w = c.shape
                srcs.append(flat[:, start:start + h * w].view(n_, h, w, ch).permute(0, 3, 1, 2))
                start += h * w
            srcs.flat = flat
        else:
            for l, feat in enumerate(features):
                src, mask = feat.decompose()
                srcs.append(self.input_proj[l](src))
                masks.append(mask)
            for l in range(len(srcs), self.num_feature_levels):
                src = self.input_proj[l](features[-1].tensors if l == len(features) else srcs[-1])
                mask = F.interpolate(samples.mask[None].float(), size=src.shape[-2:]).to(torch.bool)[0]
                pos.append(self.backbone[1](NestedTensor(src, mask)).to(src.dtype))
                srcs.append(src)
                masks.append(mask)
        query_embeds = torch.cat((self.tgt_embed.weight, self.verb_tgt_embed.weight, self.refpoint_embed.weight), dim=1)
        return self.transformer(srcs=srcs, masks=masks, pos_embeds=pos, query_embed=query_embeds, text=text,
                                encode_and_save=True)

    def forward(self, samples, encode_and_save=True, memory_cache=None, **kwargs):
        if not isinstance(samples, NestedTensor):
            if hasattr(samples, "tensors") and hasattr(samples, "mask"):     # the reference's own NestedTensor
                samples = NestedTensor(samples.tensors, samples.mask)
            else:
                samples = nested_tensor_from_tensor_list(samples)
        if encode_and_save:
            return self._encode(samples, kwargs["text"])
        return self._decode(memory_cache, **kwargs)

    # ---- phase B: decoders + heads ---------------------------------------------------------------------
    def _decode(self, memory_cache, **kwargs):
        hs_ho, hs_verb, text_dec, init_reference, inter_references, _, _, _, _ = self.transformer(
            masks=memory_cache["masks"], query_embed=memory_cache["ho_query_embed"], encode_and_save=False,
            text_memory=memory_cache["text_memory_resized"], img_memory=memory_cache["img_memory"],
            text_attention_mask=memory_cache["text_attention_mask"],
            obj_pred_names_sums=memory_cache["obj_pred_names_sums"], spatial_shapes=memory_cache["spatial_shapes"],
            level_start_index=memory_cache["level_start_index"], valid_ratios=memory_cache["valid_ratios"],
            spatial_shapes_host=memory_cache.get("spatial_shapes_host"))
        half = self.num_queries // 2
        hs_h, hs_o = hs_ho[:, :, :half], hs_ho[:, :, half:]
        sums = memory_cache["obj_pred_names_sums"]                # CPU tensor [n_tuples, 2]
        max_obj = int(sums[:, 0].max())
        max_pred = int(sums[:, 1].max())

        sub_cls, obj_cls, verb_cls, sub_boxes, obj_boxes = [], [], [], [], []
        # The box heads are the modules the pair decoder already applied for its anchor refinement (shared,
        # hoi.py:1980-1990) to the same `hs` and anchors: take its un-detached results instead of evaluating
        # two 3-layer MLPs + inverse_sigmoid per level a second time.
        refined = getattr(self.transformer.ho_decoder, "refined_boxes", None)
        self.transformer.ho_decoder.refined_boxes = None
        if not (self.with_box_refine and refined is not None and len(refined) == hs_h.shape[0]
                and getattr(self.transformer, "fusion_type", "GLIP_attn") != "MDETR_attn"):
            refined = None                 # (late fusion returns re-encoded states: the heads must see those)
        n_lvl = hs_h.shape[0]
        for lvl in range(n_lvl):
            if refined is not None:
                sub_boxes.append(refined[lvl][0])
                obj_boxes.append(refined[lvl][1])
            else:
                sub_ref, obj_ref = init_reference if lvl == 0 else inter_references[lvl - 1]
                sub_boxes.append((self.sub_bbox_embed[lvl](hs_h[lvl]) + inverse_sigmoid(sub_ref)).sigmoid())
                obj_boxes.append((self.obj_bbox_embed[lvl](hs_o[lvl]) + inverse_sigmoid(obj_ref)).sigmoid())
        if _STACKED_HEADS and torch.is_tensor(text_dec) and text_dec.dim() == 4 and text_dec.shape[0] == n_lvl:
            # the label-text heads of all decoder levels in one pass (hoi.py:2145-2163 evaluates them level by level):
            # one normalisation, one projection GEMM, one batched contraction per head instead of n_lvl of each
            text_memory = F.normalize(text_dec.transpose(1, 2), p=2, dim=-1)             # [L, bs, Tl, C]
            proj_text = dense.linear(text_memory / 2.0, self.projection_text.weight, self.projection_text.bias)
            assert max_obj + max_pred == proj_text.shape[2]
            obj_text = proj_text[:, :, :max_obj].transpose(2, 3)                          # [L, bs, 256, n_obj]
            pred_text = proj_text[:, :, max_obj:max_obj + max_pred].transpose(2, 3)
            obj_cls = list((torch.matmul(hs_o + self.bias_obj_a, obj_text) + self.bias_c).unbind(0))
            verb_cls = list((torch.matmul(hs_verb + self.bias_pred_a, pred_text) + self.bias_c).unbind(0))
            if self.subject_class:
                sub_cls = list((torch.matmul(hs_h + self.bias_obj_a, obj_text) + self.bias_c).unbind(0))
        else:
            for lvl in range(n_lvl):
                text_memory = F.normalize(text_dec[lvl].transpose(0, 1), p=2, dim=-1)
                proj_text = dense.linear(text_memory / 2.0, self.projection_text.weight, self.projection_text.bias)
                assert max_obj + max_pred == proj_text.shape[1]
                obj_text = proj_text[:, :max_obj].transpose(1, 2)                      # [bs, 256, n_obj]
                pred_text = proj_text[:, max_obj:max_obj + max_pred].transpose(1, 2)
                obj_cls.append(torch.matmul(hs_o[lvl] + self.bias_obj_a, obj_text) + self.bias_c)
                verb_cls.append(torch.matmul(hs_verb[lvl] + self.bias_pred_a, pred_text) + self.bias_c)
                if self.subject_class:
                    sub_cls.append(torch.matmul(hs_h[lvl] + self.bias_obj_a, obj_text) + self.bias_c)

        out = {"pred_obj_logits": obj_cls[-1], "pred_verb_logits": verb_cls[-1],
               "pred_sub_boxes": sub_boxes[-1], "pred_obj_boxes": obj_boxes[-1]}
        if self.subject_class:
            out = {"pred_sub_logits": sub_cls[-1], **out}
        if self.aux_loss:
            aux = []
            for i in range(len(obj_cls) - 1):
                d = {"pred_obj_logits": obj_cls[i], "pred_verb_logits": verb_cls[i],
                     "pred_sub_boxes": sub_boxes[i], "pred_obj_boxes": obj_boxes[i]}
                if self.subject_class:
                    d = {"pred_sub_logits": sub_cls[i], **d}
                aux.append(d)
            out["aux_outputs"] = aux

        if self.pseudo_verb:
            # pseudo relation labels from the distances between the pre-fusion verb text embeddings
            # (hoi.py:2197-2239, "online" mode)
            text_bf = memory_cache["text_memory_bf_resize"]
            verb_text = text_bf[:, 0][max_obj:max_obj + max_pred]
            # F.pairwise_distance(x1, x2) = ||x1 - x2 + 1e-6||_2 over all ordered pairs
            dist = (verb_text[:, None, :] - verb_text[None, :, :] + 1e-6).norm(p=2, dim=-1)
            verb_sim = dist.max(-1)[0].unsqueeze(-1) - dist
            target_verbs = torch.cat([t["verb_labels"] for t in kwargs["targets"]])
            target_verb_sim = (target_verbs.unsqueeze(-1) * verb_sim).sum(dim=1)
            if target_verbs.shape[0] > 0:
                target_verb_sim = target_verb_sim / target_verb_sim.max(-1)[0].unsqueeze(-1)
            target_verb_sim[target_verbs.bool()] = 0
            target_verb_sim = target_verb_sim * (target_verb_sim > 0.3)
            out["target_verb_sim"] = target_verb_sim
            if self.aux_loss:
                for aux in out["aux_outputs"]:
                    aux["target_verb_sim"] = target_verb_sim
        return out
