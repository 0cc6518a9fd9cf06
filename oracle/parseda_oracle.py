"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

CPU oracle of the RLIPv2-ParSeDA train step: a *functional* PyTorch restatement (plain functions over
a name -> tensor dict, no nn.Module of this repo) of the reference's per-step hot path, each function
citing the reference lines it follows.  All paths relative to /root/reference.

  alif_block            models/fuse_helper.py:365-466, 684-721 ("VXAc" gate)
  roberta_layer         models/modeling_roberta.py:149-241, 252-256, 318-336, 356-408
  msda_module           models/ops/modules/ms_deform_attn.py:82-119 (+ oracle/msda_torch_oracle.py)
  encoder_layer         models/dab_deformable/deformable_transformer.py:1283-1300
  encoder               models/deformable_transformer.py:803-884
  decoder_layer / dab_decoder   dab_deformable/deformable_transformer.py:1383-1401, 1432-1552, 1777-1802
  mbf                   dab_deformable/deformable_transformer.py:1063-1068
  forward_step          dab_deformable/deformable_transformer.py:456-744 and models/hoi.py:2034-2194
  matcher / criterion   models/matcher.py:108-202, models/hoi.py:3696-3828, 3909-4028, 4162-4193, 4481-4495, 4654-4766

Third-party arithmetic stays third-party, exactly as in the reference: torchvision ResNet-50 (with
the frozen batch-norm of models/DDETR_backbone.py:31-68), HF `RobertaModel` for the label strings,
`scipy.optimize.linear_sum_assignment`.

Pinned by tests/test_parseda_oracle.py against tests/golden/parseda_step.npz, which
oracle/gen_golden_model.py produced by running the reference's own modules (same name-keyed weights).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference may import it.
"""
import json
import math
import os
import time

import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment
from torch.nn.utils.rnn import pad_sequence

from oracle.msda_torch_oracle import msda_core

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def lin(W, p, x):
    return F.linear(x, W[p + ".weight"], W[p + ".bias"])


def ln(W, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), W[p + ".weight"], W[p + ".bias"], eps)


def mlp(W, p, x, n):
    for i in range(n):
        x = lin(W, f"{p}.layers.{i}", x)
        if i < n - 1:
            x = F.relu(x)
    return x


def inverse_sigmoid(x, eps=1e-5):                       # util/misc.py:460-464
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


# ---- ALIF ----------------------------------------------------------------------------------------------
def alif_block(W, p, v, l, pos, drop=0.0):
    """fuse_helper.py:684-721 + 365-466.  The bool masks are a no-op there (masked_fill on bool)."""
    a = p + ".b_attn"
    v = ln(W, a + ".layer_norm_v", v)
    l = ln(W, a + ".layer_norm_l", l)
    b, tv, _ = v.shape
    tl = l.shape[1]
    H, Dh = 8, 256
    q = lin(W, a + ".attn.v_proj", v + pos) * Dh ** -0.5
    k = lin(W, a + ".attn.l_proj", l)
    vv = lin(W, a + ".attn.values_v_proj", v)
    vl = lin(W, a + ".attn.values_l_proj", l)
    sh = lambda t, n: t.view(b, n, H, Dh).transpose(1, 2)
    q, k, vv, vl = sh(q, tv), sh(k, tl), sh(vv, tv), sh(vl, tl)
    s = q @ k.transpose(-1, -2)
    st = s.transpose(-1, -2)
    p_l = F.dropout((st - st.max(-1, keepdim=True)[0]).softmax(-1), drop, drop > 0)
    p_v = F.dropout(s.softmax(-1), drop, drop > 0)
    ov = (p_v @ vl).transpose(1, 2).reshape(b, tv, H * Dh)
    ol = (p_l @ vv).transpose(1, 2).reshape(b, tl, H * Dh)
    v = v + W[a + ".gamma_v"][0] * lin(W, a + ".attn.out_v_proj", ov)
    l = l + W[a + ".gamma_l"][0] * lin(W, a + ".attn.out_l_proj", ol)
    return v, l


# ---- RobertaLayer ----------------------------------------------------------------------------------------
def roberta_layer(W, p, x, keep_mask, drop=0.0):
    b, t, _ = x.shape
    ext = (1.0 - keep_mask[:, None, None, :].float()) * -10000.0
    sh = lambda y: y.view(b, t, 12, 64).transpose(1, 2)
    q, k, v = (sh(lin(W, f"{p}.attention.self.{n}", x)) for n in ("query", "key", "value"))
    pr = F.dropout(((q @ k.transpose(-1, -2)) / 8.0 + ext).softmax(-1), drop, drop > 0)
    ctx = (pr @ v).transpose(1, 2).reshape(b, t, 768)
    h = ln(W, p + ".attention.output.LayerNorm", F.dropout(lin(W, p + ".attention.output.dense", ctx), drop, drop > 0) + x)
    ff = F.gelu(lin(W, p + ".intermediate.dense", h))
    return ln(W, p + ".output.LayerNorm", F.dropout(lin(W, p + ".output.dense", ff), drop, drop > 0) + h)


# ---- MSDeformAttn module -----------------------------------------------------------------------------------
def msda_module(W, p, query, ref, inp, shapes, padding_mask):
    n, lq, _ = query.shape
    s = inp.shape[1]
    value = lin(W, p + ".value_proj", inp)
    if padding_mask is not None:
        value = value.masked_fill(padding_mask[..., None], 0.0)
    value = value.view(n, s, 8, 32)
    off = lin(W, p + ".sampling_offsets", query).view(n, lq, 8, 4, 4, 2)
    aw = lin(W, p + ".attention_weights", query).view(n, lq, 8, 16).softmax(-1).view(n, lq, 8, 4, 4)
    if ref.shape[-1] == 2:
        norm = torch.tensor([[w, h] for h, w in shapes], dtype=query.dtype)
        loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = ref[:, :, None, :, None, :2] + off / 4 * ref[:, :, None, :, None, 2:] * 0.5
    return lin(W, p + ".output_proj", msda_core(value, shapes, loc, aw))


def encoder_layer(W, p, src, pos, ref, shapes, padding_mask):
    src = ln(W, p + ".norm1", src + msda_module(W, p + ".self_attn", src + pos, ref, src, shapes, padding_mask))
    return ln(W, p + ".norm2", src + lin(W, p + ".linear2", F.relu(lin(W, p + ".linear1", src))))


def encoder(W, src, shapes, valid_ratios, pos, padding_mask, lang, lang_pad_mask, drop=0.0):
    """models/deformable_transformer.py:817-884 with fusion_interval=2, fusion_last_vis, lang_aux_loss."""
    refs = []
    for lvl, (h, w) in enumerate(shapes):
        ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h), torch.linspace(0.5, w - 0.5, w), indexing="ij")
        ry = ry.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * h)
        rx = rx.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * w)
        refs.append(torch.stack((rx, ry), -1))
    ref = torch.cat(refs, 1)[:, :, None] * valid_ratios[:, None]
    last = sum(h * w for h, w in shapes[:-1])
    keep_l = ~lang_pad_mask
    multi = []
    for idx in range(6):
        if idx % 2 == 0:
            k = idx // 2
            fused, lang = alif_block(W, f"transformer.encoder.VLFuse_layers.{k}", src[:, last:], lang, pos[:, last:], drop)
            src = torch.cat((src[:, :last], fused), 1)
            lang = roberta_layer(W, f"transformer.encoder.roberta_layers.{k}", lang, keep_l, drop)
            multi.append(lang)
        src = encoder_layer(W, f"transformer.encoder.layers.{idx}", src, pos, ref, shapes, padding_mask)
    return src, torch.stack(multi, 0)


# ---- decoders ------------------------------------------------------------------------------------------------
def sine_embed(pos_tensor):                               # deformable_transformer.py:1777-1802, 4-d boxes
    dim_t = 10000 ** (2 * torch.div(torch.arange(128, dtype=torch.float32), 2, rounding_mode="floor") / 128)
    e = []
    for col in (1, 0, 2, 3):
        p_ = (pos_tensor[:, :, col] * (2 * math.pi))[:, :, None] / dim_t
        e.append(torch.stack((p_[:, :, 0::2].sin(), p_[:, :, 1::2].cos()), 3).flatten(2))
    return torch.cat(e, 2)


def decoder_layer(W, p, tgt, qpos, ref, memory, shapes, padding_mask):
    qk = (tgt + qpos).transpose(0, 1)
    a, _ = F.multi_head_attention_forward(
        qk, qk, tgt.transpose(0, 1), 256, 8, W[p + ".self_attn.in_proj_weight"], W[p + ".self_attn.in_proj_bias"],
        None, None, False, 0.0, W[p + ".self_attn.out_proj.weight"], W[p + ".self_attn.out_proj.bias"],
        training=False, need_weights=False)
    tgt = ln(W, p + ".norm2", tgt + a.transpose(0, 1))
    tgt = ln(W, p + ".norm1", tgt + msda_module(W, p + ".cross_attn", tgt + qpos, ref, memory, shapes, padding_mask))
    return ln(W, p + ".norm3", tgt + lin(W, p + ".linear2", F.relu(lin(W, p + ".linear1", tgt))))


def dab_decoder(W, p, tgt, sub_ref, obj_ref, memory, shapes, valid_ratios, padding_mask, parse, bbox_offset):
    """deformable_transformer.py:1432-1552; box heads `sub_bbox_embed.{bbox_offset + lid}`."""
    vr4 = torch.cat([valid_ratios, valid_ratios], -1)[:, None]
    pair = obj_ref.shape[1]
    out, inter, isub, iobj = tgt, [], [], []
    for lid in range(3):
        if parse:
            ref_in = torch.cat((sub_ref[:, :, None] * vr4, obj_ref[:, :, None] * vr4), 1)
        else:
            ref_in = 0.5 * (sub_ref + obj_ref)[:, :, None] * vr4
        raw = mlp(W, p + ".ref_point_head", sine_embed(ref_in[:, :, 0, :]), 2)
        qpos = raw if lid == 0 else mlp(W, p + ".query_scale", out, 2) * raw
        out = decoder_layer(W, f"{p}.layers.{lid}", out, qpos, ref_in, memory, shapes, padding_mask)
        s_in, o_in = (out[:, :pair], out[:, pair:]) if parse else (out, out)
        sub_ref = (mlp(W, f"sub_bbox_embed.{bbox_offset + lid}", s_in, 3) + inverse_sigmoid(sub_ref)).sigmoid().detach()
        obj_ref = (mlp(W, f"obj_bbox_embed.{bbox_offset + lid}", o_in, 3) + inverse_sigmoid(obj_ref)).sigmoid().detach()
        inter.append(out)
        isub.append(sub_ref)
        iobj.append(obj_ref)
    return torch.stack(inter), isub, iobj


def mbf(W, p, a, b):                                      # deformable_transformer.py:1063-1068
    outs = [lin(W, f"{p}.fc_3.{c}", F.relu(lin(W, f"{p}.fc_1.{c}", a) * lin(W, f"{p}.fc_2.{c}", b))) for c in range(16)]
    return F.relu(torch.stack(outs).sum(0))


# ---- backbone / text (third party) ----------------------------------------------------------------------------
def build_third_party(W):
    import torchvision
    from torchvision.ops.misc import FrozenBatchNorm2d
    from transformers import RobertaModel
    from rlipv2_b200.text_encoder import roberta_base_config
    resnet = torchvision.models.resnet50(weights=None, norm_layer=FrozenBatchNorm2d)
    sd = {k[len("backbone.0.body."):]: v for k, v in W.items() if k.startswith("backbone.0.body.")}
    resnet.load_state_dict(sd, strict=False)               # fc.* absent: IntermediateLayerGetter drops it
    for n, prm in resnet.named_parameters():               # DDETR_backbone.py:75-77
        prm.requires_grad_(("layer2" in n) or ("layer3" in n) or ("layer4" in n))
    text = RobertaModel(roberta_base_config())
    text.load_state_dict({k[len("transformer.text_encoder."):]: v for k, v in W.items()
                          if k.startswith("transformer.text_encoder.")}, strict=False)
    return resnet.eval(), text


def backbone_features(resnet, x):
    x = resnet.maxpool(resnet.relu(resnet.bn1(resnet.conv1(x))))
    c2 = resnet.layer1(x)
    c3 = resnet.layer2(c2)
    c4 = resnet.layer3(c3)
    c5 = resnet.layer4(c4)
    return [c3, c4, c5]


def sine_pos(mask):                                        # models/position_encoding.py:37-58, normalize=True
    not_mask = ~mask
    y = not_mask.cumsum(1, dtype=torch.float32)
    x = not_mask.cumsum(2, dtype=torch.float32)
    y = y / (y[:, -1:, :] + 1e-6) * 2 * math.pi
    x = x / (x[:, :, -1:] + 1e-6) * 2 * math.pi
    dim_t = 10000 ** (2 * torch.div(torch.arange(128, dtype=torch.float32), 2, rounding_mode="floor") / 128)
    px, py = x[:, :, :, None] / dim_t, y[:, :, :, None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), 4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), 4).flatten(3)
    return torch.cat((py, px), 3).permute(0, 3, 1, 2)


# ---- full forward ---------------------------------------------------------------------------------------------
def forward_step(W, third, images, mask, text, tokenizer, drop=0.0):
    """phase A + phase B -> (outputs dict, extras).  images [B,3,H,W], mask [B,H,W] bool (True = pad)."""
    resnet, text_enc = third
    feats = backbone_features(resnet, images)
    srcs, masks, poss = [], [], []
    for l, f in enumerate(feats):
        m = F.interpolate(mask[None].float(), size=f.shape[-2:]).to(torch.bool)[0]
        srcs.append(F.group_norm(F.conv2d(f, W[f"input_proj.{l}.0.weight"], W[f"input_proj.{l}.0.bias"]), 32,
                                 W[f"input_proj.{l}.1.weight"], W[f"input_proj.{l}.1.bias"]))
        masks.append(m)
        poss.append(sine_pos(m))
    s3 = F.group_norm(F.conv2d(feats[-1], W["input_proj.3.0.weight"], W["input_proj.3.0.bias"], stride=2, padding=1), 32,
                      W["input_proj.3.1.weight"], W["input_proj.3.1.bias"])
    m3 = F.interpolate(mask[None].float(), size=s3.shape[-2:]).to(torch.bool)[0]
    srcs.append(s3)
    masks.append(m3)
    poss.append(sine_pos(m3))
    shapes = [tuple(s.shape[-2:]) for s in srcs]
    bs = images.shape[0]
    src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    pmask = torch.cat([m.flatten(1) for m in masks], 1)
    pos = torch.cat([p.flatten(2).transpose(1, 2) + W["transformer.level_embed"][l].view(1, 1, -1)
                     for l, p in enumerate(poss)], 1)

    def vr(m):
        _, H, Wd = m.shape
        return torch.stack([(~m[:, 0, :]).sum(1).float() / Wd, (~m[:, :, 0]).sum(1).float() / H], -1)
    valid_ratios = torch.stack([vr(m) for m in masks], 1)
    # label strings -> pooled vectors (deformable_transformer.py:489-522)
    objs, verbs = text[0]
    tok = tokenizer.batch_encode_plus(list(objs) + list(verbs), padding="longest", return_tensors="pt")
    pooled = text_enc(**tok).pooler_output
    text_memory = torch.cat([pad_sequence([pooled[:len(objs)]]), pad_sequence([pooled[len(objs):]])], 0)   # [T,1,768]
    text_pad = ~(text_memory.sum(-1) > 0)
    lang = text_memory.repeat(1, bs, 1).transpose(0, 1)
    lang_pad = text_pad.repeat(1, bs).transpose(0, 1)
    memory, lang_out = encoder(W, src, shapes, valid_ratios, pos, pmask, lang, lang_pad, drop)
    text_res = F.dropout(F.layer_norm(lin(W, "transformer.resizer.fc", lang_out.transpose(1, 2)), (256,),
                                      W["transformer.resizer.layer_norm.weight"], W["transformer.resizer.layer_norm.bias"],
                                      1e-12), drop, drop > 0)                                  # [3, T, bs, 256]
    # phase B (deformable_transformer.py:635-700)
    nq = W["tgt_embed.weight"].shape[0]
    refp = W["refpoint_embed.weight"].sigmoid()
    sub0, obj0 = refp[:nq // 2], refp[nq // 2:]
    tgt = W["tgt_embed.weight"].unsqueeze(0).expand(bs, -1, -1)
    vt = W["verb_tgt_embed.weight"].unsqueeze(0).expand(bs, -1, -1)
    hs_ho, isub, iobj = dab_decoder(W, "transformer.ho_decoder", tgt, sub0[None].repeat(bs, 1, 1), obj0[None].repeat(bs, 1, 1),
                                    memory, shapes, valid_ratios, pmask, True, 0)
    merge = mbf(W, "transformer.verb_tgt_generator", hs_ho[-1][:, :nq // 2], hs_ho[-1][:, nq // 2:]) \
        + vt[:, :nq // 2] + vt[:, nq // 2:]
    hs_verb, _, _ = dab_decoder(W, "transformer.verb_decoder", merge, isub[-1], iobj[-1], memory, shapes, valid_ratios,
                                pmask, False, 3)
    # heads (hoi.py:2116-2157)
    n_obj = len(objs)
    outs = []
    for lvl in range(3):
        sref, oref = (sub0, obj0) if lvl == 0 else (isub[lvl - 1], iobj[lvl - 1])
        hs_h, hs_o = hs_ho[lvl][:, :nq // 2], hs_ho[lvl][:, nq // 2:]
        tm = F.normalize(text_res[lvl].transpose(0, 1), p=2, dim=-1)
        pt = lin(W, "projection_text", tm / 2.0)
        ot, vtx = pt[:, :n_obj], pt[:, n_obj:]
        outs.append({
            "pred_sub_logits": torch.einsum("bcd,bed->bce", hs_h + W["bias_obj_a"], ot) + (-math.log(99.0)),
            "pred_obj_logits": torch.einsum("bcd,bed->bce", hs_o + W["bias_obj_a"], ot) + (-math.log(99.0)),
            "pred_verb_logits": torch.einsum("bcd,bed->bce", hs_verb[lvl] + W["bias_pred_a"], vtx) + (-math.log(99.0)),
            "pred_sub_boxes": (mlp(W, f"sub_bbox_embed.{lvl}", hs_h, 3) + inverse_sigmoid(sref)).sigmoid(),
            "pred_obj_boxes": (mlp(W, f"obj_bbox_embed.{lvl}", hs_o, 3) + inverse_sigmoid(oref)).sigmoid(),
        })
    out = dict(outs[-1])
    out["aux_outputs"] = outs[:-1]
    return out, {"img_memory": memory, "text_memory_resized": text_res, "text_attention_mask": text_pad.repeat(1, bs),
                 "valid_ratios": valid_ratios}


# ---- matcher + criterion ----------------------------------------------------------------------------------------
def cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def giou(a, b):                                            # util/box_ops.py:34-73
    a1 = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    a2 = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    wh = (torch.min(a[:, None, 2:], b[:, 2:]) - torch.max(a[:, None, :2], b[:, :2])).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = a1[:, None] + a2 - inter
    wh2 = (torch.max(a[:, None, 2:], b[:, 2:]) - torch.min(a[:, None, :2], b[:, :2])).clamp(min=0)
    area = wh2[..., 0] * wh2[..., 1]
    return inter / union - (area - union) / area


@torch.no_grad()
def matcher(o, targets, w_bbox=2.5, w_giou=1.0):           # matcher.py:108-202 (subject_class branch)
    bs, nq = o["pred_obj_logits"].shape[:2]
    ps = o["pred_sub_logits"].flatten(0, 1).softmax(-1)
    po = o["pred_obj_logits"].flatten(0, 1).softmax(-1)
    pv = o["pred_verb_logits"].flatten(0, 1).sigmoid()
    sb, ob = o["pred_sub_boxes"].flatten(0, 1), o["pred_obj_boxes"].flatten(0, 1)
    tv = torch.cat([t["verb_labels"] for t in targets]).permute(1, 0)
    ts_, to_ = torch.cat([t["sub_boxes"] for t in targets]), torch.cat([t["obj_boxes"] for t in targets])
    c_sub = -ps[:, torch.cat([t["sub_labels"] for t in targets])]
    c_obj = -po[:, torch.cat([t["obj_labels"] for t in targets])]
    c_verb = -(pv.matmul(tv) / (tv.sum(0, keepdim=True) + 1e-4)
               + (1 - pv).matmul(1 - tv) / ((1 - tv).sum(0, keepdim=True) + 1e-4)) / 2
    cb = torch.stack((torch.cdist(sb, ts_, p=1), torch.cdist(ob, to_, p=1) * (to_ != 0).any(1).unsqueeze(0))).max(0)[0]
    gs = -giou(cxcywh_to_xyxy(sb), cxcywh_to_xyxy(ts_))
    go = -giou(cxcywh_to_xyxy(ob), cxcywh_to_xyxy(to_)) + gs * (to_ == 0).all(1).unsqueeze(0)
    cg = torch.stack((gs, go)).max(0)[0]
    C = 1 * c_obj + 1 * c_sub + 1 * c_verb + w_bbox * cb + w_giou * cg
    C = C.view(bs, nq, -1)
    sizes = [len(t["obj_labels"]) for t in targets]
    idx = [linear_sum_assignment(c[i]) for i, c in enumerate(C.split(sizes, -1))]
    return [(torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)) for i, j in idx], cg


def criterion(out, targets, eos_coef=0.1, bbox_coef=2.5, giou_coef=1.0):
    """-> (loss dict with the reference's keys, weighted total, list of match indices per decoder layer)."""
    losses, matches = {}, []
    n_int = max(float(sum(len(t["obj_labels"]) for t in targets)), 1.0)
    layers = [({k: v for k, v in out.items() if k != "aux_outputs"}, "")] + \
             [(a, f"_{i}") for i, a in enumerate(out["aux_outputs"])]
    for o, sfx in layers:
        ind, cg = matcher(o, targets)
        matches.append(ind)
        bidx = torch.cat([torch.full_like(s, i) for i, (s, _) in enumerate(ind)])
        sidx = torch.cat([s for s, _ in ind])

        def wce(logits, key):                               # hoi.py:3704-3740
            w = torch.ones(logits.shape[-1])
            w[-1] = eos_coef
            cls = torch.full(logits.shape[:2], logits.shape[-1] - 1, dtype=torch.int64)
            matched = torch.cat([t[key][J] for t, (_, J) in zip(targets, ind)])
            cls[bidx, sidx] = matched
            return F.cross_entropy(logits.transpose(1, 2), cls, w), matched
        lo, mo = wce(o["pred_obj_logits"], "obj_labels")
        ls, ms = wce(o["pred_sub_logits"], "sub_labels")
        losses["loss_obj_ce" + sfx] = lo + ls
        if sfx == "":
            acc = lambda lg, t_: 100 - (lg[bidx, sidx].argmax(-1) == t_).float().mean() * 100
            losses["obj_class_error"] = acc(o["pred_obj_logits"], mo)
            losses["sub_class_error"] = acc(o["pred_sub_logits"], ms)
        # verb labels with GIoU soft targets + quality focal loss (hoi.py:3932-3957, 4481-4495)
        g = -cg
        q0 = t0 = 0
        soft = []
        for t, (I, J) in zip(targets, ind):
            soft.append(t["verb_labels"][J] * ((g[q0 + I, t0 + J] + 1) / 2).unsqueeze(-1))
            q0 += o["pred_verb_logits"].shape[1]
            t0 += J.shape[0]
        tgt = torch.zeros_like(o["pred_verb_logits"])
        tgt[bidx, sidx] = torch.cat(soft)
        pr = o["pred_verb_logits"].sigmoid().clamp(1e-6, 1 - 1e-6)
        ql = (tgt - pr).abs().pow(2) * ((1 - tgt) * torch.log(1 - pr) + tgt * torch.log(pr))
        npos = tgt.gt(0).float().sum()
        losses["loss_verb_ce" + sfx] = -ql.sum() / npos if npos > 0 else -ql.sum()
        # boxes (hoi.py:4162-4193)
        ssb, sob = o["pred_sub_boxes"][bidx, sidx], o["pred_obj_boxes"][bidx, sidx]
        tsb = torch.cat([t["sub_boxes"][j] for t, (_, j) in zip(targets, ind)])
        tob = torch.cat([t["obj_boxes"][j] for t, (_, j) in zip(targets, ind)])
        ex = (tob != 0).any(1)
        losses["loss_sub_bbox" + sfx] = F.l1_loss(ssb, tsb, reduction="none").sum() / n_int
        losses["loss_obj_bbox" + sfx] = (F.l1_loss(sob, tob, reduction="none") * ex.unsqueeze(1)).sum() / (ex.sum() + 1e-4)
        losses["loss_sub_giou" + sfx] = (1 - torch.diag(giou(cxcywh_to_xyxy(ssb), cxcywh_to_xyxy(tsb)))).sum() / n_int
        losses["loss_obj_giou" + sfx] = ((1 - torch.diag(giou(cxcywh_to_xyxy(sob), cxcywh_to_xyxy(tob)))) * ex).sum() / (ex.sum() + 1e-4)
        card = (o["pred_obj_logits"].argmax(-1) != o["pred_obj_logits"].shape[-1] - 1).sum(1).float()
        losses["obj_cardinality_error" + sfx] = F.l1_loss(card, torch.tensor([float(len(t["obj_labels"])) for t in targets]))
    coef = {"loss_obj_ce": 1.0, "loss_verb_ce": 1.0, "loss_sub_bbox": bbox_coef, "loss_obj_bbox": bbox_coef,
            "loss_sub_giou": giou_coef, "loss_obj_giou": giou_coef}
    base = lambda k: k[:-2] if k.endswith(("_0", "_1")) else k
    total = sum(v * coef[base(k)] for k, v in losses.items() if base(k) in coef)
    return losses, total, matches


# ---- weights + timing helper (bench.py) ---------------------------------------------------------------------------
def synthetic_weights(num_queries, seed=0):
    """Name-keyed deterministic weights with the reference's state_dict layout (shapes from the committed
    key listing; the three query-embedding tables resized to `num_queries`)."""
    from oracle.detfill import det_fill_
    shapes = json.load(open(os.path.join(GOLDEN, "parseda_state_dict_keys.json")))
    W = {}
    for k, s in shapes.items():
        if k in ("tgt_embed.weight", "verb_tgt_embed.weight", "refpoint_embed.weight"):
            s = [num_queries] + s[1:]
        W[k] = torch.zeros(s)
    det_fill_(W, seed)
    for k in list(W):                                      # bbox-head aliases share storage in the model
        if k.startswith("transformer.ho_decoder.sub_bbox_embed.") or k.startswith("transformer.ho_decoder.obj_bbox_embed."):
            W[k.replace("transformer.ho_decoder.", "")] = W[k]
        elif k.startswith("transformer.verb_decoder.sub_bbox_embed.") or k.startswith("transformer.verb_decoder.obj_bbox_embed."):
            head, idx, rest = k.replace("transformer.verb_decoder.", "").split(".", 2)
            W[f"{head}.{int(idx) + 3}.{rest}"] = W[k]
    return W


def time_train_step_sample(threads, budget_s=25.0, batch=2, steps=1, warmup=0, total_budget_s=240.0):
    """REAL optimisation steps of the oracle on the host cores: forward + losses + backward + clip_grad_norm_(0.1) + AdamW
    (the three learning-rate groups of main.py:523-539) on `batch` synthetic 3x800x1333 images, 300 queries, 256 labels -
    BASELINE config 2, the configuration bench.py's GPU arm times.  `warmup` untimed and `steps` timed steps, each one whole
    step; when the whole run would not fit `total_budget_s` (the box is slower than expected) the sample shrinks to
    ONE image per step - still a full-resolution step, reported per image - and, failing that, fewer timed steps
    (`timed_steps` says how many).  Called with steps = 1, warmup = 0 (bench.py's `cpu_baseline` leg) it times one
    single-image step within `budget_s`.  value = images processed / time, never extrapolated across resolutions."""
    from rlipv2_b200.synth import synthetic_batch, synthetic_text          # (module without any C-ABI library)
    from rlipv2_b200.text_encoder import HashTokenizer
    torch.set_num_threads(threads)
    W = synthetic_weights(300)
    third = build_third_party(W)
    resnet, text_model = third
    for k in W:
        if not k.startswith("backbone.") and not k.startswith("transformer.text_encoder."):
            W[k].requires_grad_(True)
    groups = [[v for k, v in W.items() if v.requires_grad and not k.startswith("transformer.text_encoder.")],
              [p for p in resnet.parameters() if p.requires_grad],
              [p for p in text_model.parameters() if p.requires_grad]]
    seen, uniq = set(), []
    for g in groups:                                        # bbox-head aliases share storage: one optimizer slot each
        u = []
        for p in g:
            if id(p) not in seen:
                seen.add(id(p))
                u.append(p)
        uniq.append(u)
    opt = torch.optim.AdamW([{"params": uniq[0]}, {"params": uniq[1], "lr": 1.41e-5}, {"params": uniq[2], "lr": 1.41e-5}],
                            lr=1.41e-4, weight_decay=1e-4)
    params = [p for g in uniq for p in g]
    text = synthetic_text(170, 85)
    h, w = 800, 1333

    def one(nimg):
        images, targets = synthetic_batch(nimg, h, w, pin=False)
        t0 = time.perf_counter()
        out, _ = forward_step(W, third, images, torch.zeros(nimg, h, w, dtype=torch.bool), text, HashTokenizer(), drop=0.1)
        _, total, _ = criterion(out, targets)
        opt.zero_grad(set_to_none=True)
        total.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 0.1)
        opt.step()
        return time.perf_counter() - t0

    nimg = 1 if (steps <= 1 and warmup <= 0) else batch
    first = one(nimg)                                      # first pass: doubles as warm-up and as the cost probe
    if nimg > 1 and first * (steps + max(0, warmup - 1)) > total_budget_s:
        nimg = 1                                            # bounded sample: one full-resolution image per step
        first = one(nimg)
    for _ in range(max(0, warmup - 1) if steps > 1 or warmup > 0 else 0):
        one(nimg)
    times, spent = [], 0.0
    if steps <= 1 and warmup <= 0:
        times = [first]
    else:
        while len(times) < steps and (not times or spent + times[-1] <= total_budget_s):
            times.append(one(nimg))
            spent += times[-1]
    dt = sum(times) / len(times)
    return {"value": nimg / dt, "unit": "images/s", "cores": threads, "kind": "port", "timed_steps": len(times),
            "images_per_step": nimg, "step_s": dt,
            "sample": f"oracle/parseda_oracle.py: whole optimisation steps (fwd + loss + bwd + clip + AdamW) on {nimg} "
                      f"image(s) 800x1333, 300 queries, 256 labels: {dt:.2f} s per step (mean of {len(times)})"}
