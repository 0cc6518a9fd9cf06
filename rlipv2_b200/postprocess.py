"""Evaluation post-processing of the ParSeDA outputs.

`PostProcessHOI` / `PostProcessSGG` mirror /root/reference/models/hoi.py:4769-4873 and :4876-4938
(constructed at models/detr.py:683-691): same constructor arguments, same per-image result dicts
(`labels`, `boxes`, `verb_scores`, `sub_ids`, `obj_ids`, CPU tensors) consumed by
engine.evaluate_hoi_with_text (engine.py:423-442) and the HICO / V-COCO / OI evaluators.

What differs is the execution: the reference does the arithmetic on the device and then issues four
`.to('cpu')` copies per image inside a python loop (hoi.py:4852-4872), i.e. 4*bs device->host syncs
per batch.  Here every quantity is computed batched, packed into ONE flat device buffer and read
back with a single copy; the per-image dicts are views of that host buffer.  The arithmetic
(softmax, max over the real classes, sigmoid, cxcywh->xyxy, scaling) is the reference's op for op,
so scores/boxes are bit-identical.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .nested import box_cxcywh_to_xyxy


def _scaled_boxes(boxes, target_sizes):
    """hoi.py:4838-4843: cxcywh in [0,1] -> xyxy in pixels of the original image"""
    img_h, img_w = target_sizes.unbind(1)
    scale = torch.stack([img_w, img_h, img_w, img_h], dim=1).to(boxes.device)
    return box_cxcywh_to_xyxy(boxes) * scale[:, None, :]


def _one_readback(labels, boxes, verb_scores):
    """[bs,2Q] int64, [bs,2Q,4] f32, [bs,Q,V] f32 -> the same three as CPU tensors, one D2H copy.
    Labels travel as exact fp32 integers (class ids < 2^24)."""
    bs, q2 = labels.shape
    V = verb_scores.shape[-1]
    flat = torch.cat([labels.to(torch.float32).reshape(bs, -1), boxes.reshape(bs, -1),
                      verb_scores.reshape(bs, -1)], dim=1)
    host = flat.to("cpu")
    n_l, n_b = q2, q2 * 4
    lab = host[:, :n_l].to(torch.int64)
    box = host[:, n_l:n_l + n_b].reshape(bs, q2, 4)
    vs = host[:, n_l + n_b:].reshape(bs, q2 // 2, V)
    return lab, box, vs


class PostProcessHOI(nn.Module):
    """hoi.py:4769-4873.  `obj_verb_co` (the prior the reference registers as a buffer but only uses in a
    commented-out alternative, :4787-4791, :4865) is loaded when datasets/priors/ is present and is
    otherwise left empty; it never enters the scores."""

    def __init__(self, subject_category_id, sigmoid=True, temperature=False, zero_shot_hoi_eval=False,
                 verb_curing=False, priors_path="datasets/priors/obj_verb_cooccurrence.npz"):
        super().__init__()
        self.subject_category_id = subject_category_id
        self.sigmoid = sigmoid
        self.temperature = temperature
        self.tao = 0.07 if temperature else None
        self.zero_shot_hoi_eval = zero_shot_hoi_eval
        self.verb_curing = verb_curing
        if verb_curing:
            assert sigmoid
        co = torch.zeros(0, 0)
        try:
            import numpy as np
            m = torch.tensor(np.load(priors_path)["cond_prob_co_matrices"]).float()
            m = m + 0.1 / m.shape[1]
            co = m / m.sum(dim=1).unsqueeze(dim=1)
        except (OSError, KeyError):
            pass
        self.register_buffer("obj_verb_co", co)

    def _softmax(self, logits):
        return F.softmax(logits / self.tao, -1) if self.temperature else F.softmax(logits, -1)

    @torch.no_grad()
    def forward(self, outputs, target_sizes):
        obj_logits, verb_logits = outputs["pred_obj_logits"], outputs["pred_verb_logits"]
        assert len(obj_logits) == len(target_sizes)
        assert target_sizes.shape[1] == 2
        obj_scores, obj_labels = self._softmax(obj_logits)[..., :-1].max(-1)          # :4823-4827
        if self.sigmoid:
            verb_scores = verb_logits.sigmoid()
            if self.verb_curing:
                verb_scores = verb_scores * outputs["curing_score"]
        else:
            verb_scores = verb_logits
        sub_boxes = _scaled_boxes(outputs["pred_sub_boxes"], target_sizes)
        obj_boxes = _scaled_boxes(outputs["pred_obj_boxes"], target_sizes)
        labels = torch.cat((torch.full_like(obj_labels, self.subject_category_id), obj_labels), dim=1)
        boxes = torch.cat((sub_boxes, obj_boxes), dim=1)
        vs = verb_scores * obj_scores.unsqueeze(-1)                                   # :4861
        keep = None
        if self.zero_shot_hoi_eval:                                                   # :4803-4811, :4847-4850
            assert "pred_sub_logits" in outputs
            _, sub_labels = self._softmax(outputs["pred_sub_logits"])[..., :-1].max(-1)
            keep = (sub_labels == self.subject_category_id).to("cpu")
        lab, box, vs = _one_readback(labels, boxes, vs)
        Q = obj_labels.shape[1]
        results = []
        for b in range(lab.shape[0]):
            l, bx, v = lab[b], box[b], vs[b]
            if keep is not None:
                k = keep[b]
                l = torch.cat((l[:Q][k], l[Q:][k]))
                bx = torch.cat((bx[:Q][k], bx[Q:][k]))
                v = v[k]
            ids = torch.arange(bx.shape[0])
            results.append({"labels": l, "boxes": bx, "verb_scores": v,
                            "sub_ids": ids[:ids.shape[0] // 2], "obj_ids": ids[ids.shape[0] // 2:]})
        return results


class PostProcessSGG(nn.Module):
    """hoi.py:4876-4938: subject labels are predicted too, and the triplet score is
    verb * object score * subject score."""

    def __init__(self, sigmoid=True, zero_shot_sgg_eval=False):
        super().__init__()
        self.sigmoid = sigmoid
        self.zero_shot_sgg_eval = zero_shot_sgg_eval

    @torch.no_grad()
    def forward(self, outputs, target_sizes):
        obj_logits = outputs["pred_obj_logits"]
        assert len(obj_logits) == len(target_sizes)
        assert target_sizes.shape[1] == 2
        obj_scores, obj_labels = F.softmax(obj_logits, -1)[..., :-1].max(-1)
        sub_scores, sub_labels = F.softmax(outputs["pred_sub_logits"], -1)[..., :-1].max(-1)
        verb_scores = outputs["pred_verb_logits"].sigmoid() if self.sigmoid else outputs["pred_verb_logits"]
        sub_boxes = _scaled_boxes(outputs["pred_sub_boxes"], target_sizes)
        obj_boxes = _scaled_boxes(outputs["pred_obj_boxes"], target_sizes)
        labels = torch.cat((sub_labels, obj_labels), dim=1)
        boxes = torch.cat((sub_boxes, obj_boxes), dim=1)
        vs = verb_scores * obj_scores.unsqueeze(-1) * sub_scores.unsqueeze(-1)         # :4930
        lab, box, vs = _one_readback(labels, boxes, vs)
        results = []
        for b in range(lab.shape[0]):
            ids = torch.arange(box[b].shape[0])
            results.append({"labels": lab[b], "boxes": box[b], "verb_scores": vs[b],
                            "sub_ids": ids[:ids.shape[0] // 2], "obj_ids": ids[ids.shape[0] // 2:]})
        return results


def build_postprocessors(args):
    """models/detr.py:683-691"""
    sigmoid = not (args.verb_loss_type == "focal_without_sigmoid")
    if args.hoi:
        return {"hoi": PostProcessHOI(getattr(args, "subject_category_id", 0), sigmoid=sigmoid,
                                      temperature=("with_tem" in args.obj_loss_type),
                                      zero_shot_hoi_eval=(getattr(args, "zero_shot_eval", None) in ["hico", "v-coco"]),
                                      verb_curing=getattr(args, "verb_curing", False))}
    if args.sgg:
        return {"sgg": PostProcessSGG(sigmoid=sigmoid)}
    return {}
