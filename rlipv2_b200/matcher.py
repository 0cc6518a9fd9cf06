"""HungarianMatcherHOI - bipartite matching of (subject, object, verb) predictions to ground truth.

Mirror of /root/reference/models/matcher.py:95-269: the cost matrix is built on the device in fp32
in the reference's operation order (LSAP ties flip with 1-ulp changes, so the order is part of the
contract), copied to the host once, and solved per image with the same
`scipy.optimize.linear_sum_assignment` call.  Index outputs are int64 CPU tensors and must be
bit-exact with the reference.
"""
import torch
from scipy.optimize import linear_sum_assignment
from torch import nn
from torch.nn.utils.rnn import pad_sequence

from .nested import box_cxcywh_to_xyxy, generalized_box_iou


class HungarianMatcherHOI(nn.Module):
    def __init__(self, cost_obj_class: float = 1, cost_verb_class: float = 1, cost_bbox: float = 1,
                 cost_giou: float = 1, subject_class=False):
        super().__init__()
        self.cost_obj_class = cost_obj_class
        self.cost_verb_class = cost_verb_class
        self.cost_bbox = cost_bbox
        self.cost_giou = cost_giou
        self.subject_class = subject_class
        assert cost_obj_class != 0 or cost_verb_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"

    @staticmethod
    def _verb_targets(targets, subject_class):
        """[n_verbs, total_targets] target matrix.  With subject_class the reference pads the
        per-triplet label rows to the longest label set (matcher.py:123-140); without it, a plain
        concatenation (matcher.py:217-219)."""
        if not subject_class:
            return torch.cat([v["verb_labels"] for v in targets]).permute(1, 0)
        max_len = max(v["verb_labels"].shape[1] for v in targets)
        rows, pad_row = [], False
        for v in targets:
            vl = v["verb_labels"]
            if vl.shape[0] > 0:
                rows.extend(r.reshape(-1, 1) for r in vl.split(1, dim=0))
            elif vl.shape[1] == max_len:
                pad_row = True
        if pad_row:
            rows.append(torch.zeros((max_len, 1), device=targets[0]["verb_labels"].device))
        t = pad_sequence(rows).squeeze(-1)            # [max_len, n_rows]
        if pad_row:
            t = t[:, :t.shape[1] - 1]
        return t

    KEYS = ("pred_sub_logits", "pred_obj_logits", "pred_verb_logits", "pred_sub_boxes", "pred_obj_boxes")

    @torch.no_grad()
    def compute_costs_layers(self, layers, targets):
        """Costs of several decoder layers in one pass.  Every op of `compute_costs` is row-wise over the
        (image, query) rows, so the layers' predictions are concatenated along the batch axis and the ~60
        small kernels run once instead of once per layer; each layer's entries are bit-identical to a
        per-layer call.  -> (C [n_layers, bs, nq, T], [cost_list per layer] (views))"""
        n = len(layers)
        if n == 1:
            C, cl = self.compute_costs(layers[0], targets)
            return C.unsqueeze(0), [cl]
        stacked = {k: torch.cat([l[k] for l in layers], 0) for k in self.KEYS if k in layers[0]}
        bs, nq = layers[0]["pred_obj_logits"].shape[:2]
        C, cl = self.compute_costs(stacked, targets, row_chunks=n)    # rows ordered (layer, image, query)
        C = C.view(n, bs, nq, -1)
        rows = bs * nq

        def part(x, i):
            if isinstance(x, tuple):
                return tuple(part(y, i) for y in x)
            return x[i * rows:(i + 1) * rows]
        return C, [[part(x, i) for x in cl] for i in range(n)]

    @torch.no_grad()
    def match_layers(self, layers, targets):
        """[(indices, cost_list)] for every decoder layer with ONE device->host copy of all cost tensors
        (the reference does one per layer, twice per layer with --giou_verb_label: matcher.py:185,
        hoi.py:3933)."""
        C, cost_lists = self.compute_costs_layers(layers, targets)
        sizes = [len(v["obj_labels"]) for v in targets]
        C_cpu = C.cpu()
        return [(self.solve(C_cpu[i], sizes), cost_lists[i]) for i in range(len(layers))]

    @torch.no_grad()
    def forward(self, outputs, targets, return_cost=False):
        C, cost_list = self.compute_costs(outputs, targets)
        indices = self.solve(C.cpu(), [len(v["obj_labels"]) for v in targets])   # the one device->host copy
        return (indices, cost_list) if return_cost else indices

    @staticmethod
    def solve(C_cpu, sizes):
        """scipy LSAP per image on the host copy of the cost tensor [bs, nq, sum(sizes)] (matcher.py:187-193)."""
        indices = [linear_sum_assignment(c[i]) for i, c in enumerate(C_cpu.split(sizes, -1))]
        return [(torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)) for i, j in indices]

    @torch.no_grad()
    def compute_costs(self, outputs, targets, row_chunks=1):
        """Device part (capturable in a CUDA graph): -> (C [bs, nq, T], cost_list as matcher.py:197-199)."""
        bs, num_queries = outputs["pred_obj_logits"].shape[:2]
        out_obj_prob = outputs["pred_obj_logits"].flatten(0, 1).softmax(-1)
        out_verb_prob = outputs["pred_verb_logits"].flatten(0, 1).sigmoid()
        out_sub_bbox = outputs["pred_sub_boxes"].flatten(0, 1)
        out_obj_bbox = outputs["pred_obj_boxes"].flatten(0, 1)
        tgt_obj_labels = torch.cat([v["obj_labels"] for v in targets])
        tgt_sub_boxes = torch.cat([v["sub_boxes"] for v in targets])
        tgt_obj_boxes = torch.cat([v["obj_boxes"] for v in targets])
        tgt_verb = self._verb_targets(targets, self.subject_class)          # [n_verbs, T]

        cost_obj_class = -out_obj_prob[:, tgt_obj_labels]
        if self.subject_class:
            out_sub_prob = outputs["pred_sub_logits"].flatten(0, 1).softmax(-1)
            tgt_sub_labels = torch.cat([v["sub_labels"] for v in targets])
            cost_sub_class = -out_sub_prob[:, tgt_sub_labels]
            if out_verb_prob.shape[1] - 1 == tgt_verb.shape[0]:              # trailing "no verb" column
                out_verb_prob = out_verb_prob[:, :out_verb_prob.shape[1] - 1]

        def mm(a, b):
            # the one contraction of the cost build.  GEMM libraries pick kernels (hence summation orders) by
            # row count, so the stacked call multiplies layer by layer: costs - and the LSAP tie-breaks that
            # hang on their last bit - stay identical to a per-layer call.
            if row_chunks == 1:
                return a.matmul(b)
            return torch.cat([c.matmul(b) for c in a.chunk(row_chunks, 0)], 0)

        cost_verb_class = -(mm(out_verb_prob, tgt_verb) / (tgt_verb.sum(dim=0, keepdim=True) + 1e-4)
                            + mm(1 - out_verb_prob, 1 - tgt_verb)
                            / ((1 - tgt_verb).sum(dim=0, keepdim=True) + 1e-4)) / 2

        cost_sub_bbox = torch.cdist(out_sub_bbox, tgt_sub_boxes, p=1)
        cost_obj_bbox = torch.cdist(out_obj_bbox, tgt_obj_boxes, p=1) * (tgt_obj_boxes != 0).any(dim=1).unsqueeze(0)
        if cost_sub_bbox.shape[1] == 0:
            cost_bbox = cost_sub_bbox
        else:
            cost_bbox = torch.stack((cost_sub_bbox, cost_obj_bbox)).max(dim=0)[0]

        cost_sub_giou = -generalized_box_iou(box_cxcywh_to_xyxy(out_sub_bbox), box_cxcywh_to_xyxy(tgt_sub_boxes),
                                             check=False)
        cost_obj_giou = -generalized_box_iou(box_cxcywh_to_xyxy(out_obj_bbox), box_cxcywh_to_xyxy(tgt_obj_boxes),
                                             check=False) \
            + cost_sub_giou * (tgt_obj_boxes == 0).all(dim=1).unsqueeze(0)
        if cost_sub_giou.shape[1] == 0:
            cost_giou = cost_sub_giou
        else:
            cost_giou = torch.stack((cost_sub_giou, cost_obj_giou)).max(dim=0)[0]

        if self.subject_class:
            C = self.cost_obj_class * cost_obj_class + self.cost_obj_class * cost_sub_class + \
                self.cost_verb_class * cost_verb_class + \
                self.cost_bbox * cost_bbox + self.cost_giou * cost_giou
        else:
            C = self.cost_obj_class * cost_obj_class + self.cost_verb_class * cost_verb_class + \
                self.cost_bbox * cost_bbox + self.cost_giou * cost_giou
        C = C.view(bs, num_queries, -1)
        cost_list = [cost_giou, (cost_sub_giou, cost_obj_giou), cost_bbox, (cost_sub_bbox, cost_obj_bbox),
                     cost_verb_class]
        cost_list += [cost_sub_class, cost_obj_class] if self.subject_class else [cost_obj_class]
        return C, cost_list


def build_matcher(args):
    """matcher.py:272-278 (HOI / SGG / cross-modal pre-training all use the HOI matcher)."""
    if not (args.hoi or args.sgg or args.cross_modal_pretrain):
        raise NotImplementedError("the COCO-detection HungarianMatcher is outside the ParSeDA hot path")
    return HungarianMatcherHOI(cost_obj_class=args.set_cost_obj_class, cost_verb_class=args.set_cost_verb_class,
                               cost_bbox=args.set_cost_bbox, cost_giou=args.set_cost_giou,
                               subject_class=args.subject_class)
