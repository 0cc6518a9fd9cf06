"""Small tensor utilities the hot path calls (semantics of /root/reference/util/misc.py:299-340,460-464
and util/box_ops.py:19-73, re-implemented)."""
from typing import List, Optional

import torch
from torch import Tensor


class NestedTensor:
    """Batched images padded to a common size + padding mask (True = padded pixel).
    Same attribute names as the reference's NestedTensor (util/misc.py:320-340)."""

    def __init__(self, tensors: Tensor, mask: Optional[Tensor]):
        self.tensors = tensors
        self.mask = mask

    def to(self, device, non_blocking=False):
        mask = self.mask.to(device, non_blocking=non_blocking) if self.mask is not None else None
        return NestedTensor(self.tensors.to(device, non_blocking=non_blocking), mask)

    def decompose(self):
        return self.tensors, self.mask

    def __repr__(self):
        return f"NestedTensor({tuple(self.tensors.shape)})"


def nested_tensor_from_tensor_list(tensor_list: List[Tensor]) -> NestedTensor:
    """Pad a list of [3, H_i, W_i] images to the largest H/W, top-left aligned (util/misc.py:299-317)."""
    if tensor_list[0].ndim != 3:
        raise ValueError("not supported")
    c = tensor_list[0].shape[0]
    H = max(int(t.shape[1]) for t in tensor_list)
    W = max(int(t.shape[2]) for t in tensor_list)
    b = len(tensor_list)
    out = tensor_list[0].new_zeros((b, c, H, W))
    mask = torch.ones((b, H, W), dtype=torch.bool, device=tensor_list[0].device)
    for i, img in enumerate(tensor_list):
        out[i, :, : img.shape[1], : img.shape[2]].copy_(img)
        mask[i, : img.shape[1], : img.shape[2]] = False
    return NestedTensor(out, mask)


def inverse_sigmoid(x: Tensor, eps: float = 1e-5) -> Tensor:
    """logit with both operands clamped to eps (util/misc.py:460-464)."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def box_cxcywh_to_xyxy(x: Tensor) -> Tensor:
    cx, cy, w, h = x.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def box_xyxy_to_cxcywh(x: Tensor) -> Tensor:
    x0, y0, x1, y1 = x.unbind(-1)
    return torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, x1 - x0, y1 - y0], dim=-1)


def box_iou(boxes1: Tensor, boxes2: Tensor):
    """Pairwise IoU and union of xyxy boxes: [N,4] x [M,4] -> [N,M] (util/box_ops.py:34-47)."""
    area1 = (boxes1[:, 2] - boxes1[:, 0]) * (boxes1[:, 3] - boxes1[:, 1])
    area2 = (boxes2[:, 2] - boxes2[:, 0]) * (boxes2[:, 3] - boxes2[:, 1])
    lt = torch.max(boxes1[:, None, :2], boxes2[:, :2])
    rb = torch.min(boxes1[:, None, 2:], boxes2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    union = area1[:, None] + area2 - inter
    return inter / union, union


def generalized_box_iou(boxes1: Tensor, boxes2: Tensor, check: bool = True) -> Tensor:
    """Pairwise GIoU of xyxy boxes (util/box_ops.py:50-73).  `check=False` skips the degenerate-box
    asserts, which cost a device->host sync each."""
    if check:
        assert (boxes1[:, 2:] >= boxes1[:, :2]).all()
        assert (boxes2[:, 2:] >= boxes2[:, :2]).all()
    iou, union = box_iou(boxes1, boxes2)
    lt = torch.min(boxes1[:, None, :2], boxes2[:, :2])
    rb = torch.max(boxes1[:, None, 2:], boxes2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    area = wh[:, :, 0] * wh[:, :, 1]
    return iou - (area - union) / area
