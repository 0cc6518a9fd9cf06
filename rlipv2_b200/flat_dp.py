"""Flat parameter / gradient buffers for the data-parallel step.

The reference wraps the model in DistributedDataParallel (main.py:515-517) and steps a torch AdamW over
three learning-rate groups (main.py:523-539).  Here the parameters of each group become views of one flat
buffer, so that one step needs exactly one all-reduce (NCCL on the GPUs, gloo in the CPU tests), one norm for
`clip_grad_norm_` (engine.py:165-166) and one optimizer launch per group, regardless of the number of tensors.

Pure torch + torch.distributed: the same code runs under NCCL on the GPUs and under gloo on CPU
(tests/test_flat_dp_gloo.py).  The optimizer kernel itself (csrc/fused_ops.cu) is CUDA only and lives in
train_step.py.
"""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class FlatParams:
    """groups: [(list of parameters, lr), ...].  After construction every parameter's `.data` is a view of
    `flat_param` and (unless `grad_views=False`) its `.grad` a view of `flat_grad`; each group occupies one
    contiguous, 16-byte aligned range `group_ranges[i] = (start, end, lr)`."""

    def __init__(self, groups, device, grad_views=True, moments=True):
        ranges, off = [], 0
        for plist, lr in groups:
            start = off
            off += sum(p.numel() for p in plist)
            ranges.append((start, off, lr))
            off = (off + 3) // 4 * 4
        self.flat_param = torch.zeros(off, device=device)
        self.flat_grad = torch.zeros(off, device=device)
        self.exp_avg = torch.zeros(off, device=device) if moments else None
        self.exp_avg_sq = torch.zeros(off, device=device) if moments else None
        self.param_offsets = []
        for (plist, _), (start, _, _) in zip(groups, ranges):
            o = start
            for p in plist:
                n = p.numel()
                self.flat_param[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.flat_param[o:o + n].view_as(p)
                p.grad = self.flat_grad[o:o + n].view_as(p) if grad_views else None
                self.param_offsets.append(o)
                o += n
        self.group_ranges = ranges
        self.params = [p for plist, _ in groups for p in plist]

    def broadcast_params(self, src=0):
        """identical replicas, made certain (DDP does the same at construction)"""
        if world_size() > 1:
            dist.broadcast(self.flat_param, src)

    def zero_grad(self):
        self.flat_grad.zero_()

    def allreduce_mean_(self):
        """what DDP's bucketed all-reduce amounts to: the rank-mean of the gradients, in one collective"""
        w = world_size()
        if w > 1:
            dist.all_reduce(self.flat_grad)
            self.flat_grad.div_(w)

    def allreduce_sum_(self):
        """the rank SUM of the gradients; `clip_scale` folds the 1 / world of the mean into the optimizer's read"""
        if world_size() > 1:
            dist.all_reduce(self.flat_grad)

    def clip_scale(self, max_norm):
        """1-element tensor s such that `flat_grad * s` is what `allreduce_mean_()` + `clip_(max_norm)` would have left
        in the buffer, given that `flat_grad` holds the rank SUM: s = clip coefficient of the mean gradient / world.
        Nothing is written to the gradient buffer (the optimizer kernel applies s as it reads the gradient)."""
        w = float(world_size())
        if max_norm > 0:
            coef = torch.clamp(max_norm / (self.flat_grad.norm() / w + 1e-6), max=1.0)
            return (coef / w).reshape(1)
        return torch.full((1,), 1.0 / w, device=self.flat_grad.device)

    def clip_(self, max_norm):
        """`clip_grad_norm_(parameters, max_norm)` over the used parameters == one norm + one scale of the flat
        buffer (padding elements are zero); no host sync."""
        if max_norm > 0:
            coef = torch.clamp(max_norm / (self.flat_grad.norm() + 1e-6), max=1.0)
            self.flat_grad.mul_(coef)
