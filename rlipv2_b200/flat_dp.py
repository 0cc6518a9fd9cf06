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

from . import streams


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
        # equal, 16-byte aligned shards per rank (sharded optimizer step: reduce-scatter / all-gather in place)
        unit = 4 * max(1, world_size())
        off = (off + unit - 1) // unit * unit
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

    # ---- sharded optimizer step (ZeRO-1 style): every rank reduces, clips and updates 1 / world of the flat buffer --------
    def shard_range(self):
        w, n = world_size(), self.flat_grad.numel()
        r = dist.get_rank() if w > 1 else 0
        per = n // w
        return r * per, (r + 1) * per

    def reduce_scatter_sum_(self):
        """this rank's shard of `flat_grad` <- the rank SUM of that shard (in place; the other shards keep local values).
        Half the bytes of the all-reduce; the other half is the all-gather of the updated parameters."""
        if world_size() <= 1:
            return
        s0, s1 = self.shard_range()
        if dist.get_backend() == "nccl":
            dist.reduce_scatter_tensor(self.flat_grad[s0:s1], self.flat_grad)
        else:                                   # gloo (CPU tests) has no reduce-scatter: same result through an all-reduce
            dist.all_reduce(self.flat_grad)

    def all_gather_params_(self):
        """every rank's updated parameter shard -> all ranks (in place): replicas leave the step bit-identical"""
        if world_size() <= 1:
            return
        s0, s1 = self.shard_range()
        if dist.get_backend() == "nccl":
            dist.all_gather_into_tensor(self.flat_param, self.flat_param[s0:s1])
        else:
            per = s1 - s0
            parts = [torch.empty(per, dtype=self.flat_param.dtype) for _ in range(world_size())]
            dist.all_gather(parts, self.flat_param[s0:s1].clone())
            self.flat_param.copy_(torch.cat(parts))

    def clip_scale_sharded(self, max_norm):
        """`clip_scale` when only this rank's shard holds the rank sum: squared norms of the shards are summed across ranks
        (one scalar all-reduce) - the same global norm, hence the same coefficient on every rank"""
        w = float(world_size())
        s0, s1 = self.shard_range()
        if max_norm > 0:
            sq = self.flat_grad[s0:s1].square().sum().reshape(1)
            if w > 1:
                dist.all_reduce(sq)
            coef = torch.clamp(max_norm / (sq.sqrt() / w + 1e-6), max=1.0)
            return (coef / w).reshape(1)
        return torch.full((1,), 1.0 / w, device=self.flat_grad.device)

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


class EarlyReducer:
    """All-reduce ranges of the flat gradient buffer as soon as `grad_ready` markers say they are final, on a
    communication stream beside the rest of the backward - the flat-buffer counterpart of DDP's bucket hooks
    (main.py:515-517).  Sum semantics (like `allreduce_sum_`): `clip_scale` folds the 1 / world into the optimizer.

    entries: [(tags, start, end)] - launch the all-reduce of flat_grad[start:end] once every tag of `tags` has fired.
    Stream order: a marker fires on the stream its activation was produced on; an event is recorded there at that moment
    and the launch waits for the events of all the entry's tags - NOT for the streams' later tails (when the last tag
    fires the host may already have queued the backbone's backward on the main stream; waiting for that would undo
    the overlap).  wait_streams: callable -> the SIDE streams that may hold gradient writes of the ranges (parameter-
    gradient stream of dense.py, label / value-projection / criterion streams); they carry nothing but such work, so the
    launch waits for their tails.  While a CUDA graph is being captured, side streams that are not part of the capture
    are skipped (nothing of this backward can be on them).
    Every rank must see the same entries and the same firing order (same model, same autograd graph: it does).

        reducer.begin()            # before loss.backward(); markers call reducer.on_tag(tag) from the backward
        loss.backward()
        reducer.finish()           # all-reduce whatever was not launched early, then join the communication stream
    """

    def __init__(self, flat, entries, wait_streams=None, force=False):
        self.flat = flat
        self.entries = [(frozenset(t), int(s), int(e)) for t, s, e in entries if e > s]
        for _, s, e in self.entries:
            assert 0 <= s < e <= flat.flat_grad.numel()
        ordered = sorted((s, e) for _, s, e in self.entries)
        assert all(a[1] <= b[0] for a, b in zip(ordered, ordered[1:])), "early-reduce ranges overlap"
        self.wait_streams = wait_streams or (lambda: [])
        self.active = force or world_size() > 1
        self.cuda = flat.flat_grad.is_cuda
        self.comm = streams.get(flat.flat_grad.device, "comm") if self.cuda else None
        self.fired, self.launched, self.events = set(), [], {}

    def begin(self):
        self.fired, self.launched, self.events = set(), [], {}

    def _reduce(self, start, end):
        if world_size() > 1:
            dist.all_reduce(self.flat.flat_grad[start:end])

    def on_tag(self, tag):
        if not self.active:
            return
        if tag in self.fired:
            return
        self.fired.add(tag)
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.flat.flat_grad.device))
            self.events[tag] = ev
        for tags, start, end in self.entries:
            if (start, end) in self.launched or not tags <= self.fired:
                continue
            self.launched.append((start, end))
            if self.cuda:
                for t in tags:
                    self.comm.wait_event(self.events[t])
                capturing = torch.cuda.is_current_stream_capturing()
                for s in self.wait_streams():
                    if capturing:
                        with torch.cuda.stream(s):
                            if not torch.cuda.is_current_stream_capturing():
                                continue
                    self.comm.wait_stream(s)
                with torch.cuda.stream(self.comm):
                    self._reduce(start, end)
            else:
                self._reduce(start, end)

    def remaining(self):
        """the parts of the buffer no early launch covered, as maximal ranges"""
        out, pos = [], 0
        for s, e in sorted(self.launched):
            if s > pos:
                out.append((pos, s))
            pos = max(pos, e)
        n = self.flat.flat_grad.numel()
        if pos < n:
            out.append((pos, n))
        return out

    def finish(self):
        if not self.active:
            return
        for s, e in self.remaining():
            self._reduce(s, e)
        if self.cuda and self.launched:                 # (an unused communication stream is not part of a capture)
            torch.cuda.current_stream(self.flat.flat_grad.device).wait_stream(self.comm)
