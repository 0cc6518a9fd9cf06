"""One RLIPv2-ParSeDA training step as the reference's engine runs it.

Restates the body of the hot loop of /root/reference/engine.py:68-172 (`train_one_epoch`) and the
optimizer construction of main.py:514-539 as a reusable object:

    H2D of the batch -> model(samples, encode_and_save=True) -> model(..., encode_and_save=False)
    -> criterion -> weighted sum -> zero_grad -> backward -> clip_grad_norm_(0.1) -> AdamW.step

Data-parallel: one process per GPU, the model wrapped in DistributedDataParallel over NCCL
(main.py:515-517).  The reference needs `find_unused_parameters=True` because the verb decoder's
box heads only feed detached anchors (dab_deformable/deformable_transformer.py:1511-1541); the set of
unused parameters is static, so `static_graph=True` gives the same result without the per-step graph
walk.
"""
import os

import torch
import torch.distributed as dist

from . import dense, models, streams
from .nested import NestedTensor


from .synth import synthetic_batch, synthetic_text  # noqa: E402,F401  (kept importable from here)


TEXT_TOWER_SPLIT = 6          # text-tower layers [split, 12) + pooler are reduced when the backward passes layer `split`


def early_reduce_entries(names, offsets, group_ranges, image_tags=None):
    """[(tags, start, end)] for flat_dp.EarlyReducer from the flat layout of the graphed step (groups in the order of
    main.py:525-537: everything else | backbone | text encoder; parameters in named_parameters order inside a group).
    image_tags: the 'image<i>' markers the forward applies (default: three backbone levels)."""
    tags0 = {t for t in (image_tags or {"image0", "image1", "image2"}) if t.startswith("image")} | {"text"}
    entries = [(tags0, group_ranges[0][0], group_ranges[0][1])]

    def first(prefix):
        for n, o in zip(names, offsets):
            if "text_encoder" in n and prefix in n:
                return o
        return None

    lo, mid, end = first("encoder.layer.0."), first(f"encoder.layer.{TEXT_TOWER_SPLIT}."), group_ranges[2][1]
    if lo is not None and mid is not None and group_ranges[2][0] <= lo < mid < end:
        entries.append(({"text_mid"}, mid, end))            # layers split.. and the pooler (named after the layers)
        entries.append(({"text_emb"}, lo, mid))             # layers 0 .. split-1
    return entries


class ParSeDATrainStep:
    """model + criterion + optimizer; `step(images_host, targets_host, text)` runs one iteration and
    returns the (device) total loss."""

    def __init__(self, args=None, device="cuda", precision="tf32", seed=0, ddp=None, clip_max_norm=0.1,
                 lr=1.41e-4, lr_backbone=1.41e-5, text_encoder_lr=1.41e-5, weight_decay=1e-4):
        self.device = torch.device(device)
        dense.set_matmul_precision(precision)
        torch.manual_seed(seed)
        if args is None:
            args = models.default_args(device=str(self.device), num_queries=300, synthetic_text_encoder=True)
        self.args = args
        model, criterion, _ = models.build_model(args)
        self.model = model.to(self.device).train()
        self.criterion = criterion.to(self.device).train()
        self.module = self.model
        if (os.environ.get("RLIPV2_SHARD_LABELS", "0") == "1" and dist.is_available() and dist.is_initialized()
                and dist.get_world_size() > 1 and hasattr(self.module.transformer, "shard_label_text")):
            # opt-in, for runs whose label text is THE SAME on every rank (fine-tuning on a fixed vocabulary): each rank
            # runs the text tower on 1 / world of the strings (text_encoder.pooled_text_sharded; checked on two gloo ranks,
            # not yet run under NCCL inside the captured graphs)
            self.module.transformer.shard_label_text(True)
        # NB: channels_last (NHWC) for the backbone was measured and dropped: it removes cuDNN's
        # nchwToNhwc/nhwcToNchw pairs (1.7 ms) but the frozen-BN / GroupNorm elementwise kernels get
        # slower by more than that (52.1 vs 50.3 ms per step).
        if ddp is None:
            ddp = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if ddp:
            self.model = torch.nn.parallel.DistributedDataParallel(
                self.model, device_ids=[self.device.index], static_graph=True, gradient_as_bucket_view=True)
        # main.py:523-539: three groups selected by name
        named = list(self.module.named_parameters())
        groups = [
            {"params": [p for n, p in named if "backbone" not in n and "text_encoder" not in n and p.requires_grad]},
            {"params": [p for n, p in named if "backbone" in n and p.requires_grad], "lr": lr_backbone},
            {"params": [p for n, p in named if "text_encoder" in n and p.requires_grad], "lr": text_encoder_lr},
        ]
        self.optimizer = torch.optim.AdamW(groups, lr=lr, weight_decay=weight_decay, fused=self.device.type == "cuda")
        self.clip_max_norm = clip_max_norm
        self.params = [p for g in groups for p in g["params"]]

    def to_device(self, images_host, targets_host):
        images = images_host.to(self.device, non_blocking=True)
        mask = torch.zeros(images.shape[0], images.shape[2], images.shape[3], dtype=torch.bool, device=self.device)
        targets = [{k: v.to(self.device, non_blocking=True) for k, v in t.items()} for t in targets_host]
        return NestedTensor(images, mask), targets

    def step_device(self, samples, targets, text):
        """One optimisation step on a batch that is already resident in HBM."""
        memory_cache = self.model(samples, encode_and_save=True, text=text, targets=targets)
        outputs = self.model(samples, encode_and_save=False, memory_cache=memory_cache, text=text, targets=targets)
        losses = self._weighted_total(self.criterion(outputs, targets))
        self.optimizer.zero_grad(set_to_none=True)
        losses.backward()
        if self.clip_max_norm > 0:
            torch.nn.utils.clip_grad_norm_(self.params, self.clip_max_norm, foreach=True)
        self.optimizer.step()
        return losses.detach()

    def step(self, images_host, targets_host, text):
        samples, targets = self.to_device(images_host, targets_host)
        return self.step_device(samples, targets, text)

    def _weighted_total(self, loss_dict):
        """engine.py:108 `sum(loss_dict[k] * weight_dict[k] ...)`; the criterion already formed it on the
        device as one dot product when it evaluated all decoder layers in one pass."""
        total = getattr(loss_dict, "weighted_total", None)
        if total is not None:
            return total
        wd = self.criterion.weight_dict
        return sum(loss_dict[k] * wd[k] for k in loss_dict.keys() if k in wd)


class GraphedParSeDATrainStep(ParSeDATrainStep):
    """The same optimisation step replayed from two CUDA graphs (the reference's step is launch-bound:
    ~7000 kernels and ~100 ms of host time per step at batch 2, engine.py:99-172).

        graph A   forward (phase A + phase B) + the matcher's cost tensors for the 3 decoder layers,
                  copied to pinned host memory
        host      scipy linear_sum_assignment per layer and image (models/matcher.py:193) - the only
                  host work left in the step; indices go back into static device buffers
        graph B   SetCriterionHOI from the static indices -> backward -> (NCCL all-reduce of the flat
                  gradient buffer when world > 1) -> clip_grad_norm_(0.1) -> fused AdamW

    Shapes are static: image size, label count and the number of target triplets per image are fixed at
    capture time (re-capture for a new shape bucket).  Gradients end up in one flat buffer (autograd's
    freshly produced gradient tensors are packed by one gather kernel; `RLIPV2_GATHER_GRADS=0` keeps the older
    scheme of `.grad` views accumulated into a zeroed buffer), so data-parallel training needs exactly one
    all-reduce per step and no DDP wrapper; parameters that
    never receive a gradient (the verb decoder's detached box heads) are left out of the optimizer,
    which is what the reference's `find_unused_parameters=True` + grad-is-None skip amounts to.
    """

    def __init__(self, *a, **kw):
        kw["ddp"] = False
        super().__init__(*a, **kw)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.captured = False
        self._prefetched = False
        # measured neutral on 1xB200 (37.2 vs 37.0 ms/step): inside a graph the ~600 accumulation kernels cost about what
        # the gather + the copies of non-stealable gradients cost.  Kept behind the switch.
        self.gather_grads = os.environ.get("RLIPV2_GATHER_GRADS", "0") == "1"
        # backward graph launched behind the forward graph, parked on a host flag until the assignment is solved
        # (hides the multi-millisecond launch of the ~3000-node graph; RLIPV2_FLAG_WAIT=0: launch it after the solve)
        self.flag_wait = os.environ.get("RLIPV2_FLAG_WAIT", "1") != "0"
        # clip coefficient and the 1 / world of the rank mean applied inside the AdamW kernel's gradient read instead of
        # two passes over the 850 MB gradient buffer (the torch scalar-broadcast multiply alone took 0.54 ms)
        self.fused_clip = os.environ.get("RLIPV2_FUSED_CLIP", "1") != "0"
        # gradient buffer zeroed at the head of the forward graph (where the GPU idles while the graph starts up)
        # instead of at the head of the backward graph (on the critical path behind the assignment)
        self.early_zero = os.environ.get("RLIPV2_EARLY_ZERO", "1") != "0"
        self.stamps = None                      # diagnostic globaltimer stamps (RLIPV2_STAMPS=1, tools/step_anatomy.py)
        # assignment problems solved on the device by rlipv2_lsap_f32 (bit-identical to scipy, tests/test_lsap_core.py):
        # no cost D2H, no host solve, no index H2D, no flag wait - the two graphs replay back to back without the host.
        # Opt-in until it has its A/B on the B200 (written after round 1's GPU budget was spent).
        self.device_lsap = os.environ.get("RLIPV2_DEVICE_LSAP", "0") == "1"
        # world > 1: ranges of the flat gradient buffer all-reduced on a communication stream as soon as grad_ready
        # markers say they are final (transformer + heads when the backward reaches the backbone features; text-tower
        # layers as their half of the tower finishes), instead of one all-reduce behind the whole backward.  Opt-in until
        # it has its 2- and 8-GPU A/B (written after round 1's GPU budget was spent).
        # ("force": install the markers / events / communication stream at world size 1 too - plumbing test on one GPU)
        self.overlap_allreduce = os.environ.get("RLIPV2_ALLREDUCE_OVERLAP", "0") in ("1", "force") and not self.gather_grads
        self.overlap_force = os.environ.get("RLIPV2_ALLREDUCE_OVERLAP", "0") == "force"
        self.reducer = None
        # world > 1: sharded optimizer step - reduce-scatter the flat gradient (half the all-reduce's bytes), clip and AdamW on
        # this rank's 1 / world of the buffer (the 1 ms three-launch AdamW over 213 M parameters shrinks with the rank count),
        # all-gather the updated parameters.  Same arithmetic per element, replicas stay bit-identical.
        self.shard_optimizer = os.environ.get("RLIPV2_SHARD_OPTIMIZER", "1") != "0" and not self.gather_grads

    # the piece of work each graph records -------------------------------------------------------------
    def _stamp(self, i):
        if self.stamps is not None:
            from . import fused_abi
            fused_abi.stamp(self.stamps, i)

    def _forward_and_costs(self):
        self._stamp(0)
        zero_side = None
        if self.early_zero and not self.gather_grads and getattr(self, "flat_grad", None) is not None:
            # 850 MB memset beside the (launch-bound) start of the text tower / backbone stem, joined below
            cur = torch.cuda.current_stream(self.device)
            if getattr(self, "_zero_stream", None) is None:
                self._zero_stream = streams.get(self.device, "zero")
            zero_side = self._zero_stream
            zero_side.wait_stream(cur)
            with torch.cuda.stream(zero_side):
                self.flat_grad.zero_()
        cache = self.module(self.s_samples, encode_and_save=True, text=self.s_tok, targets=self.s_targets)
        outputs = self.module(self.s_samples, encode_and_save=False, memory_cache=cache, text=self.s_tok,
                              targets=self.s_targets)
        if zero_side is not None:
            torch.cuda.current_stream(self.device).wait_stream(zero_side)
        layers = self.criterion.layers_of(outputs)
        C, cost_lists = self.criterion.matcher.compute_costs_layers(layers, self.s_targets)   # all layers, one pass
        if self.device_lsap:
            from . import lsap_abi
            lsap_abi.solve(C, self.lsap_plan, self.s_I, self.s_J)
            self.last_cost = C                  # (static graph memory: readable after a replay, e.g. by the tests)
        else:
            self.h_cost.copy_(C, non_blocking=True)
        self._stamp(1)
        giou = -torch.stack([cl[0] for cl in cost_lists]) if self.criterion.giou_verb_label else None
        return outputs, giou

    def _loss_backward_step(self, outputs, giou, graph_head=False):
        from .criterion import StackedMatches
        self._stamp(2)
        if graph_head and not self.device_lsap:
            # captured as the first nodes of graph B: wait for the host's publication of this replay, then fetch
            # the matched indices from the pinned buffers (memcpy nodes with fixed addresses)
            from . import fused_abi
            fused_abi.wait_host_flag(self.h_flag, self.d_seq, self.d_err, timeout_s=self.flag_timeout_s)
            self.s_I.copy_(self.h_I, non_blocking=True)
            self.s_J.copy_(self.h_J, non_blocking=True)
        self._stamp(3)
        matches = StackedMatches(self.s_I, self.s_J, giou, self.ks)
        total = self._weighted_total(self.criterion(outputs, self.s_targets, matches=matches))
        if self.gather_grads:
            # autograd hands every parameter a freshly produced gradient tensor (no zero-fill, no `+=`);
            # one gather kernel then packs them into the flat buffer (csrc/fused_ops.cu)
            for p in self.params:
                p.grad = None
            total.backward()
            dense.join_param_grad_stream()
            self._gather_grads()
        else:
            if not self.early_zero:
                self.flat_grad.zero_()
            if self.reducer is not None:
                self.reducer.begin()
            total.backward()
        dense.join_param_grad_stream()               # parameter gradients formed on the side stream (dense.py)
        if self.reducer is not None and self.fused_clip:
            self.reducer.finish()                    # whatever the markers did not launch early + join the comm stream
            self._adamw_step(self.flat.clip_scale(self.clip_max_norm))
        elif self.fused_clip and self.shard_optimizer and self.world > 1:
            self.flat.reduce_scatter_sum_()          # this rank's shard of the gradient sum
            self._adamw_step(self.flat.clip_scale_sharded(self.clip_max_norm), shard=self.flat.shard_range())
            self.flat.all_gather_params_()
        elif self.fused_clip:
            self.flat.allreduce_sum_()               # one NCCL all-reduce of the flat buffer (world > 1)
            self._adamw_step(self.flat.clip_scale(self.clip_max_norm))   # clip_grad_norm_ = one norm; scale in AdamW
        else:
            self.flat.allreduce_mean_()
            self.flat.clip_(self.clip_max_norm)      # clip_grad_norm_: one norm + one scale
            self._adamw_step()
        self._stamp(4)
        return total.detach()

    def _gather_grads(self):
        from . import fused_abi
        grads, offs = [], []
        for p, off in zip(self.params, self.param_offsets):
            g = p.grad
            assert g is not None, "a parameter that received a gradient in the probe step has none now"
            grads.append(g if g.is_contiguous() else g.contiguous())
            offs.append(off)
        _, n = fused_abi.gather_table(grads, offs, out=self.h_gather)
        self.d_gather.copy_(self.h_gather, non_blocking=True)      # pinned -> device; a memcpy node when captured
        fused_abi.gather_chunks(self.d_gather, n, self.flat_grad)
        for p in self.params:                                       # the buffers are free for reuse from here on
            p.grad = None

    def _install_early_reducer(self, names, force=False):
        """names: parameter names in the order of self.params / self.param_offsets (group by group)"""
        from . import grad_ready
        from .flat_dp import EarlyReducer
        entries = early_reduce_entries(names, self.param_offsets, self.group_ranges)
        self.reducer = EarlyReducer(self.flat, entries, wait_streams=self._gradient_side_streams, force=force)
        grad_ready.set_callback(self.reducer.on_tag)
        self._marker_hooks = grad_ready.install_text_tower_markers(self.module.transformer.text_encoder, TEXT_TOWER_SPLIT)
        return entries

    def _gradient_side_streams(self):
        """side streams (never the main one) that backward nodes writing parameter gradients may have been queued on"""
        from .criterion import _BRANCH_STREAMS
        streams = list(dense.pending_param_grad_streams())
        for m in self.module.modules():
            for attr in ("_lang_stream", "_value_stream"):
                st = getattr(m, attr, None)
                if st is not None:
                    streams.append(st)
        streams += _BRANCH_STREAMS.get(str(self.device), [])
        return streams

    def uninstall_early_reducer(self):
        from . import grad_ready
        if self.reducer is not None:
            grad_ready.set_callback(None)
            for h in self._marker_hooks:
                h.remove()
            self.reducer = None

    def _adamw_step(self, grad_scale=None, shard=None):
        """shard = (s0, s1): update only that range of the flat buffers (this rank's part of a sharded step)"""
        from . import fused_abi
        self.step_t.add_(1.0)
        skip = self.d_err                       # non-zero: host-flag timeout, or the dry warm-up of a later shape's capture
        for gi, (start, end, lr) in enumerate(self.group_ranges):
            if shard is not None:
                start, end = max(start, shard[0]), min(end, shard[1])
                if start >= end:
                    continue
            # lr comes from the device vector (set_lr): a host scalar would be frozen into the captured graph
            fused_abi.adamw(self.flat_param[start:end], self.flat_grad[start:end], self.exp_avg[start:end],
                            self.exp_avg_sq[start:end], lr, 0.9, 0.999, 1e-8, self.weight_decay, self.step_t,
                            grad_scale=grad_scale, lr_dev=self.lr_dev[gi:gi + 1], skip_flag=skip)

    def _solve_assignment(self):
        """host: LSAP on the pinned cost tensor [layers, bs, nq, T] -> the static device index buffers
        (stacked layout of criterion.StackedMatches: ordered layer, image, match), one H2D copy each"""
        if self.device_lsap:
            return                              # _forward_and_costs already left the indices in s_I / s_J
        o = 0
        for li in range(self.h_cost.shape[0]):
            for i, j in self.criterion.matcher.solve(self.h_cost[li], self.sizes):
                k = i.shape[0]
                self.h_I[o:o + k].copy_(i)
                self.h_J[o:o + k].copy_(j)
                o += k
        self.s_I.copy_(self.h_I, non_blocking=True)
        self.s_J.copy_(self.h_J, non_blocking=True)

    def _solve_assignment_host(self):
        """the same solve, written straight into the pinned index buffers through numpy views (no torch ops, no
        H2D: graph B copies them after its flag wait)"""
        from scipy.optimize import linear_sum_assignment
        cost, hi, hj = self.np_cost, self.np_I, self.np_J
        o = 0
        for li in range(cost.shape[0]):
            t0 = 0
            for b, n in enumerate(self.sizes):
                i, j = linear_sum_assignment(cost[li, b, :, t0:t0 + n])      # models/matcher.py:193
                k = i.shape[0]
                hi[o:o + k] = i
                hj[o:o + k] = j
                o += k
                t0 += n

    def check(self):
        """raise if a replay's flag wait timed out (the host never published its assignment)"""
        if self.device_lsap:
            from . import lsap_abi
            lsap_abi.check(self.lsap_plan)
            return
        if self.captured and self.flag_wait and int(self.d_err.item()) > 0:
            raise RuntimeError(f"backward graph replay {int(self.d_err.item())} timed out waiting for the host assignment")

    # attributes that belong to ONE captured shape (image size, triplet counts, label-token shape); `capture` of another
    # shape replaces them, `activate` swaps a saved set back in (BucketedParSeDATrainStep)
    SHAPE_STATE = ("s_tok", "s_samples", "s_targets", "sizes", "h_cost", "ks", "h_I", "h_J", "s_I", "s_J", "np_cost", "np_I",
                   "np_J", "lsap_plan", "h_flag", "np_flag", "d_seq", "flag_seq", "graph_a", "graph_b", "s_loss", "_keep",
                   "done_a", "own_launches_per_step", "h_ids", "h_am", "_tok_copied", "_copy_stream", "p_images", "p_targets", "p_tok",
                   "_p_has_text",
                   "_staging_free", "_prefetch_done", "_prefetched", "last_cost")

    def shape_state(self):
        return {k: getattr(self, k) for k in self.SHAPE_STATE if hasattr(self, k)}

    def activate(self, state):
        for k in self.SHAPE_STATE:
            if k in state:
                setattr(self, k, state[k])
            elif hasattr(self, k):
                delattr(self, k)
        self._prefetched = bool(state.get("_prefetched", False))

    def capture(self, images_host, targets_host, text, warmup=3, graphs=True):
        """Static buffers + (graphs=True) the two CUDA graphs for this batch shape.  The first call also runs the probe step
        and moves the parameters into the flat buffers; later calls (other shapes) leave the training state untouched: their
        warm-up and priming steps run with the optimizer update disabled.  graphs=False: buffers only - `replay()` then runs
        the same three pieces eagerly (rare shapes that are not worth a capture)."""
        dev = self.device
        first = getattr(self, "flat", None) is None
        for k in ("h_ids", "h_am", "_tok_copied", "_copy_stream", "p_images", "p_targets", "p_tok", "_p_has_text", "_staging_free", "_prefetch_done",
                  "graph_a", "graph_b", "lsap_plan", "last_cost"):
            if hasattr(self, k):
                delattr(self, k)                # staging / graphs of the previously active shape
        self._prefetched = False
        if isinstance(text, dict):              # already tokenised (host or device tensors)
            self.s_tok = {"input_ids": text["input_ids"].to(dev).clone(), "attention_mask": text["attention_mask"].to(dev).clone(),
                          "sums": [tuple(x) for x in text["sums"]]}
        else:
            self.s_tok = self.module.transformer.tokenize(text, dev)
        self.s_samples, self.s_targets = self.to_device(images_host, targets_host)
        self.sizes = [len(t["obj_labels"]) for t in targets_host]
        nq = self.args.num_queries // 2
        n_layers = self.args.dec_layers
        self.h_cost = torch.empty(n_layers, len(self.sizes), nq, sum(self.sizes)).pin_memory()
        self.ks = [min(nq, n) for n in self.sizes]                 # matches per image (LSAP on nq x n)
        K = n_layers * sum(self.ks)
        self.h_I, self.h_J = torch.zeros(K, dtype=torch.long).pin_memory(), torch.zeros(K, dtype=torch.long).pin_memory()
        self.s_I, self.s_J = torch.zeros(K, dtype=torch.long, device=dev), torch.zeros(K, dtype=torch.long, device=dev)
        self.np_cost, self.np_I, self.np_J = self.h_cost.numpy(), self.h_I.numpy(), self.h_J.numpy()
        if self.device_lsap:
            from . import lsap_abi
            self.lsap_plan = lsap_abi.Plan(self.sizes, nq, n_layers, dev)
            assert self.lsap_plan.K == K and self.lsap_plan.ks == self.ks
            self.flag_wait = False              # nothing for graph B to wait for
        self.h_flag = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.np_flag = self.h_flag.numpy()
        self.d_seq = torch.zeros(1, dtype=torch.int32, device=dev)
        if first:
            # error word of the host-flag wait = "skip the update" word of the AdamW kernel (shared by every shape's graphs)
            self.d_err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.flag_seq = 0
        if os.environ.get("RLIPV2_STAMPS", "0") == "1":
            self.stamps = torch.zeros(8, dtype=torch.int64, device=dev)
        self.flag_timeout_s = float(os.environ.get("RLIPV2_FLAG_TIMEOUT_S", "10"))

        # One stream for the probe, the warm-up and both captures: autograd runs every backward node
        # (AccumulateGrad included) on the stream its forward op first ran on, so all of them must be
        # the capture stream - never the legacy default stream.
        if first:
            self.cap_stream = streams.get(self.device, "capture")
        side = self.cap_stream
        side.wait_stream(torch.cuda.current_stream())
        if first:
            self._setup_flat(side)
            # the flat buffers, the learning-rate vector and the step counter were filled on the CURRENT stream; everything
            # below runs on `side` (a non-blocking stream: no implicit ordering with the default stream).  Without this wait
            # the first warm-up AdamW could read the learning rates before their copy had landed (seen once as NaN costs in a
            # full-suite run, r02h).
            side.wait_stream(torch.cuda.current_stream())
        self.graph_a = self.graph_b = None
        if not graphs:
            return self
        # warm-up / priming steps of a LATER shape must not train: the AdamW kernels see a non-zero skip word and return, the
        # step counter is put back afterwards (gradients are re-zeroed by every step anyway)
        dry = (not first) or getattr(self, "dry_captures", False)
        if dry:
            step_keep = self.step_t.clone()
            self.d_err.fill_(-1)
        self._capture_graphs(side, warmup)
        if dry:
            torch.cuda.synchronize()
            self.d_err.zero_()
            self.step_t.copy_(step_keep)
        self.check()
        return self

    def _setup_flat(self, side):
        dev = self.device
        # 1. one eager step to learn which parameters receive gradients
        for p in self.module.parameters():
            p.grad = None
        with torch.cuda.stream(side):
            outputs, giou = self._forward_and_costs_eager_probe()
        side.synchronize()
        used_ids = {id(p) for p in self.module.parameters() if p.requires_grad and p.grad is not None}
        named = [(n, p) for n, p in self.module.named_parameters() if id(p) in used_ids]
        base = self.optimizer.param_groups
        groups = [                                  # the 3 learning-rate groups of main.py:523-539
            ([p for n, p in named if "backbone" not in n and "text_encoder" not in n], base[0]["lr"]),
            ([p for n, p in named if "backbone" in n], base[1]["lr"]),
            ([p for n, p in named if "text_encoder" in n], base[2]["lr"]),
        ]
        self.weight_decay = base[0]["weight_decay"]
        # 2. flat parameter / gradient / Adam-moment buffers, one contiguous 16-byte aligned range per
        #    group: parameters and their .grad become views, so the step needs one all-reduce, one
        #    norm and three AdamW launches regardless of the number of tensors
        from .flat_dp import FlatParams
        self.flat = FlatParams(groups, dev, grad_views=not self.gather_grads)
        self.flat_param, self.flat_grad = self.flat.flat_param, self.flat.flat_grad
        self.exp_avg, self.exp_avg_sq = self.flat.exp_avg, self.flat.exp_avg_sq
        self.step_t = torch.zeros((), device=dev)
        self.param_offsets = self.flat.param_offsets
        self.group_ranges = self.flat.group_ranges
        # per-group learning rates as device scalars (read by the AdamW kernel): `set_lr` / an lr scheduler keeps working
        # after the capture (a host lr is frozen into the graph, StepLR of main.py:554 would never drop)
        self.lr_dev = torch.tensor([g[2] for g in self.group_ranges], dtype=torch.float32, device=dev)
        self.group_lrs = [float(g[2]) for g in self.group_ranges]
        self.params = self.flat.params
        if not self.gather_grads and os.environ.get("RLIPV2_FUSE_GRAD_ACC", "1") != "0":
            # dense.py's backward functions add weight / bias / LayerNorm gradients straight into these views
            # (GEMM with beta = 1, reductions without the zero-fill) instead of handing them to AccumulateGrad
            for p in self.params:
                p._fuse_grad = True
        if self.gather_grads:
            from . import fused_abi
            rows = sum((p.numel() + fused_abi.GATHER_CHUNK - 1) // fused_abi.GATHER_CHUNK for p in self.params)
            self.h_gather = torch.zeros(rows, 3, dtype=torch.int64).pin_memory()
            self.d_gather = torch.zeros(rows, 3, dtype=torch.int64, device=dev)
        self.optimizer = None                       # replaced by the flat AdamW kernel (fused_ops.cu)
        if self.overlap_allreduce and self.fused_clip and (self.world > 1 or self.overlap_force):
            other = [n for n, _ in named if "backbone" not in n and "text_encoder" not in n]
            self._install_early_reducer(other + [n for n, _ in named if "backbone" in n]
                                        + [n for n, _ in named if "text_encoder" in n], force=self.overlap_force)
        if self.world > 1:                      # identical replicas (same seed), made certain
            dist.broadcast(self.flat_param, 0)
            for p in self.module.parameters():
                if id(p) not in used_ids:
                    dist.broadcast(p.data, 0)

    def _capture_graphs(self, side, warmup):
        # 3. warm-up on a side stream (cuBLAS/cuDNN workspaces, lazy inits), then capture
        # (with the flag wait, the last warm-up step is the priming replay below: `warmup` optimizer steps either way)
        n_eager = max(1, warmup - 1) if self.flag_wait else warmup
        with torch.cuda.stream(side):
            for _ in range(n_eager):
                outputs, giou = self._forward_and_costs()
                side.synchronize()
                self._solve_assignment()
                self._loss_backward_step(outputs, giou)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import attn_abi, dense_abi, fused_abi, lsap_abi, msda_abi
        own = lambda: (msda_abi.launch_count() + dense_abi.launch_count() + fused_abi.launch_count()
                       + lsap_abi.launch_count() + attn_abi.launch_count())
        l0 = own()
        self.graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_a, stream=self.cap_stream):
            outputs, giou = self._forward_and_costs()
        self.graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_b, pool=self.graph_a.pool(), stream=self.cap_stream):
            self.s_loss = self._loss_backward_step(outputs, giou, graph_head=self.flag_wait)
        self.own_launches_per_step = own() - l0   # kernel nodes of this repo's libraries in the two graphs
        self._keep = (outputs, giou)      # the autograd graph's buffers belong to the captured pool
        self.done_a = torch.cuda.Event()
        self.captured = True
        if self.flag_wait:
            # Priming replay.  The first launch of a graph can block the host until the launch has been processed
            # (measured on B200: the whole flag timeout, once) - with the backward graph parked on a flag only the
            # host can publish, that is a deadlock bounded by the timeout.  So the first launch of graph B happens
            # here with the flag already published; every later launch is asynchronous.
            self.graph_a.replay()
            self.done_a.record()
            self.done_a.synchronize()
            self._solve_assignment_host()
            self.flag_seq += 1
            self.np_flag[0] = self.flag_seq
            self.graph_b.replay()
            torch.cuda.synchronize()

    def _forward_and_costs_eager_probe(self):
        outputs, giou = self._forward_and_costs()
        torch.cuda.current_stream().synchronize()
        self._solve_assignment()
        from .criterion import StackedMatches
        matches = StackedMatches(self.s_I, self.s_J, giou, self.ks)
        self._weighted_total(self.criterion(outputs, self.s_targets, matches=matches)).backward()
        return outputs, giou

    def replay(self):
        """One step on the batch currently held by the static buffers."""
        if self.graph_a is None:                # shape without graphs (capture(graphs=False)): the same pieces, eagerly
            cur = torch.cuda.current_stream(self.device)
            self.cap_stream.wait_stream(cur)
            with torch.cuda.stream(self.cap_stream):
                outputs, giou = self._forward_and_costs()
                if not self.device_lsap:
                    self.cap_stream.synchronize()
                    self._solve_assignment()
                loss = self._loss_backward_step(outputs, giou)
            cur.wait_stream(self.cap_stream)
            return loss
        self.graph_a.replay()
        if self.device_lsap:
            self.graph_b.replay()               # the indices are already in s_I / s_J when graph A ends
            return self.s_loss
        self.done_a.record()
        if not self.flag_wait:
            self.done_a.synchronize()           # costs are in pinned memory now
            self._solve_assignment()
            self.graph_b.replay()
            return self.s_loss
        self.graph_b.replay()                   # enqueued behind A; its head waits for the flag below
        self.done_a.synchronize()               # costs are in pinned memory now
        try:
            self._solve_assignment_host()
        finally:
            self.flag_seq += 1                  # always publish: a failed solve must not park the GPU
            self.np_flag[0] = self.flag_seq
        return self.s_loss

    # ---- optimizer surface the reference's main.py touches (lr scheduler, checkpoint save / resume) ----------------
    def set_lr(self, lrs):
        """lrs: one value (scales nothing, sets every group) or one per group in the order of main.py:525-537
        (transformer + heads | backbone | text encoder).  Takes effect at the next replay."""
        lrs = [float(lrs)] * len(self.group_lrs) if not isinstance(lrs, (list, tuple)) else [float(x) for x in lrs]
        assert len(lrs) == len(self.group_lrs)
        self.group_lrs = lrs
        self.lr_dev.copy_(torch.tensor(lrs, dtype=torch.float32), non_blocking=False)

    def step_lr(self, epoch, lr_drop, gamma=0.1, base_lrs=None):
        """torch.optim.lr_scheduler.StepLR(optimizer, lr_drop) of main.py:554, called once per epoch (main.py:724)"""
        base = base_lrs if base_lrs is not None else getattr(self, "_base_lrs", None)
        if base is None:
            base = self._base_lrs = list(self.group_lrs)
        self.set_lr([b * gamma ** (epoch // lr_drop) for b in base])

    def optimizer_state_dict(self):
        """what `optimizer.state_dict()` carries in the reference's checkpoints (main.py:609-611, 746), for the flat AdamW"""
        if self.shard_optimizer and self.world > 1:      # every rank owns the moments of its shard only: gather them
            s0, s1 = self.flat.shard_range()
            for buf in (self.exp_avg, self.exp_avg_sq):
                dist.all_gather_into_tensor(buf, buf[s0:s1].clone())
        return {"exp_avg": self.exp_avg.detach().clone(), "exp_avg_sq": self.exp_avg_sq.detach().clone(),
                "step": float(self.step_t.item()), "lrs": list(self.group_lrs),
                "param_names": [n for n, p in self.module.named_parameters() if any(p is q for q in self.params)]}

    def load_optimizer_state_dict(self, sd):
        assert sd["exp_avg"].numel() == self.exp_avg.numel(), "optimizer state of a different parameter set"
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_t.fill_(float(sd["step"]))
        self.set_lr(sd["lrs"])

    def update_text(self, text):
        """New label strings for the next replay (the reference builds the label set per batch: engine.py:92-98
        `merge_batch_data` / negative sampling, and tokenises it every step, dab_deformable/deformable_transformer.py:497).
        The captured graph holds the token ids in static buffers: same tuple sizes and a token width of at most the captured
        one are copied in (shorter rows padded with the pad id, attention 0 - the tower masks them); anything else needs a
        re-capture and raises."""
        ids, am, sums = self._tokenize_host(text)
        self._fill_token_staging(ids, am, sums)
        s_ids, s_am = self.s_tok["input_ids"], self.s_tok["attention_mask"]
        s_ids.copy_(self.h_ids, non_blocking=True)
        s_am.copy_(self.h_am, non_blocking=True)
        self._tok_copied = torch.cuda.Event()
        self._tok_copied.record()

    def _fill_token_staging(self, ids, am, sums):
        """token ids / attention mask -> the pinned staging pair (padded to the captured width); raises when they do not fit"""
        tr = self.module.transformer
        s_ids, s_am = self.s_tok["input_ids"], self.s_tok["attention_mask"]
        if [tuple(x) for x in sums] != [tuple(x) for x in self.s_tok["sums"]]:
            raise ValueError(f"label-set sizes {sums} differ from the captured {self.s_tok['sums']}: re-capture")
        if ids.shape[0] != s_ids.shape[0] or ids.shape[1] > s_ids.shape[1]:
            raise ValueError(f"token matrix {tuple(ids.shape)} does not fit the captured {tuple(s_ids.shape)}: re-capture")
        if getattr(self, "h_ids", None) is None:
            pad = getattr(getattr(tr, "tokenizer", None), "pad_token_id", None)
            self._pad_id = 1 if pad is None else int(pad)
            self.h_ids = torch.empty(s_ids.shape, dtype=s_ids.dtype).pin_memory()
            self.h_am = torch.empty(s_am.shape, dtype=s_am.dtype).pin_memory()
        if getattr(self, "_tok_copied", None) is not None:
            self._tok_copied.synchronize()           # the previous H2D of these pinned buffers has been consumed
        self.h_ids.fill_(self._pad_id)
        self.h_am.zero_()
        self.h_ids[:, :ids.shape[1]].copy_(ids)
        self.h_am[:, :am.shape[1]].copy_(am)

    def _tokenize_host(self, text):
        if isinstance(text, dict):
            return text["input_ids"], text["attention_mask"], text["sums"]
        sums, flat = [], []
        for obj_names, pred_names in text:
            sums.append((len(obj_names), len(pred_names)))
            flat += list(obj_names) + list(pred_names)
        tok = self.module.transformer.tokenizer.batch_encode_plus(flat, padding="longest", return_tensors="pt")
        return tok["input_ids"], tok["attention_mask"], sums

    def prefetch(self, images_host, targets_host, text=None):
        """Start the host->device copy of the NEXT batch on a copy stream, into staging buffers, while the current step still
        runs (the reference's DataLoader + `.to(device)` pipeline, engine.py:88-90, as a double buffer); `text`: the next
        batch's label strings, tokenised here on the host (while the GPU works) and staged the same way.  The following
        `step()` without a batch consumes it: device-to-device copies into the graphs' static buffers, then the replay."""
        dev = self.device
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = streams.get(dev, "copy")
            self.p_images = torch.empty_like(self.s_samples.tensors)
            self.p_targets = [{k: torch.empty_like(v) for k, v in t.items()} for t in self.s_targets]
            self.p_tok = None
            self._staging_free = None
        self._p_has_text = text is not None
        if text is not None:
            self._fill_token_staging(*self._tokenize_host(text))
            if self.p_tok is None:
                self.p_tok = (torch.empty_like(self.s_tok["input_ids"]), torch.empty_like(self.s_tok["attention_mask"]))
        if self._staging_free is not None:
            self._copy_stream.wait_event(self._staging_free)        # the previous consumer has read the staging buffers
        with torch.cuda.stream(self._copy_stream):
            self.p_images.copy_(images_host, non_blocking=True)
            for st, ht in zip(self.p_targets, targets_host):
                for k in st:
                    st[k].copy_(ht[k], non_blocking=True)
            if text is not None:
                self.p_tok[0].copy_(self.h_ids, non_blocking=True)
                self.p_tok[1].copy_(self.h_am, non_blocking=True)
                self._tok_copied = torch.cuda.Event()
                self._tok_copied.record(self._copy_stream)
            self._prefetch_done = torch.cuda.Event()
            self._prefetch_done.record(self._copy_stream)
        self._prefetched = True

    def step(self, images_host=None, targets_host=None, text=None):
        """H2D of a new batch (same shapes as at capture) + one replayed step.  `text`: the batch's label strings (or a
        `tokenize()` dict); None keeps the label set of the previous step.  Without a batch, the one handed to `prefetch()`
        is used (its copy has been running beside the previous step)."""
        if text is not None:
            self.update_text(text)
        if images_host is None:
            assert getattr(self, "_prefetched", False), "step() without a batch needs a preceding prefetch()"
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._prefetch_done)
            self.s_samples.tensors.copy_(self.p_images, non_blocking=True)
            for st, pt in zip(self.s_targets, self.p_targets):
                for k in st:
                    st[k].copy_(pt[k], non_blocking=True)
            if getattr(self, "_p_has_text", False):
                self.s_tok["input_ids"].copy_(self.p_tok[0], non_blocking=True)
                self.s_tok["attention_mask"].copy_(self.p_tok[1], non_blocking=True)
            self._staging_free = torch.cuda.Event()
            self._staging_free.record(cur)
            self._prefetched = False
            return self.replay()
        self.s_samples.tensors.copy_(images_host, non_blocking=True)
        for st, ht in zip(self.s_targets, targets_host):
            for k in st:
                st[k].copy_(ht[k], non_blocking=True)
        return self.replay()


class BucketedParSeDATrainStep(GraphedParSeDATrainStep):
    """The graphed step for batches whose SHAPE changes from step to step, as the reference's `train_one_epoch` produces
    them (engine.py:68-172: multi-scale resize -> another padded image size per batch, another number of annotated triplets
    per image, engine.py:700-757 `merge_batch_data` -> another label set).  A CUDA graph is bound to one shape, so this
    class keeps a small set of captured shapes:

      key = (padded image shape, triplets per image, label-set sizes, token-matrix width)
      * a shape seen for the `capture_after`-th time is captured (at most `max_shapes` live captures, least recently used
        evicted); the capture's warm-up does not train (GraphedParSeDATrainStep.capture, dry mode);
      * any other shape runs the SAME step eagerly on static buffers of its own (no padding to a bucket: ALIF lets padded
        image cells take part in its softmaxes - SURVEY quirk 1 - so a larger padded size would change the result).
    Parameters, gradients and AdamW state are the one flat buffer set of the first capture; every shape's graphs update it.
    """

    def __init__(self, *a, max_shapes=8, capture_after=2, **kw):
        super().__init__(*a, **kw)
        self.max_shapes, self.capture_after = int(max_shapes), int(capture_after)
        self.dry_captures = True                # no capture trains: exactly one optimizer step per step() call
        self._shapes, self._seen, self._clock = {}, {}, 0
        self.stats = {"graph_steps": 0, "eager_steps": 0, "captures": 0, "evictions": 0}

    def _key(self, images_host, targets_host, tok):
        return (tuple(images_host.shape), tuple(len(t["obj_labels"]) for t in targets_host),
                tuple(tuple(x) for x in tok["sums"]), tuple(tok["input_ids"].shape))

    def step(self, images_host, targets_host, text):
        tok = text if isinstance(text, dict) else self.module.transformer.tokenize(text, "cpu")
        key = self._key(images_host, targets_host, tok)
        self._clock += 1
        self._seen[key] = self._seen.get(key, 0) + 1
        entry = self._shapes.get(key)
        want_graphs = self._seen[key] >= self.capture_after or getattr(self, "flat", None) is None
        if entry is None or (want_graphs and not entry["graphs"]):
            if want_graphs:
                live = [k for k, e in self._shapes.items() if e["graphs"] and k != key]
                if len(live) >= self.max_shapes:
                    victim = min(live, key=lambda k: self._shapes[k]["used"])
                    del self._shapes[victim]                      # its graphs, pools and static buffers go with it
                    self.stats["evictions"] += 1
            self.capture(images_host, targets_host, tok, warmup=2, graphs=want_graphs)
            self.stats["captures"] += int(want_graphs)
            entry = self._shapes[key] = {"state": self.shape_state(), "graphs": want_graphs, "used": self._clock}
        else:
            self.activate(entry["state"])
        entry["used"] = self._clock
        self.stats["graph_steps" if entry["graphs"] else "eager_steps"] += 1
        loss = GraphedParSeDATrainStep.step(self, images_host, targets_host, text=tok)
        entry["state"] = self.shape_state()
        return loss
