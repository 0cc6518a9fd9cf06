"""N > 1 host logic on CPU: two `gloo` ranks.

* `FlatParams` (the flat-gradient data-parallel scheme of the graphed step): after the single all-reduce the flat
  gradient equals the gradient of the global-batch loss, `clip_` equals `clip_grad_norm_`, and the replicas stay
  bit-identical after an optimizer step on the views.
* `EarlyReducer` + `grad_ready` markers (the overlapped all-reduce): ranges reduced from inside the backward plus the
  remainder after it leave exactly what the single whole-buffer all-reduce leaves.
* `SetCriterionHOI`: `num_interactions` is all-reduced over ranks (/root/reference/models/hoi.py:4737-4740), so the
  rank-mean of the box losses equals the single-process loss over the union of the ranks' images.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.test_criterion_stacked import _setup

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model(seed=0):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))


def _groups(model, lrs=(1e-2, 1e-3)):
    ps = list(model.parameters())
    return [(ps[:2], lrs[0]), (ps[2:], lrs[1])]        # group sizes 35 and 18: exercises the 4-element alignment


def _worker(rank, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from rlipv2_b200.flat_dp import FlatParams
        torch.manual_seed(1)
        x, y = torch.randn(8, 6) * 3, torch.randn(8, 3)
        # ---- flat-gradient data parallelism
        model = _model(seed=rank)                     # replicas start different on purpose
        flat = FlatParams(_groups(model), "cpu")
        flat.broadcast_params(0)
        lo = rank * 4
        flat.zero_grad()
        torch.nn.functional.mse_loss(model(x[lo:lo + 4]), y[lo:lo + 4]).backward()
        local = flat.flat_grad.clone()
        flat.allreduce_mean_()
        g_after_reduce = flat.flat_grad.clone()
        flat.clip_(0.1)
        g_after_clip = flat.flat_grad.clone()
        # the fused variant of the graphed step: rank SUM in the buffer, clip coefficient / world applied by the optimizer
        flat.flat_grad.copy_(local)
        flat.allreduce_sum_()
        g_scaled = flat.flat_grad * flat.clip_scale(0.1)
        g_sum = flat.flat_grad.clone()
        full_scale = flat.clip_scale(0.1)
        # the sharded optimizer step: this rank's shard of the sum, the same clip coefficient from shard norms, an update of
        # the shard only, all-gather of the parameters
        flat.flat_grad.copy_(local)
        flat.reduce_scatter_sum_()
        sh0, sh1 = flat.shard_range()
        assert (sh1 - sh0) * WORLD == flat.flat_grad.numel() and sh0 % 4 == 0
        shard_sum_ok = torch.allclose(flat.flat_grad[sh0:sh1], g_sum[sh0:sh1], rtol=1e-6, atol=1e-7)
        scale_sh = flat.clip_scale_sharded(0.1)
        p_before = flat.flat_param.clone()
        flat.flat_param[sh0:sh1] -= 0.5 * scale_sh * flat.flat_grad[sh0:sh1]
        flat.all_gather_params_()
        p_sharded = flat.flat_param.clone()
        p_full = p_before - 0.5 * full_scale * g_sum
        flat.flat_param.copy_(p_before)
        sharded = (shard_sum_ok, scale_sh.clone(), full_scale.clone(), p_sharded, p_full)
        # the overlapped variant: ranges all-reduced from gradient-readiness markers inside the backward, the rest after it
        from rlipv2_b200 import grad_ready
        from rlipv2_b200.flat_dp import EarlyReducer
        (s0, e0, _), (s1, e1, _) = flat.group_ranges
        reducer = EarlyReducer(flat, [({"after_first"}, s1, e1)])          # second Linear's range: final once its input's
        assert reducer.active                                              # gradient has been formed
        grad_ready.set_callback(reducer.on_tag)
        flat.zero_grad()
        reducer.begin()
        h = grad_ready.mark(model[1](model[0](x[lo:lo + 4])), "after_first")
        torch.nn.functional.mse_loss(model[2](h), y[lo:lo + 4]).backward()
        assert reducer.launched == [(s1, e1)] and reducer.remaining()[0] == (0, s1)
        reducer.finish()
        grad_ready.set_callback(None)
        g_early = flat.flat_grad.clone()
        flat.flat_grad.copy_(g_after_clip)
        opt = torch.optim.AdamW([{"params": pl, "lr": lr} for pl, lr in _groups(model)], weight_decay=1e-4)
        opt.step()                                     # updates the views == the flat buffer
        # ---- criterion normalisation over ranks
        criterion, outputs, targets = _setup(bs=2, sizes=(3, 1), giou_verb_label=False)
        mine = {k: v[rank:rank + 1] for k, v in outputs.items() if k != "aux_outputs"}
        mine["aux_outputs"] = [{k: v[rank:rank + 1] for k, v in a.items()} for a in outputs["aux_outputs"]]
        losses = criterion(mine, targets[rank:rank + 1])
        keys = ["loss_sub_bbox", "loss_sub_giou", "loss_sub_bbox_0", "loss_sub_giou_1"]
        vec = torch.stack([losses[k].detach().reshape(()) for k in keys])
        dist.all_reduce(vec)
        q.put((rank, g_after_reduce, g_after_clip, flat.flat_param.clone(), flat.group_ranges, vec / WORLD, g_scaled,
               g_sum, g_early, sharded))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gloo_ranks():
    ctx = mp.get_context("spawn")
    res = None
    for attempt in range(3):               # a free port can be taken between probing and the rendezvous: retry
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
        for p in procs:
            p.start()
        got, waited = [], 0
        while len(got) < WORLD and waited < 240:
            try:
                got.append(q.get(timeout=5))
            except Exception:
                waited += 5
                if any(p.exitcode not in (None, 0) for p in procs):      # a rank died (e.g. the port was taken)
                    break
        res = sorted(got, key=lambda r: r[0]) if len(got) == WORLD else None
        for p in procs:
            p.join(60)
            if p.is_alive():
                p.kill()
        if res is not None and all(p.exitcode == 0 for p in procs):
            break
        res = None
    assert res is not None, "two gloo ranks did not complete in 3 attempts"
    # single-process reference: the global batch through one replica
    torch.manual_seed(1)
    x, y = torch.randn(8, 6) * 3, torch.randn(8, 3)
    ref = _model(seed=0)
    torch.nn.functional.mse_loss(ref(x), y).backward()

    def flatten(tensors, ranges):
        out = torch.zeros(res[0][1].numel())             # (includes the tail padding to 4 * world elements)
        groups = _groups(ref)
        for (plist, _), (start, _, _) in zip(groups, ranges):
            o = start
            for p, t in zip(plist, tensors[:len(plist)]):
                out[o:o + p.numel()] = t.reshape(-1)
                o += p.numel()
            tensors = tensors[len(plist):]
        return out

    ranges = res[0][4]
    assert ranges[0][0] == 0 and ranges[1][0] % 4 == 0 and ranges[1][0] >= ranges[0][1]
    g_ref = flatten([p.grad for p in ref.parameters()], ranges)
    for r in res:
        torch.testing.assert_close(r[1], g_ref, rtol=1e-5, atol=1e-6)
    total = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.1)
    assert total > 0.1                               # the clip is active in this test
    g_clip = flatten([p.grad for p in ref.parameters()], ranges)
    for r in res:
        torch.testing.assert_close(r[2], g_clip, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(r[6], g_clip, rtol=1e-5, atol=1e-7)      # sum + clip_scale == mean + clip_
        assert torch.equal(r[7], r[8])                   # marker-driven partial all-reduces == the one whole all-reduce
        ok, scale_sh, scale_full, p_sharded, p_full = r[9]
        assert ok                                        # reduce-scatter leaves the rank sum in this rank's shard
        torch.testing.assert_close(scale_sh, scale_full, rtol=1e-6, atol=0)     # same clip coefficient from shard norms
        torch.testing.assert_close(p_sharded, p_full, rtol=1e-6, atol=1e-7)     # shard updates + all-gather == full update
    assert torch.equal(res[0][9][3], res[1][9][3])       # and the replicas are bit-identical after the all-gather
    opt = torch.optim.AdamW([{"params": pl, "lr": lr} for pl, lr in _groups(ref)], weight_decay=1e-4)
    opt.step()
    p_ref = flatten([p.data for p in ref.parameters()], ranges)
    assert torch.equal(res[0][3], res[1][3])         # replicas stay bit-identical
    torch.testing.assert_close(res[0][3], p_ref, rtol=1e-5, atol=1e-6)
    # criterion: rank-mean of the num_interactions-normalised losses == single process over both images
    criterion, outputs, targets = _setup(bs=2, sizes=(3, 1), giou_verb_label=False)
    full = criterion(outputs, targets)
    keys = ["loss_sub_bbox", "loss_sub_giou", "loss_sub_bbox_0", "loss_sub_giou_1"]
    want = torch.stack([full[k].detach().reshape(()) for k in keys])
    torch.testing.assert_close(res[0][5], want, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(res[1][5], want, rtol=1e-5, atol=1e-6)
