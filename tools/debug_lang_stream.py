"""Do the flat gradients of one step depend on whether the label stream's RobertaLayer runs on its side stream?
(r02k: RLIPV2_LANG_STREAM=0 gave another loss trajectory.)  Same parameters (lr = 0), same batch: gradient buffers of the
eager flavour of the graphed step with the side stream on / off, and plain autograd gradients as the referee."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    from rlipv2_b200 import models, parseda_transformer, train_step
    args = models.default_args(device="cuda", num_queries=16, synthetic_text_encoder=True)
    ts = train_step.GraphedParSeDATrainStep(args=args, device="cuda", precision="fp32", seed=0)
    ts.module.eval()
    ts.criterion.eval()
    imgs, tg = train_step.synthetic_batch(2, 160, 192, n_obj=6, n_verb=4, triplets=3, seed=1)
    text = train_step.synthetic_text(6, 4)
    ts.capture(imgs, tg, text, warmup=2, graphs=False)
    ts.set_lr(0.0)
    names = [n for n, p in ts.module.named_parameters() if any(p is q for q in ts.params)]
    res = {}
    for tag, on in (("on", True), ("off", False), ("on2", True)):
        parseda_transformer._LANG_STREAM = on
        loss = float(ts.replay())
        torch.cuda.synchronize()
        res[tag] = (loss, ts.flat_grad.clone())
    # referee: plain autograd on the same module (no fused accumulation), side stream off
    parseda_transformer._LANG_STREAM = False
    for p in ts.params:
        p._fuse_grad = False
        p.grad = None
    samples, targets = ts.s_samples, ts.s_targets
    cache = ts.module(samples, encode_and_save=True, text=ts.s_tok, targets=targets)
    out = ts.module(samples, encode_and_save=False, memory_cache=cache, text=ts.s_tok, targets=targets)
    total = ts._weighted_total(ts.criterion(out, targets))
    total.backward()
    torch.cuda.synchronize()
    ref = torch.zeros_like(ts.flat_grad)
    for p, off in zip(ts.params, ts.param_offsets):
        if p.grad is not None:
            ref[off:off + p.numel()] = p.grad.reshape(-1)
    print("losses", {k: v[0] for k, v in res.items()}, "autograd", float(total))
    for tag in ("on", "off", "on2"):
        g = res[tag][1]
        print(tag, "vs autograd: rel", float((g - ref).norm() / ref.norm()), "| vs on:", float((g - res['on'][1]).norm() / ref.norm()))
    worst = []
    for n, p, off in zip(names, ts.params, ts.param_offsets):
        a, b, r = res["on"][1][off:off + p.numel()], res["off"][1][off:off + p.numel()], ref[off:off + p.numel()]
        d = float((a - b).norm() / (r.norm() + 1e-12))
        if d > 1e-4:
            worst.append((d, n, float((a - r).norm() / (r.norm() + 1e-12)), float((b - r).norm() / (r.norm() + 1e-12))))
    worst.sort(reverse=True)
    print("parameters whose gradient differs on/off (rel diff, name, on-vs-autograd, off-vs-autograd):")
    for w in worst[:25]:
        print("  %.3e  %-70s on %.3e  off %.3e" % w)
    print("count", len(worst))


if __name__ == "__main__":
    main()
