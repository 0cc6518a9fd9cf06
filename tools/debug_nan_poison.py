"""Reproduce 'NaN costs in the first graph replay when the process has dirty cached memory' (r02h / r02j full-suite runs):
fill the caching allocator's free blocks with NaN, build the small graphed step of tests/test_train_step_gpu.py, replay,
report which static outputs are non-finite.   python tools/debug_nan_poison.py [nographs]   (switches via RLIPV2_* env)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def poison(gb=6):
    blocks = []
    for size in (1 << 12, 1 << 16, 1 << 20, 1 << 24, 1 << 27):          # many block sizes of the allocator's pools
        n = max(4, min(256, int(gb * (1 << 30) / 5 / size)))
        blocks += [torch.full((size // 4,), float("nan"), device="cuda") for _ in range(n)]
    torch.cuda.synchronize()
    del blocks                                                           # back to the cache, contents intact


def main():
    from rlipv2_b200 import models, train_step
    graphs = "nographs" not in sys.argv
    poison()
    args = models.default_args(device="cuda", num_queries=16, synthetic_text_encoder=True)
    ts = train_step.GraphedParSeDATrainStep(args=args, device="cuda", precision="fp32", seed=0)
    ts.module.eval()
    ts.criterion.eval()
    imgs, tg = train_step.synthetic_batch(2, 160, 192, n_obj=6, n_verb=4, triplets=3, seed=1)
    text = train_step.synthetic_text(6, 4)
    poison()
    ts.capture(imgs, tg, text, warmup=2, graphs=graphs)
    poison(2)
    for it in range(3):
        try:
            loss = float(ts.replay())
        except ValueError as e:
            loss = "ValueError: " + str(e)
        torch.cuda.synchronize()
        bad = []
        if graphs:
            outputs, giou = ts._keep
            for k, v in outputs.items():
                if torch.is_tensor(v) and not torch.isfinite(v).all():
                    bad.append(k)
            for i, a in enumerate(outputs.get("aux_outputs", [])):
                bad += [f"aux{i}.{k}" for k, v in a.items() if torch.is_tensor(v) and not torch.isfinite(v).all()]
        print(f"replay {it}: loss {loss}  non-finite outputs: {bad}  params finite: {bool(torch.isfinite(ts.flat_param).all())} "
              f"grads finite: {bool(torch.isfinite(ts.flat_grad).all())}", flush=True)


if __name__ == "__main__":
    main()
