"""Instrumented reproduction of the order-dependent NaN (r02h / r02j / r02o-r: tests/test_dense_gpu.py + test_parseda_model.py +
test_postprocess.py + test_small_ops_gpu.py, then the graphed step): run those files in-process, then the body of
tests/test_train_step_gpu.py::test_graphed_step_matches_eager_step with finiteness checks and the stream handles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytest  # noqa: E402
import torch  # noqa: E402

rc = pytest.main(["tests/test_dense_gpu.py", "tests/test_parseda_model.py", "tests/test_postprocess.py", "tests/test_small_ops_gpu.py",
                  "-m", "gpu", "-q", "--tb=line", "-p", "no:cacheprovider"])
print("prefix rc", rc, flush=True)

from rlipv2_b200 import criterion as crit_mod, dense, models, train_step  # noqa: E402


def handles(ts):
    out = {"cap": ts.cap_stream.cuda_stream, "zero": getattr(getattr(ts, "_zero_stream", None), "cuda_stream", None)}
    for name, m in ts.module.named_modules():
        for attr in ("_pos_stream", "_text_stream", "_lang_stream", "_value_stream"):
            st = getattr(m, attr, None)
            if st is not None:
                out[f"{name}.{attr}"] = st.cuda_stream
    for dev, st in dense._param_grad_streams.items():
        out[f"param_grad[{dev}]"] = st.cuda_stream
    for dev, sts in crit_mod._BRANCH_STREAMS.items():
        for i, st in enumerate(sts):
            out[f"branch[{i}]"] = st.cuda_stream
    return out


def finite_report(tag, ts):
    bad = []
    if not torch.isfinite(ts.flat_param).all():
        bad.append("flat_param")
    if not torch.isfinite(ts.flat_grad).all():
        bad.append("flat_grad")
    for n, v in (("exp_avg", ts.exp_avg), ("exp_avg_sq", ts.exp_avg_sq)):
        if not torch.isfinite(v).all():
            bad.append(n)
    outputs, giou = ts._keep
    for k, v in outputs.items():
        if torch.is_tensor(v) and not torch.isfinite(v).all():
            bad.append("out." + k)
    print(tag, "non-finite:", bad, "step_t", float(ts.step_t), "d_err", int(ts.d_err), "lr", ts.lr_dev.tolist(), flush=True)


args = models.default_args(device="cuda", num_queries=16, synthetic_text_encoder=True)
eager = train_step.ParSeDATrainStep(args=args, device="cuda", precision="fp32", seed=0)
eager.module.eval(); eager.criterion.eval()
imgs, tg = train_step.synthetic_batch(2, 160, 192, n_obj=6, n_verb=4, triplets=3, seed=1)
text = train_step.synthetic_text(6, 4)
samples, targets = eager.to_device(imgs, tg)
print("eager losses", [float(eager.step_device(samples, targets, text)) for _ in range(4)], flush=True)
args = models.default_args(device="cuda", num_queries=16, synthetic_text_encoder=True)
ts = train_step.GraphedParSeDATrainStep(args=args, device="cuda", precision="fp32", seed=0)
ts.module.eval(); ts.criterion.eval()
ts.capture(imgs, tg, text, warmup=2)
torch.cuda.synchronize()
h = handles(ts)
print("stream handles:", {k: hex(v) if v else v for k, v in h.items()}, flush=True)
vals = [v for v in h.values() if v]
print("aliased handles:", sorted({hex(v) for v in vals if vals.count(v) > 1}), flush=True)
finite_report("after capture", ts)
for it in range(2):
    ts.graph_a.replay()
    torch.cuda.synchronize()
    finite_report(f"replay {it} after graph A", ts)
    cost_ok = bool(torch.isfinite(ts.h_cost).all())
    print("  h_cost finite:", cost_ok, flush=True)
    if not cost_ok:
        break
    ts._solve_assignment_host()
    ts.flag_seq += 1
    ts.np_flag[0] = ts.flag_seq
    ts.graph_b.replay()
    torch.cuda.synchronize()
    finite_report(f"replay {it} after graph B", ts)
