"""GPU (one device): plumbing of the overlapped gradient all-reduce inside the captured backward graph - markers, per-tag
events, the communication stream forked from and joined back into the capture (RLIPV2_ALLREDUCE_OVERLAP=force installs
them at world size 1, where the collectives themselves are no-ops).  The trajectory must equal the plain graphed step.
The 2-GPU A/B is tools/gpu_round2_overlap.sh.  Runs last: written after round 1's GPU budget was spent."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graphed_step_with_early_reduce_plumbing_equals_plain(monkeypatch):
    from rlipv2_b200 import dense, models, train_step

    def run(mode):
        monkeypatch.setenv("RLIPV2_ALLREDUCE_OVERLAP", mode)
        args = models.default_args(device="cuda", num_queries=16, synthetic_text_encoder=True)
        ts = train_step.GraphedParSeDATrainStep(args=args, device="cuda", precision="fp32", seed=0)
        ts.module.eval()
        ts.criterion.eval()
        imgs, tg = train_step.synthetic_batch(2, 160, 192, n_obj=6, n_verb=4, triplets=3, seed=1)
        try:
            ts.capture(imgs, tg, train_step.synthetic_text(6, 4), warmup=2)
            losses = [float(ts.replay()) for _ in range(3)]
            ts.check()
            launched = None if ts.reducer is None else sorted(ts.reducer.launched)
            entries = None if ts.reducer is None else sorted((s, e) for _, s, e in ts.reducer.entries)
        finally:
            ts.uninstall_early_reducer()
        return losses, launched, entries

    try:
        plain, none_launched, _ = run("0")
        over, launched, entries = run("force")
        assert none_launched is None
        assert launched == entries and len(entries) == 3          # every range was launched from inside the backward
        for a, b in zip(plain, over):
            assert abs(a - b) <= 2e-3 * abs(a), (plain, over)          # (fp32 reduction order differs run to run)
    finally:
        dense.set_matmul_precision("fp32")
