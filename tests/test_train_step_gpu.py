"""GPU: the CUDA-graph train step (two graphs + host LSAP, flat gradient buffer) must follow the same
trajectory as the eager step that restates engine.py:99-172 literally."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(cls):
    from rlipv2_b200 import models, train_step
    args = models.default_args(device="cuda", num_queries=16, synthetic_text_encoder=True)
    ts = cls(args=args, device="cuda", precision="fp32", seed=0)
    ts.module.eval()            # dropout off (hard-coded p=0.1 dropouts would de-correlate the two runs);
    ts.criterion.eval()         # gradients, clipping and AdamW are unaffected by eval()
    imgs, tg = train_step.synthetic_batch(2, 160, 192, n_obj=6, n_verb=4, triplets=3, seed=1)
    text = train_step.synthetic_text(6, 4)
    return ts, imgs, tg, text


def test_graphed_step_matches_eager_step():
    from rlipv2_b200 import dense, train_step
    try:
        eager, imgs, tg, text = _make(train_step.ParSeDATrainStep)
        samples, targets = eager.to_device(imgs, tg)
        eager_losses = [float(eager.step_device(samples, targets, text)) for _ in range(4)]
        graphed, imgs, tg, text = _make(train_step.GraphedParSeDATrainStep)
        graphed.capture(imgs, tg, text, warmup=2)              # 2 optimizer steps happen here
        g_losses = [float(graphed.replay()) for _ in range(2)]  # steps 3 and 4
        assert eager_losses[0] > 0
        for a, b in zip(eager_losses[2:], g_losses):
            assert abs(a - b) <= 2e-3 * abs(a), (eager_losses, g_losses)
        # parameters after 4 steps agree
        pe = dict(eager.module.named_parameters())
        pg = dict(graphed.module.named_parameters())
        for k in ("transformer.level_embed", "input_proj.0.0.weight", "transformer.encoder.layers.3.linear1.weight",
                  "transformer.encoder.VLFuse_layers.1.b_attn.attn.v_proj.weight", "tgt_embed.weight",
                  "transformer.verb_decoder.sub_bbox_embed.1.layers.0.weight"):
            d = (pe[k].detach() - pg[k].detach()).norm() / (pe[k].detach().norm() + 1e-12)
            assert float(d) < 2e-3, (k, float(d))
        # a new batch through the public step(): H2D into the static buffers, then replay
        imgs2, tg2 = train_step.synthetic_batch(2, 160, 192, n_obj=6, n_verb=4, triplets=3, seed=2)
        l5 = float(graphed.step(imgs2, tg2))
        assert l5 == l5 and l5 > 0
        graphed.check()                                         # no flag wait timed out
        assert graphed.flag_wait and int(graphed.d_seq.item()) == 4 == graphed.flag_seq   # priming replay + 3
    finally:
        dense.set_matmul_precision("fp32")
