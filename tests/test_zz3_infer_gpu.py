"""GPU: the CUDA-graph inference step equals the eager one (rlipv2_b200/infer_step.py).  Runs last: written after round
1's GPU budget was spent."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graphed_inference_equals_eager():
    from rlipv2_b200 import dense
    from rlipv2_b200.infer_step import ParSeDAInference
    from tests.test_infer_step import OBJ, VERB, _build
    try:
        model, post = _build("cuda")
        g = torch.Generator().manual_seed(9)
        imgs = [torch.randn(3, 160, 192, generator=g).cuda(), torch.randn(3, 160, 192, generator=g).cuda()]
        sizes = torch.tensor([[480, 576], [320, 384]], device="cuda")
        infer = ParSeDAInference(model, post, OBJ, VERB, batch_size=2)
        eager = infer(imgs, sizes)
        infer.capture(160, 192)
        graphed = infer(imgs, sizes)
        assert infer.graph is not None
        for a, b in zip(graphed, eager):
            assert torch.equal(a["labels"], b["labels"])
            torch.testing.assert_close(a["boxes"], b["boxes"], rtol=1e-4, atol=1e-2)
            torch.testing.assert_close(a["verb_scores"], b["verb_scores"], rtol=1e-4, atol=1e-6)
        other = infer([torch.randn(3, 160, 192, device="cuda") for _ in range(2)], sizes)      # new inputs, same graph
        assert not torch.equal(other[0]["verb_scores"], graphed[0]["verb_scores"])
    finally:
        dense.set_matmul_precision("fp32")
