"""Pins the C oracle (oracle/msda_oracle.c) to the reference: every golden fixture was produced
by the reference's own `ms_deform_attn_core_pytorch` (ms_deform_attn_func.py:47-65) + autograd in
fp64 (oracle/gen_golden_msda.py).  CPU only."""
import numpy as np
import pytest

from oracle import msda_oracle
from tests.golden_util import load_msda, msda_cases

CASES = msda_cases()


def test_fixtures_present():
    assert len(CASES) >= 9


@pytest.mark.parametrize("name", CASES)
def test_oracle_f64_matches_reference(name):
    g = load_msda(name)
    args = [g["value"].astype(np.float64), g["spatial_shapes"], g["level_start_index"],
            g["sampling_loc"].astype(np.float64), g["attn_weight"].astype(np.float64)]
    out = msda_oracle.forward(*args)
    # torch.allclose defaults, as in the reference's fp64 check (models/ops/test.py:44)
    np.testing.assert_allclose(out, g["out"], rtol=1e-5, atol=1e-8)
    gv, gl, ga = msda_oracle.backward(*args, g["grad_out"].astype(np.float64))
    np.testing.assert_allclose(gv, g["grad_value"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gl, g["grad_sampling_loc"], rtol=1e-7, atol=1e-11)
    np.testing.assert_allclose(ga, g["grad_attn_weight"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_oracle_f32_matches_reference(name):
    g = load_msda(name)
    args = [g["value"], g["spatial_shapes"], g["level_start_index"], g["sampling_loc"], g["attn_weight"]]
    out = msda_oracle.forward(*args)
    assert out.dtype == np.float32
    # north_star tolerance: 1e-3 rel fp32; the reference's own fp32 check is rtol 1e-2 / atol 1e-3
    # (models/ops/test.py:60).  fp32 arithmetic on fp32-exact inputs is far inside either.
    np.testing.assert_allclose(out, g["out"], rtol=1e-4, atol=1e-7)
    gv, gl, ga = msda_oracle.backward(*args, g["grad_out"])
    np.testing.assert_allclose(gv, g["grad_value"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gl, g["grad_sampling_loc"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(ga, g["grad_attn_weight"], rtol=1e-4, atol=1e-6)


def test_oracle_mt_equals_scalar():
    g = load_msda("parseda_small")
    args = [g["value"], g["spatial_shapes"], g["level_start_index"], g["sampling_loc"], g["attn_weight"]]
    np.testing.assert_array_equal(msda_oracle.forward(*args), msda_oracle.forward(*args, threads=3))
    a = msda_oracle.backward(*args, g["grad_out"])
    b = msda_oracle.backward(*args, g["grad_out"], threads=3)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


def test_oracle_empty_query():
    g = load_msda("reftest_d32")
    loc = g["sampling_loc"][:, :0]
    attn = g["attn_weight"][:, :0]
    out = msda_oracle.forward(g["value"], g["spatial_shapes"], g["level_start_index"], loc, attn)
    assert out.shape == (1, 0, 64)


@pytest.mark.parametrize("name", CASES)
def test_torch_oracle_matches_reference(name):
    """oracle/msda_torch_oracle.py (the CPU stand-in used by the module-level host-logic tests)."""
    import torch
    from oracle.msda_torch_oracle import msda_core
    g = load_msda(name)
    t = lambda k: torch.from_numpy(g[k]).double()
    v, l, a = t("value").requires_grad_(True), t("sampling_loc").requires_grad_(True), t("attn_weight").requires_grad_(True)
    out = msda_core(v, g["spatial_shapes"].tolist(), l, a)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=1e-9, atol=1e-12)
    out.backward(t("grad_out"))
    np.testing.assert_allclose(v.grad.numpy(), g["grad_value"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(l.grad.numpy(), g["grad_sampling_loc"], rtol=1e-7, atol=1e-11)
    np.testing.assert_allclose(a.grad.numpy(), g["grad_attn_weight"], rtol=1e-9, atol=1e-12)
