import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def msda_cpu_stub(monkeypatch):
    """Host-logic tests on a CPU-only box: swap the CUDA op for the (golden-pinned) torch oracle.
    Never used when a test runs with device='cuda'."""
    from oracle.msda_torch_oracle import CPUFunctionStub
    import rlipv2_b200.ms_deform_attn as m
    monkeypatch.setattr(m, "MSDeformAttnFunction", CPUFunctionStub)
    return CPUFunctionStub
