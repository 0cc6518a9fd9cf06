"""CPU-only: the C-ABI library builds/loads and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

from tests.conftest import ROOT


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(rlipv2_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_declares_entry_points():
    syms = declared_symbols()
    for want in ("rlipv2_msda_forward_f32", "rlipv2_msda_backward_f32",
                 "rlipv2_msda_forward_f64", "rlipv2_msda_backward_f64"):
        assert want in syms


def test_libraries_export_all_declared_symbols():
    from rlipv2_b200 import build
    build.build_all()
    libs = [ctypes.CDLL(build.lib_path(n)) for n in build.TARGETS]
    for sym in declared_symbols():
        assert any(hasattr(lib, sym) for lib in libs), f"{sym} not exported"


def test_binding_matches_abi_version():
    from rlipv2_b200 import msda_abi
    import re
    hdr = open(os.path.join(ROOT, "include", "rlipv2_msda.h")).read()
    assert msda_abi.ABI_VERSION == int(re.search(r"#define RLIPV2_MSDA_ABI_VERSION (\d+)", hdr).group(1))
    assert os.path.exists(msda_abi.library_path())
    for sym in msda_abi.EXPORTS:
        assert sym in declared_symbols()
    from rlipv2_b200 import attn_abi, dense_abi, fused_abi, lsap_abi
    for abi in (attn_abi, dense_abi, fused_abi, lsap_abi):
        assert os.path.exists(abi.library_path())
        for sym in abi.EXPORTS:
            assert sym in declared_symbols()


def test_cpu_tensors_are_rejected_like_the_reference():
    # ms_deform_attn.h:35,60: AT_ERROR("Not implemented on the CPU")
    import pytest
    import torch
    from rlipv2_b200.dropin import MultiScaleDeformableAttention as MSDA
    v = torch.zeros(1, 2, 2, 2)
    sh = torch.tensor([[1, 2]])
    ls = torch.tensor([0])
    loc = torch.zeros(1, 1, 2, 1, 1, 2)
    at = torch.zeros(1, 1, 2, 1, 1)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDA.ms_deform_attn_forward(v, sh, ls, loc, at, 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDA.ms_deform_attn_backward(v, sh, ls, loc, at, torch.zeros(1, 1, 4), 64)


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure: nothing under rlipv2_b200/ may reference it
    for path in glob.glob(os.path.join(ROOT, "rlipv2_b200", "**", "*.py"), recursive=True):
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), path
