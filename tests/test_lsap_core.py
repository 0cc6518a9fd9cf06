"""CPU: the arithmetic core of the on-device matcher (rlipv2_b200/csrc/lsap_core.h, the header lsap.cu compiles for
sm_100a) built for the host with g++ (tests/lsap_host_shim.cpp) and checked against the installed scipy - the call the
reference makes (/root/reference/models/matcher.py:193).  Index outputs must be identical, ties included, for every
number of emulated lanes (the kernel uses 32).  The GPU test of the kernel itself is tests/test_zz1_lsap_gpu.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from tests.conftest import ROOT


@pytest.fixture(scope="module")
def host_lsap(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lsap") / "liblsap_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "rlipv2_b200", "csrc"), "-o", so,
                           os.path.join(ROOT, "tests", "lsap_host_shim.cpp")])
    lib = ctypes.CDLL(so)
    lib.lsap_host_solve.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong,
                                    ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]

    def solve(view, lanes=32):
        """view: a 2-d fp32 numpy array, any strides"""
        assert view.dtype == np.float32 and view.ndim == 2
        k = min(view.shape)
        a, b = np.full(k, -7, np.int64), np.full(k, -7, np.int64)
        rc = lib.lsap_host_solve(view.ctypes.data, view.shape[0], view.shape[1], view.strides[0] // 4, view.strides[1] // 4,
                                 lanes, a.ctypes.data, b.ctypes.data)
        return rc, a, b
    return solve


def _random_cost(rng, r, c, kind):
    if kind == 0:
        return rng.random((r, c)).astype(np.float32)
    if kind == 1:
        return rng.integers(0, 3, (r, c)).astype(np.float32)                      # heavy ties
    if kind == 2:
        return rng.integers(0, 2, (r, c)).astype(np.float32) * 0.5                # almost everything ties
    if kind == 3:
        return np.round(rng.standard_normal((r, c)), 1).astype(np.float32)       # negative entries, some ties
    return np.zeros((r, c), np.float32)                                           # the all-equal matrix


@pytest.mark.parametrize("lanes", [1, 2, 32])
def test_matches_scipy_on_random_and_tied_problems(host_lsap, lanes):
    rng = np.random.default_rng(lanes)
    for trial in range(1500):
        r, c = int(rng.integers(1, 48)), int(rng.integers(1, 48))
        m = _random_cost(rng, r, c, trial % 5)
        ri, ci = linear_sum_assignment(m)
        rc, a, b = host_lsap(m, lanes)
        assert rc == 0
        np.testing.assert_array_equal(a, ri, err_msg=f"trial {trial} rows {r}x{c}")
        np.testing.assert_array_equal(b, ci, err_msg=f"trial {trial} cols {r}x{c}")


def test_matcher_shaped_problems_in_the_stacked_cost_layout(host_lsap):
    """150 queries x a few triplets per image, read in place from a [levels, bs, nq, T] tensor (row stride T, the image's
    column block) - the layout the kernel indexes; also more triplets than queries (scipy then solves untransposed)"""
    rng = np.random.default_rng(7)
    for nq, sizes in ((150, [5, 3, 0, 11]), (16, [30, 16, 2]), (64, [1, 64, 65])):
        T = sum(sizes)
        C = rng.standard_normal((3, len(sizes), nq, T)).astype(np.float32)
        C[1] = np.round(C[1])                                                     # a level full of ties
        for l in range(3):
            t0 = 0
            for b, n in enumerate(sizes):
                view = C[l, b, :, t0:t0 + n]
                t0 += n
                if n == 0:
                    continue
                ri, ci = linear_sum_assignment(view)
                rc, a, bb = host_lsap(view)
                assert rc == 0 and len(a) == min(nq, n)
                np.testing.assert_array_equal(a, ri)
                np.testing.assert_array_equal(bb, ci)


def test_costs_from_the_matcher_itself(host_lsap):
    """fp32 cost matrices produced by HungarianMatcherHOI.compute_costs on near-duplicate predictions (the ties a freshly
    initialised decoder produces: identical queries -> identical cost rows)"""
    from rlipv2_b200.matcher import HungarianMatcherHOI
    torch.manual_seed(0)
    bs, nq, n_obj, n_verb = 2, 24, 6, 5
    base = {"pred_sub_logits": torch.randn(bs, 1, 2), "pred_obj_logits": torch.randn(bs, 1, n_obj),
            "pred_verb_logits": torch.randn(bs, 1, n_verb), "pred_sub_boxes": torch.rand(bs, 1, 4) * 0.5 + 0.2,
            "pred_obj_boxes": torch.rand(bs, 1, 4) * 0.5 + 0.2}
    outputs = {k: v.expand(-1, nq, -1).clone() for k, v in base.items()}          # every query identical
    outputs["pred_obj_boxes"][:, ::3] += 0.01                                    # ... except every third
    targets = []
    for k in (4, 7):
        verbs = torch.zeros(k, n_verb)
        verbs[torch.arange(k), torch.randint(0, n_verb, (k,))] = 1
        targets.append({"obj_labels": torch.randint(0, n_obj - 1, (k,)), "sub_labels": torch.zeros(k, dtype=torch.long),
                        "verb_labels": verbs, "sub_boxes": torch.rand(k, 4) * 0.4 + 0.2, "obj_boxes": torch.rand(k, 4) * 0.4 + 0.2})
    matcher = HungarianMatcherHOI(1, 1, 2.5, 1, subject_class=True)
    C, _ = matcher.compute_costs(outputs, targets)
    want = matcher.solve(C, [4, 7])
    Cn = C.numpy()
    t0 = 0
    for b, n in enumerate((4, 7)):
        rc, a, bb = host_lsap(Cn[b, :, t0:t0 + n])
        t0 += n
        assert rc == 0
        np.testing.assert_array_equal(a, want[b][0].numpy())
        np.testing.assert_array_equal(bb, want[b][1].numpy())


def test_non_finite_costs_are_reported_not_looped_on(host_lsap):
    m = np.ones((3, 4), np.float32)
    m[:, :] = np.inf
    rc, _, _ = host_lsap(m)
    assert rc == -1                                  # scipy: "cost matrix is infeasible"
    with pytest.raises(ValueError):
        linear_sum_assignment(m)


def test_plan_layout_matches_stacked_matches():
    """offsets of lsap_abi.Plan == the (level, image, match) order criterion.StackedMatches expects"""
    from rlipv2_b200 import lsap_abi
    plan = lsap_abi.Plan([5, 0, 200, 3], nq=150, n_levels=3, device="cpu")
    assert plan.ks == [5, 0, 150, 3] and plan.K == 3 * 158 and plan.T == 208 and plan.max_count == 200
    assert plan.tgt_start.tolist() == [0, 5, 5, 205]
    assert plan.out_offset.tolist() == [o + 158 * l for l in range(3) for o in (0, 5, 5, 155)]
    with pytest.raises(RuntimeError, match="CUDA"):
        lsap_abi.solve(torch.zeros(3, 4, 150, 208), plan)
