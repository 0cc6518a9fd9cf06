"""CPU: the reduction-merging rule of the fast MSDeformAttn backward (rlipv2_b200/csrc/msda_merge.h, the header msda.cu
compiles for sm_100a) built for the host with g++ (tests/msda_merge_host_shim.cpp).

Reference semantics: every valid corner of every sampling point adds `bilinear weight x attention x grad_out row` to its
cell of grad_value (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:125,134,143,152).  The merged schedule
must produce the same cell sums with one reduction per distinct cell of a (pair, level): checked per level against a
brute-force cell map, and for whole calls against the C oracle's grad_value.  The GPU tests of the kernel itself are
tests/test_msda_gpu.py / test_msda_proj_gpu.py (merged mode)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import msda_oracle
from tests.conftest import ROOT
from tests.golden_util import load_msda


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("msda_merge") / "libmsda_merge_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I",
                           os.path.join(ROOT, "rlipv2_b200", "csrc"), "-o", so,
                           os.path.join(ROOT, "tests", "msda_merge_host_shim.cpp")])
    lib = ctypes.CDLL(so)
    vp, i = ctypes.c_void_p, ctypes.c_int
    lib.msda_merge_level.argtypes = [vp, vp, i, i, vp, vp]
    lib.msda_merge_grad_value.argtypes = [vp] * 5 + [i] * 4 + [vp]
    lib.msda_merge_grad_value.restype = ctypes.c_longlong
    return lib


def _level(lib, loc, attn, H, W):
    loc = np.ascontiguousarray(loc, np.float32)
    attn = np.ascontiguousarray(attn, np.float32)
    s = np.zeros((4, 4), np.float32)
    geo = np.zeros((4, 3), np.int32)
    lib.msda_merge_level(loc.ctypes.data, attn.ctypes.data, H, W, s.ctypes.data, geo.ctypes.data)
    return s, geo


def _brute(loc, attn, H, W):
    """cell -> [sum of weights (fp64), owner (point, corner)] with the reference's validity rules (cuh:285-288, 56-78)"""
    cells = {}
    per = np.zeros((4, 4))
    valid = np.zeros((4, 4), bool)
    key = np.full((4, 4, 2), -9, np.int64)
    for p in range(4):
        h_im = np.float32(loc[p, 1]) * np.float32(H) - np.float32(0.5)
        w_im = np.float32(loc[p, 0]) * np.float32(W) - np.float32(0.5)
        if not (h_im > -1 and w_im > -1 and h_im < H and w_im < W):
            continue
        h0, w0 = int(np.floor(h_im)), int(np.floor(w_im))
        lh, lw = np.float32(h_im - np.float32(h0)), np.float32(w_im - np.float32(w0))
        for k, (cy, cx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
            y, x = h0 + cy, w0 + cx
            if 0 <= y <= H - 1 and 0 <= x <= W - 1:
                w = float(lh if cy else 1 - lh) * float(lw if cx else 1 - lw) * float(attn[p])
                valid[p, k] = True
                key[p, k] = (y, x)
                per[p, k] = w
                c = cells.setdefault((y, x), [0.0, (p, k)])
                c[0] += w
    return cells, per, valid, key


def _cases(rng, n):
    for t in range(n):
        H, W = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        kind = t % 5
        c = rng.uniform(-0.3, 1.3, 2)
        if kind == 0:       # the initialisation's ring: whole-cell steps along one direction
            d = np.array([(1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1)][t // 5 % 8], float)
            loc = c + np.arange(1, 5)[:, None] * d / np.array([W, H])
        elif kind == 1:     # ring + noise
            loc = c + (np.arange(1, 5)[:, None] * np.array([1.0, 0.0]) + rng.normal(0, 0.6, (4, 2))) / np.array([W, H])
        elif kind == 2:     # all four in one cell or its neighbours
            loc = c + rng.uniform(-0.7, 0.7, (4, 2)) / np.array([W, H])
        elif kind == 3:     # exact cell centres (fractional parts 0 -> zero-weight corners), repeated points
            ij = rng.integers(-1, 9, (4, 2))
            ij[rng.integers(0, 4)] = ij[rng.integers(0, 4)]
            loc = (ij + 0.5) / np.array([W, H])
        else:               # anywhere
            loc = rng.uniform(-0.5, 1.5, (4, 2))
        attn = rng.uniform(0, 1, 4)
        if t % 7 == 0:
            attn[rng.integers(0, 4)] = 0.0
        yield loc.astype(np.float32), attn.astype(np.float32), H, W


def test_level_merge_equals_brute_force_cell_map(shim):
    rng = np.random.default_rng(5)
    merged_any = zero_any = 0
    for loc, attn, H, W in _cases(rng, 6000):
        s, geo = _level(shim, loc, attn, H, W)
        cells, per, valid, key = _brute(loc, attn, H, W)
        # every corner that is not the owner of its cell, and every corner outside, carries exactly zero
        for p in range(4):
            for k in range(4):
                if not valid[p, k]:
                    assert s[p, k] == 0.0
                    continue
                tot, owner = cells[tuple(key[p, k])]
                if owner == (p, k):
                    assert s[p, k] == pytest.approx(tot, rel=2e-6, abs=1e-7), (loc, attn, H, W, p, k)
                    merged_any += tot != pytest.approx(per[p, k], rel=1e-9, abs=0)
                else:
                    assert s[p, k] == 0.0, (loc, attn, H, W, p, k)
        zero_any += int(((s == 0) & valid).sum() > 0)
        assert float(s.sum()) == pytest.approx(per.sum(), rel=1e-5, abs=1e-6)
        # at most one non-zero entry per cell
        assert int((s != 0).sum()) <= len(cells)
    assert merged_any > 1000 and zero_any > 1000


def test_point_geometry_matches_the_reference_rules(shim):
    rng = np.random.default_rng(6)
    for loc, attn, H, W in _cases(rng, 2000):
        _, geo = _level(shim, loc, attn, H, W)
        cells, per, valid, key = _brute(loc, attn, H, W)
        for p in range(4):
            bits = int(geo[p, 2])
            got = [(bits & 5) == 5, (bits & 9) == 9, (bits & 6) == 6, (bits & 10) == 10]
            assert got == list(valid[p]), (loc[p], H, W, bits)
            for k, (cy, cx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
                if valid[p, k]:
                    assert (geo[p, 0] + cy, geo[p, 1] + cx) == tuple(key[p, k])


def _encoder_like(rng, shapes, N, noise):
    """one query per cell, ring offsets + noise (rlipv2_b200/synth.py::encoder_inputs, numpy)"""
    M, P = 8, 4
    ref = np.concatenate([np.stack(np.meshgrid((np.arange(W) + 0.5) / W, (np.arange(H) + 0.5) / H), -1).reshape(-1, 2)
                          for H, W in shapes])
    S = ref.shape[0]
    th = np.arange(M) * 2 * np.pi / M
    ring = np.stack([np.cos(th), np.sin(th)], -1)
    ring = ring / np.abs(ring).max(-1, keepdims=True)
    offs = ring.reshape(1, 1, M, 1, 1, 2) * np.arange(1, P + 1).reshape(1, 1, 1, 1, P, 1)
    offs = offs + noise * rng.standard_normal((N, S, M, len(shapes), P, 2))
    norm = np.array([[w, h] for h, w in shapes], float).reshape(1, 1, 1, -1, 1, 2)
    loc = ref.reshape(1, S, 1, 1, 1, 2) + offs / norm
    return loc.astype(np.float32), S


@pytest.mark.parametrize("noise", [0.0, 0.3, 1.0])
def test_merged_grad_value_equals_the_oracle(shim, noise):
    rng = np.random.default_rng(int(noise * 10))
    shapes = [(12, 20), (6, 10), (3, 5), (2, 3)]
    N, M = 2, 8
    loc, S = _encoder_like(rng, shapes, N, noise)
    sh = np.asarray(shapes, np.int64)
    lsi = np.concatenate([[0], np.cumsum(sh.prod(1))[:-1]]).astype(np.int64)
    value = rng.standard_normal((N, S, M, 32)).astype(np.float32)
    attn = rng.uniform(0, 1, (N, S, M, 4, 4)).astype(np.float32)
    attn /= attn.sum((-1, -2), keepdims=True)
    gout = rng.standard_normal((N, S, M * 32)).astype(np.float32)
    gv64, _, _ = msda_oracle.backward(value.astype(np.float64), sh, lsi, loc.astype(np.float64), attn.astype(np.float64),
                                      gout.astype(np.float64))
    got = np.zeros((N, S, M, 32), np.float64)
    issued = shim.msda_merge_grad_value(sh.ctypes.data, lsi.ctypes.data, loc.ctypes.data, attn.ctypes.data, gout.ctypes.data,
                                        N, S, M, S, got.ctypes.data)
    scale = np.abs(gv64).max()
    assert np.abs(got - gv64).max() <= 2e-6 * scale
    # and it does merge: the unmerged schedule issues one reduction per valid corner
    valid = 0
    for l, (H, W) in enumerate(shapes):
        x = loc[..., l, :, 0] * np.float32(W) - np.float32(0.5)
        y = loc[..., l, :, 1] * np.float32(H) - np.float32(0.5)
        inside = (x > -1) & (y > -1) & (x < W) & (y < H)
        x0, y0 = np.floor(x), np.floor(y)
        for cy in (0, 1):
            for cx in (0, 1):
                valid += int((inside & (y0 + cy >= 0) & (y0 + cy <= H - 1) & (x0 + cx >= 0) & (x0 + cx <= W - 1)).sum())
    assert issued < (0.9 if noise >= 1.0 else 0.8) * valid, (issued, valid)


def test_merged_grad_value_on_the_reference_fixture(shim):
    g = load_msda("parseda_small")
    v = g["value"]
    N, S, M, D = v.shape
    Lq = g["sampling_loc"].shape[1]
    assert D == 32 and g["sampling_loc"].shape[3:5] == (4, 4)
    loc = np.ascontiguousarray(g["sampling_loc"], np.float32)
    attn = np.ascontiguousarray(g["attn_weight"], np.float32)
    gout = np.ascontiguousarray(g["grad_out"], np.float32)
    sh = np.ascontiguousarray(g["spatial_shapes"], np.int64)
    lsi = np.ascontiguousarray(g["level_start_index"], np.int64)
    got = np.zeros((N, S, M, 32), np.float64)
    shim.msda_merge_grad_value(sh.ctypes.data, lsi.ctypes.data, loc.ctypes.data, attn.ctypes.data, gout.ctypes.data,
                               N, S, M, Lq, got.ctypes.data)
    ref = g["grad_value"].astype(np.float64)
    assert np.abs(got - ref).max() <= 3e-6 * np.abs(ref).max()


def test_mode_switch_and_env_variable_without_a_gpu():
    """the schedule switch is host state of the library: default 2 (merged for encoder-shaped calls), RLIPV2_MSDA_BWD_MERGE
    sets it at import (r02u: the switch was read before the module had defined its error check)"""
    import sys
    code = "from rlipv2_b200 import msda_abi; print(msda_abi.get_backward_mode())"
    for env, want in (({}, "2"), ({"RLIPV2_MSDA_BWD_MERGE": "0"}, "0"), ({"RLIPV2_MSDA_BWD_MERGE": "1"}, "1")):
        e = {k: v for k, v in os.environ.items() if k != "RLIPV2_MSDA_BWD_MERGE"}
        e.update(env)
        out = subprocess.run([sys.executable, "-c", code], env=e, cwd=ROOT, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr[-500:]
        assert out.stdout.strip().splitlines()[-1] == want
