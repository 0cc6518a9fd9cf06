"""MSDeformAttn micro-benchmark (BASELINE config 5 + encoder-shaped calls), one GPU.

    python tools/msda_microbench.py [--ref] [--iters 50] [--bwd-modes 0,1]

Timing: CUDA events on the current stream around `iters` back-to-back launches that rotate over
enough independent input sets to exceed the L2 (>= 2x126 MB working set), after 5 warm-up
launches.  GB/s = algorithmic bytes (rlipv2_b200.synth.msda_bytes) / mean launch time.
`--ref` also times the reference's own CUDA kernel (oracle/_ref, built by oracle/build_ref.py);
`--bwd-modes 0,1` times the backward under each schedule of rlipv2_msda_set_backward_mode (0 = one reduction per corner,
1 = same-cell corners of a pair merged) and restores the mode it found."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlipv2_b200 import synth  # noqa: E402
from rlipv2_b200.dropin import MultiScaleDeformableAttention as MSDA  # noqa: E402

L2_BYTES = 126e6


def time_rot(fn_list, iters):
    for f in fn_list[:5]:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn_list[i % len(fn_list)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def run_case(name, make, N, S, Lq, iters, impls, peak, bwd_modes=()):
    fwd_b, bwd_b = synth.msda_bytes(N, S, Lq)
    ncopies = max(2, min(24, int(2.2 * L2_BYTES / fwd_b) + 1))
    sets = [make(seed) for seed in range(ncopies)]
    res = {"case": name, "N": N, "S": S, "Lq": Lq, "fwd_MB": fwd_b / 1e6, "bwd_MB": bwd_b / 1e6,
           "copies": ncopies}
    for iname, mod in impls.items():
        f = [lambda s=s: mod.ms_deform_attn_forward(s[0], s[1], s[2], s[3], s[4], 64) for s in sets]
        b = [lambda s=s: mod.ms_deform_attn_backward(s[0], s[1], s[2], s[3], s[4], s[5], 64) for s in sets]
        tf = time_rot(f, iters)
        tb = time_rot(b, iters)
        res[iname] = {"fwd_us": tf * 1e6, "fwd_GBs": fwd_b / tf / 1e9, "fwd_frac": fwd_b / tf / 1e9 / peak,
                      "bwd_us": tb * 1e6, "bwd_GBs": bwd_b / tb / 1e9, "bwd_frac": bwd_b / tb / 1e9 / peak}
    if bwd_modes:
        from rlipv2_b200 import msda_abi
        before = msda_abi.get_backward_mode()
        b = [lambda s=s: MSDA.ms_deform_attn_backward(s[0], s[1], s[2], s[3], s[4], s[5], 64) for s in sets]
        for mode in bwd_modes:
            msda_abi.set_backward_mode(mode)
            tb = time_rot(b, iters)
            res[f"ours_bwd_mode{mode}"] = {"bwd_us": tb * 1e6, "bwd_GBs": bwd_b / tb / 1e9, "bwd_frac": bwd_b / tb / 1e9 / peak}
        msda_abi.set_backward_mode(before)
    print(json.dumps(res))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--cases", default="dec16,dec2,encrand2,enc2,enc2init,enc2n025")
    ap.add_argument("--bwd-modes", default="", help="comma-separated backward schedules to time (rlipv2_msda_set_backward_mode)")
    args = ap.parse_args()
    peak = 6581.2
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    impls = {"ours": MSDA}
    if args.ref:
        from oracle import build_ref
        ref = build_ref.load()
        if ref is not None:
            impls["ref"] = ref
        else:
            print("# oracle/_ref not built; skipping --ref", file=sys.stderr)
    Sm = sum(h * w for h, w in synth.LEVELS_MICRO)
    Se = sum(h * w for h, w in synth.LEVELS_800x1333)
    cases = {
        "dec16": ("config5 decoder-shaped N=16 Lq=300 (rand loc)",
                  lambda seed: synth.random_inputs(16, 300, synth.LEVELS_MICRO, seed=seed), 16, Sm, 300),
        "dec2": ("config5 decoder-shaped N=2 Lq=300 (rand loc)",
                 lambda seed: synth.random_inputs(2, 300, synth.LEVELS_MICRO, seed=seed), 2, Sm, 300),
        "encrand2": ("config5 encoder-shaped N=2 Lq=S=13294 (rand loc)",
                     lambda seed: synth.random_inputs(2, Sm, synth.LEVELS_MICRO, seed=seed), 2, Sm, Sm),
        "enc2": ("encoder call 800x1333 N=2 Lq=S=22223 (local sampling, offsets = ring + 1 cell of noise)",
                 lambda seed: synth.encoder_inputs(2, synth.LEVELS_800x1333, seed=seed), 2, Se, Se),
        "enc2init": ("encoder call 800x1333 N=2 Lq=S=22223 (offsets = ring pattern of the initialisation, no noise)",
                     lambda seed: synth.encoder_inputs(2, synth.LEVELS_800x1333, seed=seed, noise_px=0.0), 2, Se, Se),
        "enc2n025": ("encoder call 800x1333 N=2 Lq=S=22223 (ring + 0.25 cell of noise)",
                     lambda seed: synth.encoder_inputs(2, synth.LEVELS_800x1333, seed=seed, noise_px=0.25), 2, Se, Se),
    }
    for key in args.cases.split(","):
        name, make, N, S, Lq = cases[key]
        run_case(name, make, N, S, Lq, args.iters, impls, peak, [int(m) for m in args.bwd_modes.split(",") if m])


if __name__ == "__main__":
    main()
