"""bench.py - headline measurement (driver contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads
  msda_step  (default until the full ParSeDA train step lands in this file): the 12 MSDeformAttn
             calls (forward + backward) one RLIPv2-ParSeDA R50 train step issues on a batch of two
             3x800x1333 images with 300 queries (BASELINE config 2): 6 encoder calls
             (Lq = S = 22223), 3 pair-decoder calls (Lq = 300), 3 verb-decoder calls (Lq = 150).
             One "step" = one pass over that batch; images/sec = batch / step time.

Every rank runs the same per-GPU batch (weak scaling, no data-path collective: the op shards by
image); timing = CUDA events around exactly K steps, bracketed by barrier + synchronize, max over
ranks.  L2 hygiene: each step rotates over enough independent input sets that the working set
exceeds 2x the 126 MB L2 ("l2": "inputs>L2" in config).

`--impl reference` times the reference's CPU path: the C port oracle/msda_oracle.c on all host
cores over a bounded sample of the same workload (the Python reference cannot travel to the GPU
box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

LEVELS = [(100, 167), (50, 84), (25, 42), (13, 21)]     # R50, 3x800x1333
BATCH = 2
NQ_PAIR, NQ_VERB = 300, 150
M, D, L, P = 8, 32, 4, 4


def peaks():
    p = {"hbm_gbs": 6650.0, "src": "fallback (B200_PROFILING.md)"}
    try:
        j = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        p = {"hbm_gbs": float(j["hbm_gbs"]), "src": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(float(s[0])) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.samples[0][1])), "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workload: the MSDeformAttn calls of one train step
# ------------------------------------------------------------------------------------------------
def msda_calls():
    S = sum(h * w for h, w in LEVELS)
    return [("enc", S)] * 6 + [("pair_dec", NQ_PAIR)] * 3 + [("verb_dec", NQ_VERB)] * 3, S


def make_inputs_device(seed):
    """One independent input set for the 3 distinct call shapes (device tensors)."""
    from rlipv2_b200 import synth
    S = sum(h * w for h, w in LEVELS)
    sets = {"enc": synth.encoder_inputs(BATCH, LEVELS, seed=seed)}
    for name, lq in (("pair_dec", NQ_PAIR), ("verb_dec", NQ_VERB)):
        v, sh, lsi, loc, attn, gout = synth.random_inputs(BATCH, lq, LEVELS, seed=seed + 100)
        sets[name] = (sets["enc"][0], sh, lsi, loc, attn, gout)      # decoder samples the same memory
    return sets, S


def run_ours(args, rank, world, device):
    from rlipv2_b200 import msda_abi, synth
    from rlipv2_b200.dropin import MultiScaleDeformableAttention as MSDA
    calls, S = msda_calls()
    fwd_e, bwd_e = synth.msda_bytes(BATCH, S, S)
    ncopies = max(2, int(2.2 * 126e6 / fwd_e) + 1)
    sets = [make_inputs_device(seed)[0] for seed in range(ncopies)]

    def step(i):
        s = sets[i % ncopies]
        for name, _ in calls:
            v, sh, lsi, loc, attn, gout = s[name]
            MSDA.ms_deform_attn_forward(v, sh, lsi, loc, attn, 64)
            MSDA.ms_deform_attn_backward(v, sh, lsi, loc, attn, gout, 64)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    l0 = msda_abi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(device.index) as clk:
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (msda_abi.launch_count() - l0)
    # memsets of grad_value are issued by our library too (cudaMemsetAsync), not counted as kernels

    # dominant kernel: encoder backward.  Time it alone (same rotation) for the roofline entry.
    def time_kernel(fn, iters=20):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters * 1e-3

    def enc_bwd(i):
        v, sh, lsi, loc, attn, gout = sets[i % ncopies]["enc"]
        MSDA.ms_deform_attn_backward(v, sh, lsi, loc, attn, gout, 64)

    def enc_fwd(i):
        v, sh, lsi, loc, attn, gout = sets[i % ncopies]["enc"]
        MSDA.ms_deform_attn_forward(v, sh, lsi, loc, attn, 64)

    t_bwd, t_fwd = time_kernel(enc_bwd), time_kernel(enc_fwd)
    pk = peaks()
    roofline = {"bound": "hbm", "kernel": "msda_bwd_d32_l4p4 (encoder call, N=2, S=Lq=22223)",
                "achieved": bwd_e / t_bwd / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": bwd_e / t_bwd / 1e9 / pk["hbm_gbs"], "peak_src": pk["src"],
                "traffic": 345.4e6, "traffic_src": "ncu dram__bytes_read+write, profiles/msda_r01.md",
                "fwd": {"kernel": "msda_fwd_d32_l4p4", "achieved": fwd_e / t_fwd / 1e9,
                        "frac": fwd_e / t_fwd / 1e9 / pk["hbm_gbs"], "us": t_fwd * 1e6},
                "us": t_bwd * 1e6}

    # e2e: the same step through the C-ABI with HOST (pinned) buffers: H2D of every call's inputs,
    # D2H of every call's outputs, inside the timed region.
    host = {k: [t.cpu().pin_memory() for t in v] for k, v in sets[0].items()}
    dev = {k: [torch.empty_like(t) for t in v] for k, v in sets[0].items()}
    h2d = d2h = 0
    outs_host = {}

    def e2e_step():
        nonlocal h2d, d2h
        h2d = d2h = 0
        for name, _ in calls:
            for hs, ds in zip(host[name], dev[name]):
                ds.copy_(hs, non_blocking=True)
                h2d += hs.numel() * hs.element_size()
            v, sh, lsi, loc, attn, gout = dev[name]
            o = MSDA.ms_deform_attn_forward(v, sh, lsi, loc, attn, 64)
            g = MSDA.ms_deform_attn_backward(v, sh, lsi, loc, attn, gout, 64)
            for j, t in enumerate([o] + g):
                key = (name, j)
                if key not in outs_host:
                    outs_host[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                outs_host[key].copy_(t, non_blocking=True)
                d2h += t.numel() * t.element_size()

    e2e_step()
    barrier()
    e0.record()
    n_e2e = max(1, min(args.steps, 5))
    for _ in range(n_e2e):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / n_e2e

    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    return {
        "metric": "images/sec (MSDeformAttn fwd+bwd calls of one RLIPv2-ParSeDA R50 train step)",
        "value": BATCH * world / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "msda_step: 12 MSDeformAttn fwd+bwd calls (6x Lq=S=22223, 3x Lq=300, 3x Lq=150), "
                               "batch 2 x 3x800x1333 per GPU, 8 heads x 32 ch, 4 levels x 4 points",
                   "per_gpu_batch": BATCH, "global_batch": BATCH * world, "l2": "inputs>L2 (rotating input sets)",
                   "parallelism": f"replicas x{world} (op shards by image; no collective)"},
        "clocks": clk.summary(), "gpu_launches": int(launches),
        "e2e": {"value": BATCH * world / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e},
        "roofline": roofline,
    }


def cpu_port_step_time(threads, sample_queries=2048):
    """Reference CPU path (C port) on a bounded sample: one image, `sample_queries` encoder queries,
    forward + backward; scaled to a full step by the query count (cost is linear in queries)."""
    import numpy as np
    from oracle import msda_oracle
    S = sum(h * w for h, w in LEVELS)
    rng = np.random.default_rng(0)
    value = rng.standard_normal((1, S, M, D), dtype=np.float32)
    shapes = np.array(LEVELS, dtype=np.int64)
    lsi = np.concatenate(([0], np.cumsum(shapes.prod(1))[:-1])).astype(np.int64)
    Lq = sample_queries
    loc = rng.random((1, Lq, M, L, P, 2), dtype=np.float32)
    attn = rng.random((1, Lq, M, L, P), dtype=np.float32)
    attn /= attn.sum((-1, -2), keepdims=True)
    gout = rng.standard_normal((1, Lq, M * D), dtype=np.float32)
    msda_oracle.forward(value, shapes, lsi, loc[:, :64], attn[:, :64], threads=threads)      # warm
    t0 = time.perf_counter()
    msda_oracle.forward(value, shapes, lsi, loc, attn, threads=threads)
    # backward port parallelises over images only; run it per-thread on a query slice instead
    t1 = time.perf_counter()
    msda_oracle.backward(value, shapes, lsi, loc[:, :Lq // max(1, threads)], attn[:, :Lq // max(1, threads)],
                         gout[:, :Lq // max(1, threads)], threads=1)
    t2 = time.perf_counter()
    calls, S = msda_calls()
    q_total = BATCH * sum(lq for _, lq in calls)
    per_q = (t1 - t0) / Lq + (t2 - t1) / (Lq // max(1, threads)) / max(1, threads)
    return per_q * q_total, {"fwd_s": t1 - t0, "bwd_slice_s": t2 - t1, "sample_queries": Lq}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    threads = os.cpu_count() or 1
    times = []
    for _ in range(max(1, min(args.steps, 3))):
        t, detail = cpu_port_step_time(threads)
        times.append(t)
    t = min(times)
    val = BATCH / t
    cb = {"value": val, "unit": "images/s", "cores": threads, "kind": "port",
          "sample": f"1 image, {detail['sample_queries']} encoder-shaped queries fwd (all cores) + bwd slice, "
                    "scaled linearly in queries to the 12-call step"}
    return {"impl": "reference", "metric": "images/sec (MSDeformAttn fwd+bwd calls of one RLIPv2-ParSeDA R50 train step)",
            "value": val, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "msda_step (CPU port oracle/msda_oracle.c, bounded sample)", "per_gpu_batch": BATCH},
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        line = run_reference(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return

    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=device)
    line = run_ours(args, rank, world, device)
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            t, detail = cpu_port_step_time(threads)
            line["cpu_baseline"] = {"value": BATCH / t, "unit": "images/s", "cores": threads, "kind": "port",
                                    "sample": f"1 image, {detail['sample_queries']} encoder-shaped queries, scaled "
                                              "linearly to the 12-call step (oracle/msda_oracle.c, pthreads)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
