"""bench.py - headline measurement (driver contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train_step|msda_step]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads
  train_step (default) BASELINE config 2: one full RLIPv2-ParSeDA R50 optimisation step (phase A +
             phase B + SetCriterionHOI/matcher + backward + clip + AdamW) on a per-GPU batch of two
             synthetic 3x800x1333 images, 300 queries, 256 label strings (170 objects + 'no objects' +
             85 relations), random-init weights, fp32 storage with TF32 tensor-core contractions
             (what the reference's pinned torch 1.10 does by default).  images/sec = global batch / step.
  msda_step  the 12 MSDeformAttn forward+backward calls of that step in isolation (micro-benchmark).

Every rank runs the same per-GPU batch (weak scaling; gradients all-reduce over NCCL).  `value`: K
steps with the batch already resident in HBM, CUDA events, barrier + synchronize on both sides, max
over ranks.  `e2e`: the same step through `ParSeDATrainStep.step()` with the batch in pinned host
memory (H2D inside the timed region) and the loss read back to the host every step.
L2 hygiene: a train step streams > 2 GB of activations per image, far beyond the 126 MB L2
("l2": "working set >> L2").

`--impl reference` times the reference's CPU path through the oracle port (oracle/) on the host
cores, on a bounded sample of the same workload.
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
NUM_QUERIES = 300
M, D, L, P = 8, 32, 4, 4
METRIC = "images/sec RLIPv2-ParSeDA R50 800px train step"


TRAINABLE_PARAMS_R50 = 212737454          # parameters that receive a gradient in the BASELINE config-2 step (flat buffer)


def train_config(batch, world, nparams, graphs=True, backbone=None, pretrain=False):
    """the `config` object of a train_step line; the reference arm prints the SAME object for the same workload"""
    return {"workload": ("train_step: RLIPv2-ParSeDA R50 HICO-DET fine-tune step (BASELINE config 2), batch 2 x 3x800x1333 "
                         "per GPU, 300 queries, 256 label strings, AdamW, clip 0.1, random-init weights") if backbone is None else
                        (f"train_step: RLIPv2-ParSeDA {backbone}, {'relational pre-train' if pretrain else 'HICO-DET fine-tune'} "
                         f"flags, batch {batch} x 3x800x1333 per GPU, 300 queries, 256 label strings, AdamW, clip 0.1, random-init weights"),
            "per_gpu_batch": batch, "global_batch": batch * world, "trainable_params": nparams,
            "l2": "working set >> L2 (activations > 2 GB per image)",
            "execution": "2 CUDA graphs per step + host LSAP" if graphs else "eager",
            "parallelism": (f"dp{world} (flat-gradient NCCL all-reduce in graph)" if graphs else
                            f"dp{world} (DDP static_graph, NCCL)") if world > 1 else "dp1"}


def peaks():
    p = {"hbm_gbs": 6650.0, "src": "fallback (B200_PROFILING.md)"}
    try:
        j = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        p = {"hbm_gbs": float(j["hbm_gbs"]), "src": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def __enter__(self):
        if os.environ.get("RLIPV2_BENCH_NO_CLOCKS") == "1":      # diagnostic only: how much the polling costs
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", os.environ.get("RLIPV2_BENCH_CLOCKS_MS", "200")], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.25)
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) >= 6:
                self.samples.append(f)

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()

    def summary(self):
        good = []
        for s in self.samples:
            try:
                good.append((int(float(s[0])), int(float(s[1])), s[2:6]))
            except ValueError:
                pass
        if not good:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(g[0] for g in good)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(g[2][i].lower().startswith("active") for g in good)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": good[0][1], "reasons": reasons, "samples": len(sm)}


def dist_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def timed(fn, steps, world):
    """CUDA-event time of exactly `steps` calls, barrier + synchronize on both sides -> ms per step."""
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("timed")          # ncu --nvtx --nvtx-include "timed/" selects this region
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.nvtx.range_pop()
    barrier(world)
    return e0.elapsed_time(e1) / steps


def max_over_ranks(values, device, world):
    t = torch.tensor(values, device=device, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return [float(x) for x in t]


# ------------------------------------------------------------------------------------------------
# dominant own kernel: MSDeformAttn encoder call, timed alone for the roofline entry
# ------------------------------------------------------------------------------------------------
def measured_traffic():
    """DRAM bytes per launch from the last ncu run (tools/ncu_traffic.sh -> profiles/ncu_traffic.json); None when absent"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return None


def msda_roofline(device):
    from rlipv2_b200 import synth
    from rlipv2_b200.dropin import MultiScaleDeformableAttention as MSDA
    S = sum(h * w for h, w in LEVELS)
    fwd_b, bwd_b = synth.msda_bytes(BATCH, S, S)
    ncopies = 3                                              # 3 x 273 MB of operands >> L2
    sets = [synth.encoder_inputs(BATCH, LEVELS, seed=s) for s in range(ncopies)]

    def t(fn, iters=15):
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

    tb = t(lambda i: MSDA.ms_deform_attn_backward(*sets[i % ncopies][:5], sets[i % ncopies][5], 64))
    tf = t(lambda i: MSDA.ms_deform_attn_forward(*sets[i % ncopies][:5], 64))
    del sets
    pk = peaks()
    # BASELINE config 5 (the isolated micro-benchmark the metric names): levels {100,50,25,13}^2, 300 queries, batch 16,
    # inputs as models/ops/test.py:32-40, three rotating input sets (3 x 218 MB of value maps >> L2)
    config5 = None
    try:
        c5_levels = [(100, 100), (50, 50), (25, 25), (13, 13)]
        c5_S = sum(h * w for h, w in c5_levels)
        c5_f, c5_b = synth.msda_bytes(16, c5_S, 300)
        c5_sets = [synth.random_inputs(16, 300, c5_levels, seed=3 + s) for s in range(ncopies)]
        c5_tb = t(lambda i: MSDA.ms_deform_attn_backward(*c5_sets[i % ncopies][:5], c5_sets[i % ncopies][5], 64), iters=30)
        c5_tf = t(lambda i: MSDA.ms_deform_attn_forward(*c5_sets[i % ncopies][:5], 64), iters=30)
        del c5_sets
        tr5 = (measured_traffic() or {}).get("config5_N16_Lq300", {})
        config5 = {"workload": "BASELINE config 5: 4 levels {100,50,25,13}^2, Lq = 300, 8 heads x 4 points, N = 16",
                   "traffic": {k: v.get("traffic_bytes") for k, v in tr5.items()} or None,
                   "fwd_us": c5_tf * 1e6, "bwd_us": c5_tb * 1e6, "fwd_gbs": c5_f / c5_tf / 1e9, "bwd_gbs": c5_b / c5_tb / 1e9,
                   "fwd_bwd_gbs": (c5_f + c5_b) / (c5_tf + c5_tb) / 1e9,
                   "fwd_bwd_frac": (c5_f + c5_b) / (c5_tf + c5_tb) / 1e9 / pk["hbm_gbs"],
                   "algorithmic_mb": {"fwd": c5_f / 1e6, "bwd": c5_b / 1e6}}
    except Exception as e:                       # an extra, never allowed to take the headline entry down
        config5 = {"error": repr(e)}
    tr = (measured_traffic() or {}).get("encoder_call_N2_S22223", {})
    return {"config5": config5,
            "bound": "hbm", "kernel": "msda_bwd_d32_l4p4 (encoder call, N=2, S=Lq=22223): 273.1 MB algorithmic / launch",
            "achieved": bwd_b / tb / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": bwd_b / tb / 1e9 / pk["hbm_gbs"],
            "peak_src": pk["src"], "traffic": tr.get("bwd", {}).get("traffic_bytes"),
            "traffic_src": "dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/ncu_traffic.json (tools/ncu_traffic.sh)"
                           if tr else "not measured (profiles/ncu_traffic.json absent)",
            "algorithmic_bytes": bwd_b,
            "us": tb * 1e6,
            "fwd": {"kernel": "msda_fwd_d32_l4p4: 159.3 MB algorithmic / launch", "achieved": fwd_b / tf / 1e9,
                    "frac": fwd_b / tf / 1e9 / pk["hbm_gbs"], "us": tf * 1e6, "algorithmic_bytes": fwd_b,
                    "traffic": tr.get("fwd", {}).get("traffic_bytes")}}


def alif_tensor_roofline(device):
    """One whole ALIF fusion layer on the tcgen05 kernels (forward, batch 2: Tv = 273 image tokens, Tl = 256 labels,
    E = 2048 = 8 heads x 256): the six projections (linear_tf32_kernel) AND the bidirectional attention core
    (attn_fwd_kernel x 2: S = Q K^T in tensor memory, both softmaxes, two P V products) = SURVEY 8d's 4.13 GFLOP per image
    and layer.  Achieved TF32 FLOP/s vs the tensor peak; the driver's MEASURED_PEAKS.json holds the dense bf16 figure and
    TF32 runs at half the bf16 rate on this part (B200_PROFILING.md: 2.25 vs 1.1 PFLOP/s nominal): peak_tf32 = bf16 / 2."""
    from rlipv2_b200 import attn_abi, dense_abi
    B, Tv, Tl, E, H = BATCH, 273, 256, 2048, 8
    g = torch.Generator(device=device).manual_seed(0)
    mk = lambda *s: torch.randn(*s, device=device, generator=g)
    v, l = mk(B, Tv, 256), mk(B, Tl, 768)
    W = {n: (mk(o, i) * i ** -0.5, mk(o)) for n, o, i in (("q", E, 256), ("k", E, 768), ("vv", E, 256), ("vl", E, 768),
                                                       ("ov", 256, E), ("ol", 768, E))}
    splitk = os.environ.get("RLIPV2_SPLITK_FWD", "1") != "0"

    def lin(x, name):
        w, b = W[name]
        x2 = x.reshape(-1, x.shape[-1])
        sp = dense_abi.splitk_splits(x2.shape[0], w.shape[0], w.shape[1]) if splitk else 1
        y = dense_abi.linear_splitk_tf32(x2, w, b, sp) if sp > 1 else dense_abi.linear_tf32(x2, w, b, 0)
        return y.view(*x.shape[:-1], w.shape[0])

    from rlipv2_b200 import alif as alif_mod, streams
    two = alif_mod._ALIF_STREAMS                       # issue the layer the way the model does (alif.py::_two_stream_call)

    def layer():
        if not two:
            q, k, vv, vl = lin(v, "q"), lin(l, "k"), lin(v, "vv"), lin(l, "vl")
            ov, _, _ = attn_abi.forward(q, k, vl, H, None, 256 ** -0.5, 0.0, None, 0)
            ol, _, _ = attn_abi.forward(k, q, vv, H, None, 256 ** -0.5, 0.0, None, 1)
            return lin(ov, "ov"), lin(ol, "ol")
        cur, side = torch.cuda.current_stream(device), streams.get(device, "alif")
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            k, vl = lin(l, "k"), lin(l, "vl")
        q, vv = lin(v, "q"), lin(v, "vv")
        cur.wait_stream(side)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            ol, _, _ = attn_abi.forward(k, q, vv, H, None, 256 ** -0.5, 0.0, None, 1)
            o_l = lin(ol, "ol")
        ov, _, _ = attn_abi.forward(q, k, vl, H, None, 256 ** -0.5, 0.0, None, 0)
        o_v = lin(ov, "ov")
        cur.wait_stream(side)
        return o_v, o_l

    for _ in range(3):
        layer()
    torch.cuda.synchronize()
    # the kernels take a few us each - less than the python / ctypes launch path - so the layer is timed the way the train
    # step issues it: as nodes of a CUDA graph (10 layers' worth per replay)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(10):
                layer()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 50 * 1e-3
    proj = sum(2.0 * m * n * k for m, n, k in ((B * Tv, E, 256), (B * Tl, E, 768), (B * Tv, E, 256), (B * Tl, E, 768),
                                               (B * Tv, 256, E), (B * Tl, 768, E)))
    attn = 3 * 2.0 * B * Tv * Tl * E                      # S = Q K^T once per direction is counted once + 2 P V
    flops = proj + attn
    peak = 1590.0 / 2
    src = "fallback (B200_PROFILING.md) / 2"
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]) / 2
        src = "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 = half the bf16 rate)"
    except Exception:
        pass
    return {"bound": "tensor",
            "kernel": "one ALIF fusion layer forward, batch 2: 6 x linear_tf32_kernel"
                      + (" / split-K gemm_tf32_kernel for K >= 768" if splitk else "")
                      + " + 2 x attn_fwd_kernel (tcgen05.mma kind::tf32; scores and probabilities in tensor memory)",
            "achieved": flops / t / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": flops / t / 1e12 / peak, "peak_src": src,
            "us_per_layer": t * 1e6, "flops_per_layer": flops, "flops_per_image_layer": flops / B,
            "ncu": "profiles/attn_r02_ncu.txt, profiles/dense_r01_final_alif_ncu.txt (sm__pipe_tensor_cycles_active per kernel); "
                   "512-546 row problems fill 16-112 of 148 SMs: latency-bound, not pipe-bound",
            "streams": 2 if two else 1,
            "timing": "CUDA-graph replay of the 8 launches x 10, CUDA events"
                      + (" (label chain beside the image chain, as alif.py issues it)" if two else "")}


# ------------------------------------------------------------------------------------------------
# workload: full train step
# ------------------------------------------------------------------------------------------------
def run_train_step(args, rank, world, device):
    from rlipv2_b200 import attn_abi, dense, dense_abi, fused_abi, lsap_abi, msda_abi, train_step
    text = train_step.synthetic_text(170, 85)
    batch = args.per_gpu_batch or BATCH
    images_h, targets_h = train_step.synthetic_batch(batch, 800, 1333, seed=rank)
    model_args = None
    if args.backbone != "resnet50" or args.pretrain:
        # other BASELINE configs on the same step (not the headline line): config 3 = --pretrain (relational pre-training
        # flags), config 4 = --backbone swin_large --per-gpu-batch 1 (drop_path_rate 0.5)
        from rlipv2_b200 import models
        extra = dict(hoi=False, cross_modal_pretrain=True, pseudo_verb=True) if args.pretrain else {}
        model_args = models.default_args(device=str(device), num_queries=NUM_QUERIES, synthetic_text_encoder=True,
                                         backbone=args.backbone, drop_path_rate=0.5 if "swin" in args.backbone else 0.2,
                                         **extra)
    own = lambda: (msda_abi.launch_count() + dense_abi.launch_count() + fused_abi.launch_count()
                   + lsap_abi.launch_count() + attn_abi.launch_count())
    loss = None
    if args.graphs:
        ts = train_step.GraphedParSeDATrainStep(args=model_args, device=str(device), precision=args.precision, seed=0)
        ts.capture(images_h, targets_h, text, warmup=max(1, args.warmup - 1))
        per_step = ts.own_launches_per_step            # this repo's kernel nodes in the two captured graphs

        def step(i):
            nonlocal loss
            loss = ts.replay()

        h_loss = [torch.zeros(1).pin_memory() for _ in range(2)]
        loss_ev = [None, None]
        e2e_losses = []

        def e2e(i):
            # Every step, inside the timed region: this step's batch and label strings come from pinned host memory through
            # prefetch() (issued during the previous step: double-buffered H2D; the strings are tokenised on the host like the
            # reference does every step, dab_deformable/deformable_transformer.py:497), and the step's loss goes back to the
            # host (4 bytes D2H into a pinned slot) - read one step late, so that the host never stalls the GPU; the last
            # one is read after the loop, inside the timed region (e2e_finish).
            if not ts._prefetched:
                ts.prefetch(images_h, targets_h, text)
            loss_dev = ts.step()
            slot = i & 1
            h_loss[slot].copy_(loss_dev.reshape(1), non_blocking=True)
            loss_ev[slot] = torch.cuda.Event()
            loss_ev[slot].record()
            ts.prefetch(images_h, targets_h, text)          # next step's batch + label tokens: staged while this step computes
            prev = slot ^ 1
            if loss_ev[prev] is not None:
                loss_ev[prev].synchronize()
                e2e_losses.append(float(h_loss[prev]))

        def e2e_finish(i):
            slot = (i - 1) & 1
            loss_ev[slot].synchronize()
            e2e_losses.append(float(h_loss[slot]))
    else:
        ts = train_step.ParSeDATrainStep(args=model_args, device=str(device), precision=args.precision, seed=0)
        samples, targets = ts.to_device(images_h, targets_h)
        per_step = None

        def step(i):
            nonlocal loss
            loss = ts.step_device(samples, targets, text)

        def e2e(i):
            float(ts.step(images_h, targets_h, text))

        e2e_finish = None

    for i in range(args.warmup):
        step(i)
    l0 = own()
    with ClockSampler(device.index) as clk:
        ms = timed(step, args.steps, world)
    launches = per_step * args.steps if per_step is not None else own() - l0
    final_loss = float(loss)
    h2d = images_h.numel() * 4 + sum(v.numel() * v.element_size() for t in targets_h for v in t.values())
    if args.graphs:
        h2d += sum(ts.s_tok[k].numel() * ts.s_tok[k].element_size() for k in ("input_ids", "attention_mask"))
    e2e(0)
    n_e2e = max(2, min(args.steps, 10))

    def e2e_timed(i):
        e2e(i + 1)
        if e2e_finish is not None and i == n_e2e - 1:
            e2e_finish(i + 2)                            # the last step's loss is read before the clock stops

    ms_e2e = timed(e2e_timed, n_e2e, world)
    if args.graphs:
        ts.check()                                   # no replay's host-flag wait timed out
    ms, ms_e2e = max_over_ranks([ms, ms_e2e], device, world)
    nparams = sum(p.numel() for p in ts.params)
    line = {
        "metric": METRIC if model_args is None else METRIC.replace("R50", args.backbone) + (" (pre-train flags)" if args.pretrain else ""),
        "value": batch * world / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 storage, " + ("tf32 tensor-core products" if dense.matmul_precision() == "tf32" else "fp32 products"),
        "data": "synthetic",
        "config": train_config(batch, world, nparams, args.graphs, args.backbone if model_args is not None else None,
                               args.pretrain),
        "clocks": clk.summary(), "gpu_launches": int(launches), "final_loss": final_loss,
        "e2e": {"value": batch * world / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
    }
    if rank == 0 and not args.no_roofline:
        del ts, step, e2e
        torch.cuda.empty_cache()
        line["roofline"] = msda_roofline(device)
        line["roofline_alif_tensor"] = alif_tensor_roofline(device)
    return line


# ------------------------------------------------------------------------------------------------
# workload: the MSDeformAttn calls of one train step (micro-benchmark)
# ------------------------------------------------------------------------------------------------
def run_msda_step(args, rank, world, device):
    from rlipv2_b200 import msda_abi, synth
    from rlipv2_b200.dropin import MultiScaleDeformableAttention as MSDA
    S = sum(h * w for h, w in LEVELS)
    calls = [("enc", S)] * 6 + [("pair_dec", NUM_QUERIES)] * 3 + [("verb_dec", NUM_QUERIES // 2)] * 3
    sets = []
    for seed in range(3):
        d = {"enc": synth.encoder_inputs(BATCH, LEVELS, seed=seed)}
        for name, lq in (("pair_dec", NUM_QUERIES), ("verb_dec", NUM_QUERIES // 2)):
            v, sh, lsi, loc, attn, gout = synth.random_inputs(BATCH, lq, LEVELS, seed=seed + 100)
            d[name] = (d["enc"][0], sh, lsi, loc, attn, gout)
        sets.append(d)

    def step(i):
        s = sets[i % len(sets)]
        for name, _ in calls:
            v, sh, lsi, loc, attn, gout = s[name]
            MSDA.ms_deform_attn_forward(v, sh, lsi, loc, attn, 64)
            MSDA.ms_deform_attn_backward(v, sh, lsi, loc, attn, gout, 64)

    for i in range(args.warmup):
        step(i)
    l0 = msda_abi.launch_count()
    with ClockSampler(device.index) as clk:
        ms = timed(step, args.steps, world)
    launches = msda_abi.launch_count() - l0
    (ms,) = max_over_ranks([ms], device, world)
    line = {"metric": "images/sec (MSDeformAttn fwd+bwd calls of one RLIPv2-ParSeDA R50 train step)",
            "value": BATCH * world / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "msda_step: 12 MSDeformAttn fwd+bwd calls (6x Lq=S=22223, 3x Lq=300, 3x Lq=150), batch 2",
                       "per_gpu_batch": BATCH, "l2": "inputs>L2 (rotating input sets)", "parallelism": f"replicas x{world}"},
            "clocks": clk.summary(), "gpu_launches": int(launches)}
    if rank == 0:
        del sets
        torch.cuda.empty_cache()
        line["roofline"] = msda_roofline(device)
    return line


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference(workload, budget_s=25.0, steps=1, warmup=0):
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    if workload == "train_step":
        from oracle import parseda_oracle
        return parseda_oracle.time_train_step_sample(threads=threads, budget_s=budget_s, steps=steps, warmup=warmup)
    from oracle import msda_oracle_bench
    return msda_oracle_bench.time_msda_step_sample(threads=threads)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train_step", choices=["train_step", "msda_step"])
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--backbone", default="resnet50", help="train_step only: e.g. swin_large (BASELINE config 4; not the headline)")
    ap.add_argument("--per-gpu-batch", type=int, default=0, help="train_step only: images per GPU (default 2)")
    ap.add_argument("--pretrain", action="store_true", help="train_step only: --cross_modal_pretrain --pseudo_verb flags (config 3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="A/B runs: skip the kernel micro-benchmarks")
    ap.add_argument("--no-graphs", dest="graphs", action="store_false", help="eager step instead of CUDA graphs")
    args = ap.parse_args()
    rank, world, local = dist_info()
    if os.environ.get("RLIPV2_BENCH_FAULT_S"):          # diagnostics: dump every thread's stack and exit if the run hangs
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["RLIPV2_BENCH_FAULT_S"]), exit=True)

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference(args.workload, steps=args.steps, warmup=args.warmup)
        if args.workload == "train_step":
            # whole optimisation steps of the CPU port on the host cores (oracle/parseda_oracle.py time_train_step_sample):
            # `steps` = the steps really timed, ms_per_step = the measured time of one of them, value = images / second
            ms = cb["step_s"] * 1e3
            cfg = train_config(BATCH, 1, TRAINABLE_PARAMS_R50)
        else:
            ms = BATCH / cb["value"] * 1e3
            cfg = {"workload": "msda_step: 12 MSDeformAttn fwd+bwd calls (6x Lq=S=22223, 3x Lq=300, 3x Lq=150), batch 2",
                   "per_gpu_batch": BATCH, "l2": "inputs>L2 (rotating input sets)", "parallelism": f"replicas x{world}"}
        line = {"impl": "reference", "metric": METRIC if args.workload == "train_step" else "images/sec (MSDeformAttn fwd+bwd calls of one RLIPv2-ParSeDA R50 train step)",
                "value": cb["value"], "unit": "images/s", "n_gpus": world, "steps": cb.get("timed_steps", args.steps),
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    args.warmup = max(args.warmup, 3)
    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=device)
    line = (run_train_step if args.workload == "train_step" else run_msda_step)(args, rank, world, device)
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_reference(args.workload)
            except Exception as e:       # the baseline is reported, never allowed to hide the GPU number
                line["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        # no destroy_process_group(): with collectives captured on a side (communication) stream it waits forever on work
        # objects that only exist inside the CUDA graphs (stack dump of both ranks: gpurun_out/r02f_2gpu_overlap.err); the
        # measurement is complete and printed, leave without running the NCCL teardown
        os._exit(0)


if __name__ == "__main__":
    main()
