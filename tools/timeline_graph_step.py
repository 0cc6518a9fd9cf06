"""Timeline of one graph-replayed ParSeDA train step (torch.profiler / CUPTI kernel records).

Answers what the kernel-sum tables cannot: how much of the step the GPU is idle or running one tiny
kernel at a time.  Prints the union-busy time over all streams, per-stream busy time, the idle gaps,
a 1-ms bucket view (busy fraction + the kernels that own the bucket) and the kernel table of the
replayed step.  usage: python tools/timeline_graph_step.py [bucket_ms]
"""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlipv2_b200 import train_step  # noqa: E402

bucket_ms = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
ts = train_step.GraphedParSeDATrainStep(device="cuda", precision="tf32", seed=0)
text = train_step.synthetic_text(170, 85)
images_h, targets_h = train_step.synthetic_batch(2, 800, 1333, seed=0)
ts.capture(images_h, targets_h, text, warmup=2)
for _ in range(3):
    ts.replay()
torch.cuda.synchronize()
ts.check()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ts.replay()
    torch.cuda.synchronize()

evs = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        dur = e.device_time if hasattr(e, "device_time") else e.cuda_time
        evs.append((e.time_range.start, e.time_range.start + dur, e.name, getattr(e, "device_resource_id", 0)))
evs.sort()
if os.environ.get("TIMELINE_DUMP"):
    with open(os.environ["TIMELINE_DUMP"], "w") as f:
        f.write("start_us,dur_us,stream,kernel\n")
        for s_, e_, n_, st_ in evs:
            f.write(f"{s_ - evs[0][0]:.1f},{e_ - s_:.1f},{st_},\"{n_[:110]}\"\n")
t0, t1 = evs[0][0], max(e[1] for e in evs)
span = t1 - t0
print(f"kernels {len(evs)}  span {span / 1e3:.2f} ms  kernel-time sum {sum(e[1] - e[0] for e in evs) / 1e3:.2f} ms")


def union(intervals):
    tot, cur_s, cur_e = 0.0, None, None
    gaps = []
    for s, e in sorted(intervals):
        if cur_e is None:
            cur_s, cur_e = s, e
        elif s <= cur_e:
            cur_e = max(cur_e, e)
        else:
            gaps.append((cur_e, s))
            tot += cur_e - cur_s
            cur_s, cur_e = s, e
    if cur_e is not None:
        tot += cur_e - cur_s
    return tot, gaps


busy, gaps = union([(s, e) for s, e, _, _ in evs])
print(f"union busy {busy / 1e3:.2f} ms  idle inside the span {(span - busy) / 1e3:.2f} ms over {len(gaps)} gaps")
hist = defaultdict(lambda: [0, 0.0])
for a, b in gaps:
    g = b - a
    k = "<2us" if g < 2 else "<5us" if g < 5 else "<20us" if g < 20 else "<100us" if g < 100 else ">=100us"
    hist[k][0] += 1
    hist[k][1] += g
for k in ("<2us", "<5us", "<20us", "<100us", ">=100us"):
    print(f"   gaps {k:8s} n={hist[k][0]:5d}  {hist[k][1] / 1e3:7.3f} ms")
print("largest gaps (offset ms, length us, next kernel):")
for a, b in sorted(gaps, key=lambda g: g[0] - g[1])[:12]:
    nxt = next((n for s, _, n, _ in evs if s >= b), "?")
    print(f"   @{(a - t0) / 1e3:7.2f} ms  {b - a:8.1f} us  -> {nxt[:90]}")
per_stream = defaultdict(list)
for s, e, _, st in evs:
    per_stream[st].append((s, e))
for st, iv in sorted(per_stream.items(), key=lambda x: -len(x[1])):
    b, _ = union(iv)
    print(f"stream {st}: {len(iv)} kernels, busy {b / 1e3:.2f} ms, from {(min(i[0] for i in iv) - t0) / 1e3:.2f} to {(max(i[1] for i in iv) - t0) / 1e3:.2f} ms")

# kernel size classes: how much of the step is made of tiny kernels
cls = defaultdict(lambda: [0, 0.0])
for s, e, _, _ in evs:
    d = e - s
    k = "<3us" if d < 3 else "<10us" if d < 10 else "<50us" if d < 50 else "<200us" if d < 200 else ">=200us"
    cls[k][0] += 1
    cls[k][1] += d
for k in ("<3us", "<10us", "<50us", "<200us", ">=200us"):
    print(f"   kernels {k:8s} n={cls[k][0]:5d}  {cls[k][1] / 1e3:7.3f} ms")

print(f"---- {bucket_ms} ms buckets: busy fraction (union), kernels started, top kernels by time")
nb = int(span / 1e3 / bucket_ms) + 1
for b in range(nb):
    lo, hi = t0 + b * bucket_ms * 1e3, t0 + (b + 1) * bucket_ms * 1e3
    inside = [(max(s, lo), min(e, hi), n) for s, e, n, _ in evs if e > lo and s < hi]
    bz, _ = union([(s, e) for s, e, _ in inside])
    agg = defaultdict(float)
    for s, e, n in inside:
        agg[n] += e - s
    top = sorted(agg.items(), key=lambda x: -x[1])[:3]
    started = sum(1 for s, _, _, _ in evs if lo <= s < hi)
    print(f"{b * bucket_ms:6.1f} ms  busy {bz / (bucket_ms * 1e3):4.2f}  n={started:4d}  " +
          " | ".join(f"{t:5.0f}us {n[:48]}" for n, t in top))

agg = defaultdict(lambda: [0, 0.0])
for s, e, n, _ in evs:
    agg[n][0] += 1
    agg[n][1] += e - s
tot = sum(v[1] for v in agg.values())
print("---- kernel table of the replayed step")
for name, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(os.environ.get("TOPK", "120"))]:
    print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}%  n={n:4d}  {name[:150]}")
