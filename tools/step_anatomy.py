"""Where the graphed step's time goes: forward graph, host assignment, backward graph, and the GPU idle time
between the two graphs.  Host timestamps (perf_counter) and CUDA events, no profiler attached.
usage: python tools/step_anatomy.py [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlipv2_b200 import train_step  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ts = train_step.GraphedParSeDATrainStep(device="cuda", precision="tf32", seed=0)
text = train_step.synthetic_text(170, 85)
images_h, targets_h = train_step.synthetic_batch(2, 800, 1333, seed=0)
ts.capture(images_h, targets_h, text, warmup=2)
for _ in range(3):
    ts.replay()
torch.cuda.synchronize()

ev = lambda: torch.cuda.Event(enable_timing=True)

# GPU-side anatomy from %globaltimer stamps captured inside the graphs (RLIPV2_STAMPS=1):
#   0 forward graph starts   1 forward graph done (costs copied to the host)   2 backward graph's head reached
#   3 flag seen + indices copied   4 optimizer step done
assert ts.stamps is not None, "run with RLIPV2_STAMPS=1"
rows, host = [], []
torch.cuda.synchronize()
s0, s1 = ev(), ev()
s0.record()
prev_end = None
for _ in range(steps):
    t0 = time.perf_counter()
    ts.replay()
    host.append(time.perf_counter() - t0)
    torch.cuda.synchronize()          # stamps are read per step: the steps below do not pipeline on the host
    st = ts.stamps.tolist()
    rows.append([st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3], st[4] - st[0]])
s1.record()
torch.cuda.synchronize()
names = ["forward graph", "A end -> B head", "flag wait + index copies", "backward graph", "total"]
for i, n in enumerate(names):
    v = sorted(r[i] for r in rows[2:])
    print(f"{n:28s} median {v[len(v) // 2] / 1e6:7.3f} ms   min {v[0] / 1e6:7.3f}  max {v[-1] / 1e6:7.3f}")
print(f"host time inside replay(): median {sorted(host)[len(host) // 2] * 1e3:.2f} ms")

# the free-running loop (no per-step sync), as bench.py times it
torch.cuda.synchronize()
s0.record()
for _ in range(steps):
    ts.replay()
s1.record()
torch.cuda.synchronize()
print(f"free-running step: {s0.elapsed_time(s1) / steps:.2f} ms   flag_wait={ts.flag_wait}")
ts.check()
