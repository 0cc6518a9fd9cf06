"""Kernel-level time breakdown of one eager ParSeDA train step (torch.profiler, CUDA kernels only)."""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlipv2_b200 import train_step  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "tf32"
ts = train_step.ParSeDATrainStep(device="cuda", precision=precision, seed=0)
text = train_step.synthetic_text(170, 85)
images_h, targets_h = train_step.synthetic_batch(2, 800, 1333, seed=0)
samples, targets = ts.to_device(images_h, targets_h)
for _ in range(3):
    ts.step_device(samples, targets, text)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    ts.step_device(samples, targets, text)
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name][0] += 1
        agg[e.name][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"total GPU kernel time per step: {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches")
for name, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(os.environ.get('TOPK', '400'))]:
    print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}%  n={n:4d}  {name[:150]}")

if len(sys.argv) > 2 and sys.argv[2] == "big":
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    print("---- kernels >= 40 us in launch order")
    for e in evs:
        t = e.device_time if hasattr(e, "device_time") else e.cuda_time
        if t >= 40:
            print(f"{t:8.1f} us  {e.name[:140]}")

# which aten ops (with which operand shapes) own the device time: attribution of the copies / elementwise kernels
print("---- aten ops by (name, input shapes), self device time")
ka = prof.key_averages(group_by_input_shape=True)
rows = sorted(ka, key=lambda e: -e.self_device_time_total)[:int(os.environ.get("TOPOPS", "120"))]
for e in rows:
    print(f"{e.self_device_time_total / 1e3:8.3f} ms  n={e.count:4d}  {e.key[:40]:40s} {str(e.input_shapes)[:150]}")
