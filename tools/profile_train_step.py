"""torch.profiler breakdown of the ParSeDA train step (kernel time by name), one GPU."""
import os
import sys

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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        ts.step_device(samples, targets, text)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
print("max memory allocated GB:", torch.cuda.max_memory_allocated() / 1e9)
