import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlipv2_b200 import train_step
ts = train_step.GraphedParSeDATrainStep(device="cuda", precision="tf32", seed=0)
text = train_step.synthetic_text(170, 85)
images_h, targets_h = train_step.synthetic_batch(2, 800, 1333, seed=0)
ts.flag_timeout_s = 3.0
os.environ["RLIPV2_FLAG_TIMEOUT_S"] = "3"
ts.capture(images_h, targets_h, text, warmup=2)
torch.cuda.synchronize()
print("after capture: d_err", ts.d_err.item(), "d_seq", ts.d_seq.item(), "flag_seq", ts.flag_seq, flush=True)
for it in range(3):
    t0 = time.perf_counter()
    ts.graph_a.replay(); ts.done_a.record()
    t1 = time.perf_counter()
    ts.graph_b.replay()
    t2 = time.perf_counter()
    ts.done_a.synchronize()
    t3 = time.perf_counter()
    ts._solve_assignment_host()
    t4 = time.perf_counter()
    ts.flag_seq += 1; ts.np_flag[0] = ts.flag_seq
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    print(f"replay {it}: launchA {1e3*(t1-t0):.2f} launchB {1e3*(t2-t1):.2f} waitA {1e3*(t3-t2):.2f} solve {1e3*(t4-t3):.2f} rest {1e3*(t5-t4):.2f} ms | d_err {ts.d_err.item()} d_seq {ts.d_seq.item()} loss {float(ts.s_loss):.4f}", flush=True)
