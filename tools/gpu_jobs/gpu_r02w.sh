#!/bin/bash
# r02w: occupancy experiment for the fast MSDeformAttn backward - 3 CTAs per SM (80 registers) against 2 (128), merged and
# unmerged (rlipv2_msda_set_backward_mode 3 / 4, plain op only)
mkdir -p gpurun_out
timeout 100 python tools/msda_microbench.py --cases enc2,enc2init,enc2n025,dec16 --iters 30 --bwd-modes 0,1,3,4 > gpurun_out/r02w_msda_microbench.jsonl 2> gpurun_out/r02w_msda_microbench.err
python - <<PY
import json
for line in open("gpurun_out/r02w_msda_microbench.jsonl"):
    j = json.loads(line)
    print(j["case"][:64], "| fwd", round(j["ours"]["fwd_us"], 1), "| bwd", {m: round(j[f"ours_bwd_mode{m}"]["bwd_us"], 1) for m in (0, 1, 3, 4)})
PY
tail -3 gpurun_out/r02w_msda_microbench.err
