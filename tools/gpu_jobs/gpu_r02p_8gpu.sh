#!/bin/bash
# BASELINE configs 3 and 4 on 8 GPUs (and the headline config 2 with the final defaults), one box
TAG=${1:-r02p}
N=${2:-8}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 15 --warmup 3 --no-roofline"
export RLIPV2_BENCH_FAULT_S=240
timeout -s KILL 300 $T > gpurun_out/${TAG}_${N}gpu_config2.json 2> gpurun_out/${TAG}_${N}gpu_config2.err
timeout -s KILL 300 $T --pretrain > gpurun_out/${TAG}_${N}gpu_config3.json 2> gpurun_out/${TAG}_${N}gpu_config3.err
timeout -s KILL 300 $T --backbone swin_large --per-gpu-batch 1 > gpurun_out/${TAG}_${N}gpu_config4.json 2> gpurun_out/${TAG}_${N}gpu_config4.err
python - <<PY
import json
for f in ("config2", "config3", "config4"):
    try:
        j = json.loads(open("gpurun_out/${TAG}_${N}gpu_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "N=$N", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2), "global batch", j["config"]["global_batch"])
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/${TAG}_${N}gpu_config4.err
