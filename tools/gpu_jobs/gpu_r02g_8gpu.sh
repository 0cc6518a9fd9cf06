#!/bin/bash
# 8-GPU A/B/C of the data-parallel step variants on ONE box, 15 timed steps each
TAG=${1:-r02g}
N=${2:-8}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 15 --warmup 3 --no-roofline"
export RLIPV2_BENCH_FAULT_S=200
RLIPV2_SHARD_OPTIMIZER=0 timeout -s KILL 260 $T > gpurun_out/${TAG}_${N}gpu_allreduce.json 2> gpurun_out/${TAG}_${N}gpu_allreduce.err
timeout -s KILL 260 $T > gpurun_out/${TAG}_${N}gpu_sharded.json 2> gpurun_out/${TAG}_${N}gpu_sharded.err
RLIPV2_ALLREDUCE_OVERLAP=1 timeout -s KILL 260 $T > gpurun_out/${TAG}_${N}gpu_overlap.json 2> gpurun_out/${TAG}_${N}gpu_overlap.err
timeout -s KILL 260 $T > gpurun_out/${TAG}_${N}gpu_sharded2.json 2> gpurun_out/${TAG}_${N}gpu_sharded2.err
python - <<PY
import json
for f in ("allreduce", "sharded", "overlap", "sharded2"):
    try:
        j = json.loads(open("gpurun_out/${TAG}_${N}gpu_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "N=$N", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2), "loss", round(j["final_loss"], 3))
    except Exception as e:
        print(f, "FAILED", e)
PY
