#!/bin/bash
# N-GPU A/B of the data-parallel step variants: one all-reduce behind the backward (round 1), sharded optimizer step
# (reduce-scatter + 1/N AdamW + all-gather; the round-2 default), marker-driven overlapped all-reduce.  Each run under its own
# timeout so that a collective-order bug cannot hang the box.     gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_jobs/gpu_r02e_2gpu.sh r02e 2'
TAG=${1:-r02e}
N=${2:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-roofline"
RLIPV2_SHARD_OPTIMIZER=0 timeout -s KILL 300 $T > gpurun_out/${TAG}_${N}gpu_allreduce.json 2> gpurun_out/${TAG}_${N}gpu_allreduce.err
timeout -s KILL 300 $T > gpurun_out/${TAG}_${N}gpu_sharded.json 2> gpurun_out/${TAG}_${N}gpu_sharded.err
RLIPV2_ALLREDUCE_OVERLAP=1 timeout -s KILL 300 $T > gpurun_out/${TAG}_${N}gpu_overlap.json 2> gpurun_out/${TAG}_${N}gpu_overlap.err
for f in allreduce sharded overlap; do python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_${N}gpu_$f.json").read().strip().splitlines()[-1])
    print("$f N=$N", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2), "loss", j["final_loss"])
except Exception as e:
    print("$f FAILED", e)
PY
done
tail -3 gpurun_out/${TAG}_${N}gpu_sharded.err gpurun_out/${TAG}_${N}gpu_overlap.err
