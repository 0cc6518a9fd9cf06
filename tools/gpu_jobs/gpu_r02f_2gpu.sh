#!/bin/bash
# overlap variant alone, with a stack dump if it hangs at teardown; then the single-GPU tests that changed
TAG=${1:-r02f}
N=${2:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-roofline"
RLIPV2_BENCH_FAULT_S=150 RLIPV2_ALLREDUCE_OVERLAP=1 timeout -s KILL 280 $T > gpurun_out/${TAG}_${N}gpu_overlap.json 2> gpurun_out/${TAG}_${N}gpu_overlap.err
echo "overlap exit $?"
grep -n "File \|Thread\|Current thread" gpurun_out/${TAG}_${N}gpu_overlap.err | head -40
tail -c 300 gpurun_out/${TAG}_${N}gpu_overlap.json
RLIPV2_BENCH_FAULT_S=200 timeout -s KILL 280 $T > gpurun_out/${TAG}_${N}gpu_sharded.json 2> gpurun_out/${TAG}_${N}gpu_sharded.err
echo "sharded exit $?"
python - <<PY
import json
for f in ("overlap", "sharded"):
    try:
        j = json.loads(open("gpurun_out/${TAG}_${N}gpu_%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2))
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout -s KILL 600 python -m pytest tests/test_train_step_gpu.py tests/test_parseda_model.py tests/test_attn_gpu.py -m gpu -q --tb=short > gpurun_out/${TAG}_tests.log 2>&1; tail -15 gpurun_out/${TAG}_tests.log
