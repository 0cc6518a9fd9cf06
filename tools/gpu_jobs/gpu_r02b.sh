#!/bin/bash
# Round 2, second GPU call: the fused attention kernels' parity tests (own process: a trap must not take the rest down),
# the fixed graphed-inference test, the new train-step / MSDA tests, then the whole suite and a bench run.
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_attn_gpu.py -m gpu -q -x --tb=short > gpurun_out/${TAG}_attn.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_attn.log
tail -40 gpurun_out/${TAG}_attn.log
timeout -s KILL 600 python -m pytest tests/test_attn_gpu.py -m gpu -q --tb=line > gpurun_out/${TAG}_attn_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_attn_all.log
tail -25 gpurun_out/${TAG}_attn_all.log
timeout -s KILL 600 python -m pytest tests/test_zz3_infer_gpu.py tests/test_train_step_gpu.py tests/test_msda_gpu.py -m gpu -q --tb=short > gpurun_out/${TAG}_new.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_new.log
tail -30 gpurun_out/${TAG}_new.log
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=line --deselect tests/test_attn_gpu.py > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -12 gpurun_out/${TAG}_pytest_gpu.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline"
timeout 400 $B > gpurun_out/${TAG}_base.json 2> gpurun_out/${TAG}_base.err
RLIPV2_FUSED_ATTN=0 timeout 400 $B > gpurun_out/${TAG}_noattn.json 2> gpurun_out/${TAG}_noattn.err
for f in base noattn; do python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_$f.json").read().strip().splitlines()[-1])
    print("$f", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2), "launches", j["gpu_launches"], "loss", j["final_loss"])
except Exception as e:
    print("$f FAILED", e)
PY
done
tail -5 gpurun_out/${TAG}_base.err
