#!/bin/bash
# A/B of the step-level switches (one bench line each).  usage: bash tools/gpu_ab.sh <tag>
TAG=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/${TAG}_${name}.json").read().strip().splitlines()[-1])
    print("${name}", round(j["ms_per_step"],2), "ms/step", round(j["value"],2), "img/s  e2e", round(j["e2e"]["ms_per_step"],2), "loss", j["final_loss"])
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-1500:])
PY
}
run all_on A=1
run no_gather RLIPV2_GATHER_GRADS=0
run no_prologue RLIPV2_MSDA_FUSED_PROLOGUE=0
run no_lang_stream RLIPV2_LANG_STREAM=0
run nhwc_nofusedconv RLIPV2_FUSED_CONV=0
