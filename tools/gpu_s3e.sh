#!/bin/bash
TAG=${1:-r01s3e}
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/${TAG}_${name}.json").read().strip().splitlines()[-1])
    print("${name}", round(j["ms_per_step"],2), "ms/step", round(j["value"],2), "img/s  e2e", round(j["e2e"]["ms_per_step"],2), "loss", j["final_loss"], "launches", j["gpu_launches"], j["clocks"])
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-1500:])
PY
}
run clocks200 A=1
run noclocks RLIPV2_BENCH_NO_CLOCKS=1
run clocks100 RLIPV2_BENCH_CLOCKS_MS=100
run clocks200_noflag RLIPV2_FLAG_WAIT=0
run noclocks_noflag RLIPV2_BENCH_NO_CLOCKS=1 RLIPV2_FLAG_WAIT=0
