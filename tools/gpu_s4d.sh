#!/bin/bash
# session-4 GPU visit D: decoder value projections on a side stream, encoder-size parameter gradients on the side stream
TAG=${1:-r01s4d}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/${TAG}_pytest.log | head -20
grep -E "^E   " gpurun_out/${TAG}_pytest.log | head -20
run() {
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/${TAG}_${name}.json").read().strip().splitlines()[-1])
    print("${name}", round(j["ms_per_step"],2), "ms/step", round(j["value"],2), "img/s  e2e", round(j["e2e"]["ms_per_step"],2), "loss", j["final_loss"], "launches", j["gpu_launches"])
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-1500:])
PY
}
run base A=1
run no_value_stream RLIPV2_VALUE_STREAM=0
run wgrad_side_all RLIPV2_WGRAD_STREAM_MAX_ROWS=1000000
run base2 A=1
