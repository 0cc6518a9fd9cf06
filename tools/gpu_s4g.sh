#!/bin/bash
# session-4 GPU visit G: short-sequence attention + library GEMMs for the text tower
TAG=${1:-r01s4g}
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
run hf_eager_attn RLIPV2_SHORT_ATTN=0
run text_tcgen05 RLIPV2_TEXT_LIBRARY_GEMM=0
run base2 A=1
TIMELINE_DUMP=gpurun_out/${TAG}_kernels.csv timeout 300 python tools/timeline_graph_step.py 1.0 > gpurun_out/${TAG}_timeline.txt 2>&1
head -4 gpurun_out/${TAG}_timeline.txt
