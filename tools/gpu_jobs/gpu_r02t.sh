#!/bin/bash
TAG=${1:-r02t}
mkdir -p gpurun_out
timeout 600 python tools/debug_nan_sequence.py 2>&1 | grep -v "^tests/\|passed\|warnings\|Warning\|^$\|^  " | tail -12 | cut -c1-600
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -8 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<PY
import json
j = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("bench", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2), "launches", j["gpu_launches"])
print("roofline", j["roofline"]["frac"], j["roofline"]["us"], j["roofline"]["traffic"], "alif", j["roofline_alif_tensor"]["frac"], j["roofline_alif_tensor"]["us_per_layer"])
PY
