#!/bin/bash
TAG=${1:-bwd3}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed|^FAILED" gpurun_out/${TAG}_pytest_gpu.log | head -20
grep -E "^E   " gpurun_out/${TAG}_pytest_gpu.log | head -8
timeout -s KILL 300 python tools/bwd_gemm_microbench.py > gpurun_out/${TAG}_microbench.jsonl 2> gpurun_out/${TAG}_microbench.err
cut -c1-200 gpurun_out/${TAG}_microbench.jsonl; tail -3 gpurun_out/${TAG}_microbench.err
run() {
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
run base A=1
run own_bwd RLIPV2_OWN_BWD=1
run no_own_wgrad RLIPV2_OWN_WGRAD=0
