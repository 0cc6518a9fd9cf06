#!/bin/bash
# tcgen05 backward GEMMs: parity tests (own process, hard timeout), micro-benchmark, step A/B, kernel table
TAG=${1:-bwd}
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_dense_gpu.py -q --maxfail=20 > gpurun_out/${TAG}_pytest_dense.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_dense.log
tail -25 gpurun_out/${TAG}_pytest_dense.log
timeout -s KILL 300 python tools/bwd_gemm_microbench.py > gpurun_out/${TAG}_microbench.jsonl 2> gpurun_out/${TAG}_microbench.err
cat gpurun_out/${TAG}_microbench.jsonl; tail -3 gpurun_out/${TAG}_microbench.err
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
timeout 300 python tools/profile_train_step.py tf32 big > gpurun_out/${TAG}_kernels.txt 2>&1
head -3 gpurun_out/${TAG}_kernels.txt
