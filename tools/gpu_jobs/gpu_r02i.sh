#!/bin/bash
# persistent tcgen05 linear: bit-identity test, micro-benchmark, step A/B; full GPU suite after the capture stream-order fix
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_dense_gpu.py -m gpu -q --tb=short -k "persistent" > gpurun_out/${TAG}_persistent_test.log 2>&1; tail -6 gpurun_out/${TAG}_persistent_test.log
timeout 400 python tools/dense_microbench.py "enc " > gpurun_out/${TAG}_dense_microbench.jsonl 2> gpurun_out/${TAG}_dense_microbench.err
python - <<PY
import json
for l in open("gpurun_out/${TAG}_dense_microbench.jsonl"):
    j = json.loads(l); print(j["shape"][:34].ljust(34), "ours %.1f" % j["ours_us"], "persistent %.1f" % j["ours_persistent_us"] if "ours_persistent_us" in j else "", "cublas %.1f" % j["cublas_us"])
PY
tail -3 gpurun_out/${TAG}_dense_microbench.err
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-roofline"
timeout 400 $B > gpurun_out/${TAG}_base.json 2> gpurun_out/${TAG}_base.err
RLIPV2_DENSE_PERSISTENT=296 timeout 400 $B > gpurun_out/${TAG}_persistent.json 2> gpurun_out/${TAG}_persistent.err
timeout 400 $B > gpurun_out/${TAG}_base2.json 2> gpurun_out/${TAG}_base2.err
RLIPV2_DENSE_PERSISTENT=296 timeout 400 $B > gpurun_out/${TAG}_persistent2.json 2> gpurun_out/${TAG}_persistent2.err
for f in base persistent base2 persistent2; do python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_$f.json").read().strip().splitlines()[-1])
    print("$f", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2), "launches", j["gpu_launches"], "loss", j["final_loss"])
except Exception as e:
    print("$f FAILED", e)
PY
done
tail -3 gpurun_out/${TAG}_persistent.err
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -15 gpurun_out/${TAG}_pytest_gpu.log
