#!/bin/bash
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_dense_gpu.py -m gpu -q --tb=short -k "persistent" > gpurun_out/${TAG}_persistent_test.log 2>&1; tail -4 gpurun_out/${TAG}_persistent_test.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-roofline"
for rep in 1 2; do
RLIPV2_DENSE_PERSISTENT=0 timeout 400 $B > gpurun_out/${TAG}_off$rep.json 2> gpurun_out/${TAG}_off$rep.err
RLIPV2_DENSE_PERSISTENT=296 timeout 400 $B > gpurun_out/${TAG}_lin$rep.json 2> gpurun_out/${TAG}_lin$rep.err
RLIPV2_DENSE_PERSISTENT=296 RLIPV2_DENSE_PERSISTENT_DGRAD=1 timeout 400 $B > gpurun_out/${TAG}_lindgrad$rep.json 2> gpurun_out/${TAG}_lindgrad$rep.err
RLIPV2_DENSE_PERSISTENT=1000 timeout 400 $B > gpurun_out/${TAG}_linwide$rep.json 2> gpurun_out/${TAG}_linwide$rep.err
done
for f in off1 lin1 lindgrad1 linwide1 off2 lin2 lindgrad2 linwide2; do python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_$f.json").read().strip().splitlines()[-1])
    print("$f", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2))
except Exception as e:
    print("$f FAILED", e)
PY
done
