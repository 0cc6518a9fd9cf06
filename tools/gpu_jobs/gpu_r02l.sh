#!/bin/bash
TAG=${1:-r02l}
mkdir -p gpurun_out
timeout 300 python tools/debug_lang_stream.py > gpurun_out/${TAG}_lang_stream.log 2>&1; tail -40 gpurun_out/${TAG}_lang_stream.log | cut -c1-220
timeout -s KILL 600 python -m pytest tests/test_dense_gpu.py -m gpu -q --tb=short -k "persistent" > gpurun_out/${TAG}_persistent_test.log 2>&1; tail -6 gpurun_out/${TAG}_persistent_test.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-roofline"
RLIPV2_DENSE_PERSISTENT=0 timeout 400 $B > gpurun_out/${TAG}_nopersist.json 2> gpurun_out/${TAG}_nopersist.err
RLIPV2_DENSE_PERSISTENT=296 timeout 400 $B > gpurun_out/${TAG}_persist.json 2> gpurun_out/${TAG}_persist.err
RLIPV2_DENSE_PERSISTENT=0 timeout 400 $B > gpurun_out/${TAG}_nopersist2.json 2> gpurun_out/${TAG}_nopersist2.err
RLIPV2_DENSE_PERSISTENT=296 timeout 400 $B > gpurun_out/${TAG}_persist2.json 2> gpurun_out/${TAG}_persist2.err
for f in nopersist persist nopersist2 persist2; do python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_$f.json").read().strip().splitlines()[-1])
    print("$f", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2), "loss", j["final_loss"])
except Exception as e:
    print("$f FAILED", e)
PY
done
tail -3 gpurun_out/${TAG}_persist.err
