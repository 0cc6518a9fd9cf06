#!/bin/bash
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_attn_gpu.py -m gpu -q --tb=short > gpurun_out/${TAG}_attn.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_attn.log
tail -25 gpurun_out/${TAG}_attn.log
timeout 300 python tools/attn_microbench.py > gpurun_out/${TAG}_attn_microbench.jsonl 2> gpurun_out/${TAG}_attn_microbench.err
cat gpurun_out/${TAG}_attn_microbench.jsonl; tail -3 gpurun_out/${TAG}_attn_microbench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_ -c 60 --csv --log-file gpurun_out/${TAG}_attn_launches.csv python tools/attn_microbench.py > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_attn_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
for r in rows[1:41]:
    print(r[ki][:60], r[gi], r[vi])
PY
timeout -s KILL 600 python -m pytest tests/test_parseda_model.py tests/test_dense_gpu.py -m gpu -q --tb=short > gpurun_out/${TAG}_model.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_model.log
tail -30 gpurun_out/${TAG}_model.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline"
timeout 400 $B > gpurun_out/${TAG}_base.json 2> gpurun_out/${TAG}_base.err
RLIPV2_FUSED_ATTN=0 timeout 400 $B > gpurun_out/${TAG}_noattn.json 2> gpurun_out/${TAG}_noattn.err
for f in base noattn; do python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_$f.json").read().strip().splitlines()[-1])
    print("$f", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2), "launches", j["gpu_launches"], "loss", j["final_loss"])
except Exception as e:
    print("$f FAILED", e)
PY
done
