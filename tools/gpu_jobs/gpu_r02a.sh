#!/bin/bash
# Round 2, first GPU call: whole GPU suite without -x, graph-vs-eager inference diagnosis, opt-in A/Bs, configs 3/4.
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout -s KILL 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -15 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python tools/debug_infer_graph.py fp32 > gpurun_out/${TAG}_infer_fp32.log 2>&1
timeout 300 python tools/debug_infer_graph.py tf32 > gpurun_out/${TAG}_infer_tf32.log 2>&1
RLIPV2_VALUE_STREAM=0 RLIPV2_POS_STREAM=0 timeout 300 python tools/debug_infer_graph.py fp32 > gpurun_out/${TAG}_infer_fp32_nostreams.log 2>&1
tail -12 gpurun_out/${TAG}_infer_fp32.log gpurun_out/${TAG}_infer_tf32.log gpurun_out/${TAG}_infer_fp32_nostreams.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline"
timeout 400 $B > gpurun_out/${TAG}_base.json 2> gpurun_out/${TAG}_base.err
RLIPV2_DEVICE_LSAP=1 timeout 400 $B > gpurun_out/${TAG}_device_lsap.json 2> gpurun_out/${TAG}_device_lsap.err
RLIPV2_GN_TOKENS=1 timeout 400 $B > gpurun_out/${TAG}_gn_tokens.json 2> gpurun_out/${TAG}_gn_tokens.err
RLIPV2_SDPA_QUERY_ATTN=1 timeout 400 $B > gpurun_out/${TAG}_sdpa_query.json 2> gpurun_out/${TAG}_sdpa_query.err
RLIPV2_SPLITK_FWD=1 timeout 400 $B > gpurun_out/${TAG}_splitk.json 2> gpurun_out/${TAG}_splitk.err
RLIPV2_ALLREDUCE_OVERLAP=force timeout 400 $B > gpurun_out/${TAG}_overlap_force.json 2> gpurun_out/${TAG}_overlap_force.err
for f in base device_lsap gn_tokens sdpa_query splitk overlap_force; do python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_$f.json").read().strip().splitlines()[-1])
    print("$f", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2), "launches", j["gpu_launches"], "loss", j["final_loss"])
except Exception as e:
    print("$f FAILED", e)
PY
done
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline --pretrain > gpurun_out/${TAG}_pretrain.json 2> gpurun_out/${TAG}_pretrain.err
timeout 700 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline --backbone swin_large --per-gpu-batch 1 > gpurun_out/${TAG}_swin_large.json 2> gpurun_out/${TAG}_swin_large.err
tail -c 300 gpurun_out/${TAG}_pretrain.json gpurun_out/${TAG}_swin_large.json
tail -3 gpurun_out/${TAG}_pretrain.err gpurun_out/${TAG}_swin_large.err
