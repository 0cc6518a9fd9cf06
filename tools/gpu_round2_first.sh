#!/bin/bash
# First GPU call of round 2: everything that was written after round 1's GPU budget was spent, in the order of what a
# failure would invalidate.  One B200.   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh r02a'
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
# 1. the verified suite first (no -x: see every failure), then the unverified files on their own
timeout -s KILL 900 python -m pytest tests -m gpu -q --deselect tests/test_zz1_lsap_gpu.py --deselect tests/test_zz4_swin_backbone.py \
    --deselect tests/test_zz2_overlap_gpu.py --deselect tests/test_zz3_infer_gpu.py --deselect tests/test_zz5_splitk_gpu.py \
    > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
timeout -s KILL 300 python -m pytest tests/test_zz1_lsap_gpu.py -m gpu -q > gpurun_out/${TAG}_pytest_lsap.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_lsap.log
timeout -s KILL 300 python -m pytest tests/test_zz2_overlap_gpu.py tests/test_zz3_infer_gpu.py tests/test_zz5_splitk_gpu.py -m gpu -q > gpurun_out/${TAG}_pytest_overlap_infer.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_overlap_infer.log
timeout -s KILL 600 python -m pytest tests/test_zz4_swin_backbone.py -m gpu -q > gpurun_out/${TAG}_pytest_swin.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_swin.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log gpurun_out/${TAG}_pytest_lsap.log gpurun_out/${TAG}_pytest_overlap_infer.log gpurun_out/${TAG}_pytest_swin.log
# 2. A/B on one box: base, assignment on the device, token-major GroupNorm input projections, both
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline"
timeout 400 $B > gpurun_out/${TAG}_base.json 2> gpurun_out/${TAG}_base.err
RLIPV2_DEVICE_LSAP=1 timeout 400 $B > gpurun_out/${TAG}_device_lsap.json 2> gpurun_out/${TAG}_device_lsap.err
RLIPV2_GN_TOKENS=1 timeout 400 $B > gpurun_out/${TAG}_gn_tokens.json 2> gpurun_out/${TAG}_gn_tokens.err
RLIPV2_DEVICE_LSAP=1 RLIPV2_GN_TOKENS=1 timeout 400 $B > gpurun_out/${TAG}_both.json 2> gpurun_out/${TAG}_both.err
RLIPV2_SDPA_QUERY_ATTN=1 timeout 400 $B > gpurun_out/${TAG}_sdpa_query.json 2> gpurun_out/${TAG}_sdpa_query.err
RLIPV2_SPLITK_FWD=1 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_splitk.json 2> gpurun_out/${TAG}_splitk.err
for f in base device_lsap gn_tokens both sdpa_query splitk; do python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_$f.json").read().strip().splitlines()[-1])
    print("$f", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s launches", j["gpu_launches"], "loss", j["final_loss"])
except Exception as e:
    print("$f FAILED", e)
PY
done
# 3. the other BASELINE configs on the same step (not the headline): config 3 flags, config 4 (Swin-L, batch 1)
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline --pretrain > gpurun_out/${TAG}_pretrain.json 2> gpurun_out/${TAG}_pretrain.err
timeout 700 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline --backbone swin_large --per-gpu-batch 1 > gpurun_out/${TAG}_swin_large.json 2> gpurun_out/${TAG}_swin_large.err
tail -c 400 gpurun_out/${TAG}_pretrain.json gpurun_out/${TAG}_swin_large.json
tail -5 gpurun_out/${TAG}_swin_large.err
