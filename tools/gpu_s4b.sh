#!/bin/bash
# session-4 GPU visit B: parity of the paired MSDA backward / fused box loss / fused clip, MSDA sweep, step A/B
TAG=${1:-r01s4b}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/${TAG}_pytest.log | head -20
grep -E "^E   " gpurun_out/${TAG}_pytest.log | head -20
timeout 300 python tools/msda_microbench.py --iters 30 --cases enc2,enc2init,enc2n025,encrand2,dec16 > gpurun_out/${TAG}_msda_microbench.jsonl 2> gpurun_out/${TAG}_msda_microbench.err
python - <<PY
import json
for l in open("gpurun_out/${TAG}_msda_microbench.jsonl"):
    try:
        j=json.loads(l); o=j["ours"]
        print(j["case"][:60].ljust(60), f"fwd {o['fwd_us']:.0f}us ({o['fwd_frac']:.3f})  bwd v0 {o.get('bwd_variant0_us',0):.0f}us ({o.get('bwd_variant0_frac',0):.3f})  v1 {o.get('bwd_variant1_us',0):.0f}us ({o.get('bwd_variant1_frac',0):.3f})")
    except Exception as e:
        print("bad line", e)
PY
tail -3 gpurun_out/${TAG}_msda_microbench.err
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
run bwd_paired RLIPV2_MSDA_BWD_VARIANT=1
run serial_losses RLIPV2_PARALLEL_LOSSES=0
run torch_box_loss RLIPV2_FUSED_BOX_LOSS=0
run per_level_heads RLIPV2_STACKED_HEADS=0
run old_clip_zero RLIPV2_FUSED_CLIP=0 RLIPV2_EARLY_ZERO=0
TIMELINE_DUMP=gpurun_out/${TAG}_kernels.csv timeout 300 python tools/timeline_graph_step.py 1.0 > gpurun_out/${TAG}_timeline.txt 2>&1
head -4 gpurun_out/${TAG}_timeline.txt
