#!/bin/bash
# session-4 GPU visit A: parity of the new paths, small-grid GEMM sweep, step A/B
TAG=${1:-r01s4a}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/${TAG}_pytest.log | head -20
grep -E "^E   " gpurun_out/${TAG}_pytest.log | head -12
timeout 300 python tools/dense_microbench.py > gpurun_out/${TAG}_dense_microbench.jsonl 2> gpurun_out/${TAG}_dense_microbench.err
python - <<PY
import json
for l in open("gpurun_out/${TAG}_dense_microbench.jsonl"):
    try:
        j=json.loads(l)
        print(j["shape"][:28].ljust(28), " ".join(f"{k[5:-3]}={j[k]:.1f}" for k in j if k.startswith("ours_mode")), f"default={j['ours_us']:.1f} cublas={j['cublas_us']:.1f}")
    except Exception as e:
        print("bad line", e)
PY
tail -3 gpurun_out/${TAG}_dense_microbench.err
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
run no_ref4 RLIPV2_MSDA_FUSED_PROLOGUE_REF4=0
run small0 RLIPV2_DENSE_SMALL_MODE=0
run small2 RLIPV2_DENSE_SMALL_MODE=2
run ffn_hybrid RLIPV2_FFN_BWD=hybrid
TIMELINE_DUMP=gpurun_out/${TAG}_kernels.csv timeout 300 python tools/timeline_graph_step.py 1.0 > gpurun_out/${TAG}_timeline.txt 2>&1
head -12 gpurun_out/${TAG}_timeline.txt
