#!/bin/bash
# Round 2, final GPU call: the library defaults are now merged MSDeformAttn backward for encoder-shaped calls
# (rlipv2_msda_set_backward_mode 2) and two-stream ALIF (RLIPV2_ALIF_STREAMS=1).  Parity of the new paths, the model-level tests
# with the defaults, step A/B against each switch turned off, evidence (DRAM traffic, bench line, ncu counters, launch list),
# BASELINE configs 3 / 4 on one GPU.  Ordered by priority.
TAG=${1:-r02v}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout -s KILL 200 python -m pytest tests/test_msda_merge_gpu.py tests/test_alif_streams.py -m gpu -q --tb=short > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_new.log
tail -5 gpurun_out/${TAG}_pytest_new.log | cut -c1-300; el new tests
timeout -s KILL 330 python -m pytest tests/test_parseda_model.py tests/test_train_step_gpu.py tests/test_zz3_infer_gpu.py tests/test_zz1_lsap_gpu.py tests/test_attn_gpu.py tests/test_msda_gpu.py tests/test_msda_proj_gpu.py tests/test_zz2_overlap_gpu.py -m gpu -q --tb=short > gpurun_out/${TAG}_pytest_models.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_models.log
tail -8 gpurun_out/${TAG}_pytest_models.log | cut -c1-300; el model-level tests with the new defaults
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline"
run() { name=$1; shift; env "$@" timeout 150 $B > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2), "loss", j.get("final_loss"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/${TAG}_$name.err").read()[-600:])
PY
}
run default1 X=0
run nomerge1 RLIPV2_MSDA_BWD_MERGE=0
run noalif1 RLIPV2_ALIF_STREAMS=0
run default2 X=0
run nomerge2 RLIPV2_MSDA_BWD_MERGE=0
run noalif2 RLIPV2_ALIF_STREAMS=0
el step A/B
timeout 150 bash tools/ncu_traffic.sh > gpurun_out/${TAG}_traffic.log 2>&1; tail -c 500 gpurun_out/${TAG}_traffic.log; echo
cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json 2>/dev/null; el traffic
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_final.json 2> gpurun_out/${TAG}_bench_final.err; tail -c 1500 gpurun_out/${TAG}_bench_final.json; tail -2 gpurun_out/${TAG}_bench_final.err; el final bench
timeout 150 ncu --set full --clock-control none --import-source on -k regex:msda_bwd -s 2 -c 1 -o gpurun_out/${TAG}_msda_bwd_merge_prof python tools/msda_profile_target.py --case enc2 --iters 3 > gpurun_out/${TAG}_msda_bwd_prof.log 2>&1
timeout 60 python tools/ncu_summary.py gpurun_out/${TAG}_msda_bwd_merge_prof.ncu-rep --all-stalls > gpurun_out/${TAG}_msda_bwd_merge_ncu.txt 2>&1; head -34 gpurun_out/${TAG}_msda_bwd_merge_ncu.txt | cut -c1-200; el ncu full
timeout 60 python tools/msda_microbench.py --cases enc2,enc2init,enc2n025,dec16 --iters 30 --bwd-modes 0,1,2 > gpurun_out/${TAG}_msda_microbench.jsonl 2> gpurun_out/${TAG}_msda_microbench.err; cut -c1-50,150-700 gpurun_out/${TAG}_msda_microbench.jsonl; el microbench
timeout 120 python bench.py --pretrain --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/${TAG}_config3_1gpu.json 2> gpurun_out/${TAG}_config3_1gpu.err; tail -c 700 gpurun_out/${TAG}_config3_1gpu.json; tail -2 gpurun_out/${TAG}_config3_1gpu.err; el config 3
timeout 150 python bench.py --backbone swin_large --per-gpu-batch 1 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/${TAG}_config4_1gpu.json 2> gpurun_out/${TAG}_config4_1gpu.err; tail -c 700 gpurun_out/${TAG}_config4_1gpu.json; tail -2 gpurun_out/${TAG}_config4_1gpu.err; el config 4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; el smoke
RLIPV2_DEVICE_LSAP=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --graph-profiling node \
    --csv --log-file gpurun_out/${TAG}_launches_train_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_families.py gpurun_out/${TAG}_launches_train_step.csv 2 > gpurun_out/${TAG}_launches_summary.md 2>&1; head -12 gpurun_out/${TAG}_launches_summary.md; el launch list
