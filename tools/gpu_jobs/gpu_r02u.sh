#!/bin/bash
# Round 2, last GPU call(s): merged-reduction MSDeformAttn backward (RLIPV2_MSDA_BWD_MERGE) and two-stream ALIF
# (RLIPV2_ALIF_STREAMS): parity, micro-benchmark, step A/B, the model-level tests under both switches, evidence (ncu counters,
# DRAM traffic, bench line), then BASELINE configs 3 / 4 on one GPU.  Ordered by priority: the call's time limit may cut the tail.
TAG=${1:-r02u}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout -s KILL 300 python -m pytest tests/test_msda_merge_gpu.py tests/test_alif_streams.py -m gpu -q --tb=short > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_new.log
tail -6 gpurun_out/${TAG}_pytest_new.log | cut -c1-300; el new tests
timeout 150 python tools/msda_microbench.py --cases enc2,enc2init,enc2n025,dec16 --iters 30 --bwd-modes 0,1 > gpurun_out/${TAG}_msda_microbench.jsonl 2> gpurun_out/${TAG}_msda_microbench.err
python - <<PY
import json
for line in open("gpurun_out/${TAG}_msda_microbench.jsonl"):
    try:
        j = json.loads(line)
        print(j["case"][:60], "| fwd", round(j["ours"]["fwd_us"], 1), "| bwd mode0", round(j["ours_bwd_mode0"]["bwd_us"], 1), "mode1", round(j["ours_bwd_mode1"]["bwd_us"], 1))
    except Exception as e:
        print("microbench line failed", e)
PY
el microbench
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline"
run() { name=$1; shift; env "$@" timeout 200 $B > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2), "loss", j.get("final_loss"))
except Exception as e:
    print("$name FAILED", e)
PY
}
run base X=0
run merge RLIPV2_MSDA_BWD_MERGE=1
run both RLIPV2_MSDA_BWD_MERGE=1 RLIPV2_ALIF_STREAMS=1
run alif RLIPV2_ALIF_STREAMS=1
el step A/B
RLIPV2_MSDA_BWD_MERGE=1 RLIPV2_ALIF_STREAMS=1 timeout -s KILL 420 python -m pytest tests/test_parseda_model.py tests/test_train_step_gpu.py tests/test_zz3_infer_gpu.py -m gpu -q --tb=short > gpurun_out/${TAG}_pytest_switched.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_switched.log
tail -6 gpurun_out/${TAG}_pytest_switched.log | cut -c1-300; el model-level tests under both switches
RLIPV2_MSDA_BWD_MERGE=1 timeout 200 bash tools/ncu_traffic.sh > gpurun_out/${TAG}_traffic.log 2>&1; tail -c 700 gpurun_out/${TAG}_traffic.log; echo
cp gpurun_out/ncu_traffic.json gpurun_out/${TAG}_ncu_traffic_merge.json 2>/dev/null; el traffic
cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json 2>/dev/null
RLIPV2_MSDA_BWD_MERGE=1 RLIPV2_ALIF_STREAMS=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_final.json 2> gpurun_out/${TAG}_bench_final.err; tail -c 1200 gpurun_out/${TAG}_bench_final.json; tail -2 gpurun_out/${TAG}_bench_final.err; el final bench
RLIPV2_MSDA_BWD_MERGE=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:msda_bwd -s 2 -c 1 -o gpurun_out/${TAG}_msda_bwd_merge_prof python tools/msda_profile_target.py --case enc2 --iters 3 > gpurun_out/${TAG}_msda_bwd_prof.log 2>&1
timeout 60 python tools/ncu_summary.py gpurun_out/${TAG}_msda_bwd_merge_prof.ncu-rep > gpurun_out/${TAG}_msda_bwd_merge_ncu.txt 2>&1; head -40 gpurun_out/${TAG}_msda_bwd_merge_ncu.txt | cut -c1-200; el ncu full
RLIPV2_MSDA_BWD_MERGE=1 RLIPV2_ALIF_STREAMS=1 timeout 200 python bench.py --pretrain --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/${TAG}_config3_1gpu.json 2> gpurun_out/${TAG}_config3_1gpu.err; tail -c 600 gpurun_out/${TAG}_config3_1gpu.json; tail -2 gpurun_out/${TAG}_config3_1gpu.err; el config 3
RLIPV2_MSDA_BWD_MERGE=1 RLIPV2_ALIF_STREAMS=1 timeout 240 python bench.py --backbone swin_large --per-gpu-batch 1 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/${TAG}_config4_1gpu.json 2> gpurun_out/${TAG}_config4_1gpu.err; tail -c 600 gpurun_out/${TAG}_config4_1gpu.json; tail -2 gpurun_out/${TAG}_config4_1gpu.err; el config 4
run base2 X=0
run both2 RLIPV2_MSDA_BWD_MERGE=1 RLIPV2_ALIF_STREAMS=1
el extras
