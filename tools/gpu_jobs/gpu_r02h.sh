#!/bin/bash
# Round 2 evidence run on one B200: whole GPU suite, TMA-staged MSDA variant (timing + ncu), DRAM traffic of the MSDA kernels,
# launch list of the step (device LSAP: no host flag under ncu's serialisation), final bench lines, reference arm smoke.
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=line > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -8 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python tools/msda_tma_experiment.py > gpurun_out/${TAG}_msda_tma.jsonl 2> gpurun_out/${TAG}_msda_tma.err; cat gpurun_out/${TAG}_msda_tma.jsonl; tail -3 gpurun_out/${TAG}_msda_tma.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 4 -c 2 -o gpurun_out/${TAG}_msda_tma_prof python tools/msda_tma_experiment.py --once > gpurun_out/${TAG}_msda_tma_prof.log 2>&1
timeout 600 bash tools/ncu_traffic.sh > gpurun_out/${TAG}_traffic.log 2>&1; tail -2 gpurun_out/${TAG}_traffic.log
timeout 300 python tools/attn_microbench.py > gpurun_out/${TAG}_attn_microbench.jsonl 2>/dev/null; cat gpurun_out/${TAG}_attn_microbench.jsonl
cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json 2>/dev/null
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 900 gpurun_out/${TAG}_bench_reference.json; tail -3 gpurun_out/${TAG}_bench_reference.err
RLIPV2_DEVICE_LSAP=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --graph-profiling node \
    --csv --log-file gpurun_out/${TAG}_launches_train_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_families.py gpurun_out/${TAG}_launches_train_step.csv 2 > gpurun_out/${TAG}_launches_summary.md 2>&1; head -12 gpurun_out/${TAG}_launches_summary.md
