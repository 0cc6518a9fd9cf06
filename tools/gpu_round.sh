#!/bin/bash
# One GPU-box visit: parity tests, bench line, kernel breakdown, ncu launch list of the bench command.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [skip-tests]
TAG=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
  tail -3 gpurun_out/${TAG}_pytest_gpu.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json
timeout 300 python tools/profile_train_step.py tf32 big > gpurun_out/${TAG}_kernels.txt 2>&1
timeout 300 python tools/timeline_graph_step.py 1.0 > gpurun_out/${TAG}_timeline.txt 2>&1
head -30 gpurun_out/${TAG}_timeline.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" \
   --graph-profiling node --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
wc -l gpurun_out/${TAG}_launches.csv
