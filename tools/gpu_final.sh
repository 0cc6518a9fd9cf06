#!/bin/bash
# Round-end evidence run on one B200: parity tests, smoke, the bench line (both arms), micro-benchmarks against the
# reference's own CUDA kernel, the ncu launch list of the bench command and ncu --set full captures of the top kernels.
TAG=${1:-r01final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1
tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -c 600 gpurun_out/${TAG}_bench_reference.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3500 gpurun_out/${TAG}_bench.json
timeout 400 python tools/msda_microbench.py --ref --iters 50 > gpurun_out/${TAG}_msda_microbench.jsonl 2> gpurun_out/${TAG}_msda_microbench.err
timeout 300 python tools/dense_microbench.py > gpurun_out/${TAG}_dense_microbench.jsonl 2> gpurun_out/${TAG}_dense_microbench.err
TIMELINE_DUMP=gpurun_out/${TAG}_kernels.csv timeout 300 python tools/timeline_graph_step.py 1.0 > gpurun_out/${TAG}_timeline.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_ -c 2 -f -o gpurun_out/${TAG}_msda_enc \
   python tools/msda_profile_target.py --case enc2 --iters 1 > gpurun_out/${TAG}_ncu_msda.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tf32 -c 6 -f -o gpurun_out/${TAG}_alif_linear \
   python tools/dense_microbench.py ALIF > gpurun_out/${TAG}_ncu_alif.log 2>&1
# RLIPV2_FLAG_WAIT=0: under ncu every graph node is replayed alone, so a backward graph parked on the host flag would sit
# out its 10 s timeout per step (the r01final capture took 15 minutes that way); with the flag wait off the host solves
# the assignment before it launches the backward graph - same kernels, same launch list.
RLIPV2_FLAG_WAIT=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" \
   --graph-profiling node --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/${TAG}_ncu_bench.log 2>&1
wc -l gpurun_out/${TAG}_launches.csv
ls -la gpurun_out | tail -20
