#!/bin/bash
# 2-GPU A/B of the overlapped gradient all-reduce (RLIPV2_ALLREDUCE_OVERLAP=1), each run under its own timeout so that a
# collective-order bug cannot hang the box.   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_round2_overlap.sh r02b'
TAG=${1:-r02b}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-roofline"
timeout -s KILL 300 $T > gpurun_out/${TAG}_2gpu_base.json 2> gpurun_out/${TAG}_2gpu_base.err
RLIPV2_ALLREDUCE_OVERLAP=1 timeout -s KILL 300 $T > gpurun_out/${TAG}_2gpu_overlap.json 2> gpurun_out/${TAG}_2gpu_overlap.err
echo "overlap exit $?"
for f in 2gpu_base 2gpu_overlap; do tail -c 700 gpurun_out/${TAG}_$f.json; echo; done
tail -5 gpurun_out/${TAG}_2gpu_overlap.err
