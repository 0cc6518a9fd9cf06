#!/bin/bash
TAG=${1:-bwd2}
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_dense_gpu.py -q --maxfail=30 -k "wgrad or dgrad" > gpurun_out/${TAG}_pytest_dense.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_dense.log
grep -E "passed|failed|^FAILED" gpurun_out/${TAG}_pytest_dense.log | head -20
grep -E "^E       Assertion" gpurun_out/${TAG}_pytest_dense.log | head -8
timeout -s KILL 300 python tools/bwd_gemm_microbench.py > gpurun_out/${TAG}_microbench.jsonl 2> gpurun_out/${TAG}_microbench.err
cut -c1-330 gpurun_out/${TAG}_microbench.jsonl; tail -3 gpurun_out/${TAG}_microbench.err
timeout 300 python tools/profile_train_step.py tf32 > gpurun_out/${TAG}_kernels.txt 2>&1
head -3 gpurun_out/${TAG}_kernels.txt
