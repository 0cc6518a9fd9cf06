#!/bin/bash
# r02x: last call of round 2 - the final dispatch of the MSDeformAttn backward (mode 2: merged for encoder-shaped calls,
# 3 CTAs per SM for the plain op's decoder-shaped calls) under the kernel-level and model-level tests, and the final bench line.
mkdir -p gpurun_out
T0=$(date +%s)
timeout -s KILL 140 python -m pytest tests/test_msda_gpu.py tests/test_msda_merge_gpu.py tests/test_msda_proj_gpu.py tests/test_alif_streams.py tests/test_parseda_model.py tests/test_train_step_gpu.py -m gpu -q --tb=short > gpurun_out/r02x_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02x_pytest.log
tail -6 gpurun_out/r02x_pytest.log | cut -c1-300; echo "[$(( $(date +%s) - T0 )) s]"
timeout 90 python bench.py --steps 20 --warmup 5 > gpurun_out/r02x_bench_final.json 2> gpurun_out/r02x_bench_final.err; tail -c 400 gpurun_out/r02x_bench_final.json; tail -2 gpurun_out/r02x_bench_final.err; echo "[$(( $(date +%s) - T0 )) s]"
