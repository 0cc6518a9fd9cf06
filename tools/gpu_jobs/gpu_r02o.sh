#!/bin/bash
TAG=${1:-r02o}
mkdir -p gpurun_out
T=tests/test_train_step_gpu.py::test_graphed_step_matches_eager_step
run() { name=$1; shift; timeout -s KILL 400 python -m pytest "$@" $T -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/${TAG}_$name.log 2>&1; echo "== $name: $(tail -1 gpurun_out/${TAG}_$name.log)"; }
run backbone tests/test_backbone_gpu.py
run fused tests/test_fused_gpu.py
run msda tests/test_msda_gpu.py
run all_before tests/test_attn_gpu.py tests/test_backbone_gpu.py tests/test_dense_gpu.py tests/test_fused_gpu.py tests/test_msda_gpu.py tests/test_msda_proj_gpu.py tests/test_parseda_model.py tests/test_postprocess.py tests/test_small_ops_gpu.py
