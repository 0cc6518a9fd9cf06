#!/bin/bash
# which earlier test makes tests/test_train_step_gpu.py::test_graphed_step_matches_eager_step see NaN costs?
TAG=${1:-r02m}
mkdir -p gpurun_out
T=tests/test_train_step_gpu.py::test_graphed_step_matches_eager_step
run() { name=$1; shift; timeout -s KILL 400 python -m pytest "$@" $T -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/${TAG}_$name.log 2>&1; echo "== $name: $(tail -1 gpurun_out/${TAG}_$name.log)"; }
run alone
run small_ops tests/test_small_ops_gpu.py
run pretrain "tests/test_parseda_model.py::test_pretrain_step_golden"
run model_all tests/test_parseda_model.py
run tf32 "tests/test_parseda_model.py::test_full_step_golden_tf32"
run x3tf32 "tests/test_parseda_model.py::test_full_step_golden_3xtf32"
run postproc tests/test_postprocess.py tests/test_msda_proj_gpu.py
run attn_dense tests/test_attn_gpu.py tests/test_dense_gpu.py
