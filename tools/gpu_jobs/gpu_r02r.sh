#!/bin/bash
TAG=${1:-r02r}
mkdir -p gpurun_out
T=tests/test_train_step_gpu.py::test_graphed_step_matches_eager_step
A=tests/test_attn_gpu.py; B=tests/test_backbone_gpu.py; D=tests/test_dense_gpu.py
P=tests/test_parseda_model.py; O=tests/test_postprocess.py; S=tests/test_small_ops_gpu.py
run() { name=$1; shift; timeout -s KILL 400 python -m pytest "$@" $T -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/${TAG}_$name.log 2>&1; echo "== $name: $(tail -1 gpurun_out/${TAG}_$name.log)"; }
run attn_g3 $A $P $O $S
run backbone_g3 $B $P $O $S
run dense_g3 $D $P $O $S
run g1_model $A $B $D $P
run g1_post $A $B $D $O
run g1_small $A $B $D $S
