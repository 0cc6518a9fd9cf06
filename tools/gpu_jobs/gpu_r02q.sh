#!/bin/bash
TAG=${1:-r02q}
mkdir -p gpurun_out
T=tests/test_train_step_gpu.py::test_graphed_step_matches_eager_step
G1="tests/test_attn_gpu.py tests/test_backbone_gpu.py tests/test_dense_gpu.py"
G2="tests/test_fused_gpu.py tests/test_msda_gpu.py tests/test_msda_proj_gpu.py"
G3="tests/test_parseda_model.py tests/test_postprocess.py tests/test_small_ops_gpu.py"
run() { name=$1; shift; timeout -s KILL 400 python -m pytest "$@" $T -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/${TAG}_$name.log 2>&1; echo "== $name: $(tail -1 gpurun_out/${TAG}_$name.log)"; }
run g1 $G1
run g2 $G2
run g3 $G3
run g12 $G1 $G2
run g23 $G2 $G3
run g13 $G1 $G3
# stream-pool hypothesis: 40 throw-away streams before the test (torch hands out 32 pool streams round-robin)
cat > /tmp/conftest_streams.py <<'PY'
PY
timeout 300 python - <<'PY' > gpurun_out/${TAG}_streams.log 2>&1
import sys, torch, pytest
keep = [torch.cuda.Stream() for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 45)]
for s in keep:
    with torch.cuda.stream(s):
        torch.zeros(8, device="cuda")
torch.cuda.synchronize()
sys.exit(pytest.main(["tests/test_train_step_gpu.py::test_graphed_step_matches_eager_step", "-m", "gpu", "-q", "--tb=line", "-p", "no:cacheprovider"]))
PY
echo "== streams45: $(tail -1 gpurun_out/${TAG}_streams.log)"
