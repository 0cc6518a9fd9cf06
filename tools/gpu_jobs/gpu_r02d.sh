#!/bin/bash
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_parseda_model.py -m gpu -q --tb=short > gpurun_out/${TAG}_model.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_model.log
tail -12 gpurun_out/${TAG}_model.log
timeout 400 python tools/dense_microbench.py > gpurun_out/${TAG}_dense_microbench.jsonl 2> gpurun_out/${TAG}_dense_microbench.err
python - <<PY
import json
for l in open("gpurun_out/${TAG}_dense_microbench.jsonl"):
    j = json.loads(l); print(j["shape"][:34].ljust(34), "ours %.1f" % j["ours_us"], "splitk %.1f" % j["ours_splitk_us"] if "ours_splitk_us" in j else "", "cublas %.1f" % j["cublas_us"])
PY
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline"
timeout 400 $B > gpurun_out/${TAG}_base.json 2> gpurun_out/${TAG}_base.err
RLIPV2_SPLITK_FWD=0 timeout 400 $B > gpurun_out/${TAG}_nosplitk.json 2> gpurun_out/${TAG}_nosplitk.err
for f in base nosplitk; do python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_$f.json").read().strip().splitlines()[-1])
    print("$f", round(j["ms_per_step"], 3), "ms/step", round(j["value"], 2), "img/s e2e", round(j["e2e"]["value"], 2), "launches", j["gpu_launches"], "loss", j["final_loss"])
except Exception as e:
    print("$f FAILED", e)
PY
done
# ncu --set full of the attention kernels (tensor-pipe utilisation, registers, stalls)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 45 -c 15 -o gpurun_out/${TAG}_attn_prof python tools/attn_profile_target.py > gpurun_out/${TAG}_attn_prof.log 2>&1
ls -la gpurun_out/${TAG}_attn_prof.ncu-rep
# compute-sanitizer: memcheck over the kernel-level parity tests of the five libraries, racecheck over the shared-memory
# heavy ones (mbarrier / TMEM hand-offs, MSDA prep records, LSAP state)
export RLIPV2_SANITIZER=1
timeout -s KILL 1500 compute-sanitizer --tool memcheck --error-exitcode 97 --log-file gpurun_out/${TAG}_memcheck.txt \
    python -m pytest tests/test_attn_gpu.py tests/test_msda_gpu.py tests/test_msda_proj_gpu.py tests/test_dense_gpu.py tests/test_zz5_splitk_gpu.py \
    tests/test_fused_gpu.py tests/test_small_ops_gpu.py tests/test_zz1_lsap_gpu.py -m gpu -q -x --tb=line \
    -k "not full_size and not full_step and not gradcheck and not graphed" > gpurun_out/${TAG}_memcheck_pytest.log 2>&1; echo "memcheck exit $?" >> gpurun_out/${TAG}_memcheck_pytest.log
tail -5 gpurun_out/${TAG}_memcheck_pytest.log; tail -8 gpurun_out/${TAG}_memcheck.txt
timeout -s KILL 900 compute-sanitizer --tool racecheck --error-exitcode 97 --log-file gpurun_out/${TAG}_racecheck.txt \
    python -m pytest tests/test_attn_gpu.py tests/test_msda_gpu.py tests/test_zz1_lsap_gpu.py tests/test_dense_gpu.py -m gpu -q -x --tb=line \
    -k "(forward_matches or backward_matches or golden or wgrad or dgrad or lsap) and not full" > gpurun_out/${TAG}_racecheck_pytest.log 2>&1; echo "racecheck exit $?" >> gpurun_out/${TAG}_racecheck_pytest.log
tail -5 gpurun_out/${TAG}_racecheck_pytest.log; tail -8 gpurun_out/${TAG}_racecheck.txt
