#!/bin/bash
# DRAM traffic per launch of the MSDeformAttn encoder-call kernels (the `roofline.traffic` entries of bench.py), measured with
# ncu on the GPU box and written to profiles/ncu_traffic.json, which bench.py reads (it reports null when the file is absent -
# never a stale constant).      gpurun --timeout 600 -- 'bash tools/ncu_traffic.sh'   then copy gpurun_out/ncu_traffic.json
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:msda_ -s 2 -c 4 --csv \
    --log-file gpurun_out/ncu_traffic.csv python tools/msda_profile_target.py --case enc2 --iters 3 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:msda_ -s 2 -c 4 --csv \
    --log-file gpurun_out/ncu_traffic_dec16.csv python tools/msda_profile_target.py --case dec16 --iters 3 > /dev/null 2>&1
python - <<'PY'
import csv, json, subprocess

def collect(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    acc = {}
    for r in rows[1:]:
        kind = "bwd" if "msda_bwd" in r[ki] else "fwd"
        v = float(r[vi].replace(",", ""))
        u = r[ui].lower()
        if "byte" in u:
            v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
        acc.setdefault(kind, {}).setdefault(r[mi], []).append(v)
    out = {}
    for kind, m in acc.items():
        n = len(m["dram__bytes_read.sum"])
        out[kind] = {"dram_read_bytes": sum(m["dram__bytes_read.sum"]) / n, "dram_write_bytes": sum(m["dram__bytes_write.sum"]) / n,
                     "launches": n}
        out[kind]["traffic_bytes"] = out[kind]["dram_read_bytes"] + out[kind]["dram_write_bytes"]
    return out

res = {"how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum per launch (tools/ncu_traffic.sh), mean over the profiled launches",
       "encoder_call_N2_S22223": collect("gpurun_out/ncu_traffic.csv"),
       "config5_N16_Lq300": collect("gpurun_out/ncu_traffic_dec16.csv"),
       "gpu": subprocess.run(["nvidia-smi", "--query-gpu=name", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()}
json.dump(res, open("gpurun_out/ncu_traffic.json", "w"), indent=1)
print(json.dumps(res))
PY
