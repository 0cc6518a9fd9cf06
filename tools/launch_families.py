"""Family table of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`): whose kernels the step's time goes to.
    python tools/launch_families.py gpurun_out/launches.csv [steps] > profiles/launches_r02_summary.md"""
import csv
import sys

OWN = ("msda_", "linear_tf32_kernel", "gemm_tf32_kernel", "attn_fwd_kernel", "attn_bwd_ds_kernel", "attn_bgemm3_kernel",
       "add_layernorm", "layernorm_bwd", "relu_bwd_colsum", "rowmask_bwd_colsum", "adamw_kernel", "box_pair_loss", "box_refine",
       "sine_embed", "short_attn", "gn_tok", "gather_chunks", "wait_host_flag", "stamp_kernel", "lsap_kernel")


def family(name):
    if any(k in name for k in OWN):
        return "own (this repo's sm_100a kernels)"
    if "cutlass" in name or "gemm" in name.lower() or "cublas" in name.lower() or "magma" in name:
        if "implicit_gemm" in name or "conv" in name or "xmma" in name:
            return "cuDNN convolutions (backbone, north_star)"
        return "cuBLAS GEMM"
    if "cudnn" in name.lower() or "xmma" in name or "conv" in name.lower() or "nchw" in name.lower() or "nhwc" in name.lower():
        return "cuDNN convolutions (backbone, north_star)"
    if "nccl" in name.lower():
        return "NCCL"
    return "torch elementwise / copy / reduce"


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    per, fam = {}, {}
    for r in rows[1:]:
        if "wait_host_flag" in r[ki]:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(r[ui], 1e-3)
        n, t = per.get(r[ki], (0, 0.0))
        per[r[ki]] = (n + 1, t + v)
        f = family(r[ki])
        n, t = fam.get(f, (0, 0.0))
        fam[f] = (n + 1, t + v)
    total = sum(t for _, t in per.values())
    launches = sum(n for n, _ in per.values())
    print(f"total: {launches} launches, {total / 1e3:.2f} ms of kernel time in {steps} steps "
          f"({launches // steps} launches / step; per-launch times are cold-cache and serialised: compare shares)\n")
    print("| family | launches / step | ms / step | share |\n|---|---|---|---|")
    for f, (n, t) in sorted(fam.items(), key=lambda x: -x[1][1]):
        print(f"| {f} | {n // steps} | {t / steps / 1e3:.2f} | {100 * t / total:.1f} % |")
    print("\n| kernel | launches / step | us / step | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(per.items(), key=lambda x: -x[1][1])[:30]:
        print(f"| `{k[:90]}` | {n // steps} | {t / steps:.0f} | {100 * t / total:.1f} % |")


if __name__ == "__main__":
    main()
