"""Compile the CUDA sources under rlipv2_b200/csrc/ for sm_100a into rlipv2_b200/lib/*.so.

Plain ``nvcc -shared``: the libraries expose a C ABI (include/*.h) and do not link against torch,
so the build takes seconds and cross-compiles on a box without a GPU.

    python -m rlipv2_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
INCLUDE = os.path.join(ROOT, "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
    "-I", INCLUDE, "-I", CSRC,
]

# library name -> (sources, extra flags)
TARGETS = {
    "librlipv2_msda.so": (["msda.cu"], []),
    "librlipv2_dense.so": (["dense_tf32.cu"], []),
    "librlipv2_fused.so": (["fused_ops.cu"], []),
    "librlipv2_lsap.so": (["lsap.cu"], []),
    "librlipv2_attn.so": (["attn_tf32.cu"], []),
}


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def lib_path(name):
    return os.path.join(LIB, name)


def build_all(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    built = []
    for name, (srcs, extra) in TARGETS.items():
        out = lib_path(name)
        paths = [os.path.join(CSRC, s) for s in srcs]
        deps = paths + [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)] + \
               [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
        if (not force and os.path.exists(out)
                and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps)):
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + paths
        subprocess.check_call(cmd)
        built.append(name)
    return built


if __name__ == "__main__":
    b = build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", b if b else "(up to date)")
