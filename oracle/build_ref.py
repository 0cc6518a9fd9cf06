"""TEST INFRASTRUCTURE - builds the reference's OWN CUDA op for sm_100a into oracle/_ref/.

Sources are compiled where they lie under /root/reference/models/ops/src (nothing is copied into
the repo); the only addition is the forced-include shim oracle/ref_compat.h that adapts one
dispatch macro to torch 2.11.  The result, oracle/_ref/msda_reference_cuda*.so, is a torch
extension exporting the reference's `ms_deform_attn_forward/backward` (vision.cpp:13-16).  It is
git-ignored but travels to the GPU box with gpurun, where tests use it as a second oracle and
bench.py times it as the "reference GPU kernel" bar-to-beat.

Only runs where /root/reference exists (this container).  `python oracle/build_ref.py`
"""
import glob
import os
import shutil
import sys

REF_SRC = "/root/reference/models/ops/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "msda_reference_cuda"


def built_path():
    hits = glob.glob(os.path.join(OUT, NAME + "*.so"))
    return hits[0] if hits else None


def build(force=False):
    if not os.path.isdir(REF_SRC):
        return built_path()
    if built_path() and not force:
        return built_path()
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils import cpp_extension
    shim = os.path.join(HERE, "ref_compat.h")
    build_dir = os.path.join(OUT, "_build")
    os.makedirs(build_dir, exist_ok=True)
    cpp_extension.load(
        name=NAME,
        sources=[os.path.join(REF_SRC, "vision.cpp"),
                 os.path.join(REF_SRC, "cpu", "ms_deform_attn_cpu.cpp"),
                 os.path.join(REF_SRC, "cuda", "ms_deform_attn_cuda.cu")],
        extra_include_paths=[REF_SRC],
        extra_cflags=["-DWITH_CUDA", "-include", shim, "-w"],
        extra_cuda_cflags=["-DWITH_CUDA", "-include", shim, "-w", "-lineinfo",
                           "-gencode", "arch=compute_100a,code=sm_100a"],
        build_directory=build_dir, is_python_module=False, verbose=False)
    so = os.path.join(build_dir, NAME + ".so")
    shutil.copy2(so, os.path.join(OUT, NAME + ".so"))
    shutil.rmtree(build_dir, ignore_errors=True)
    return built_path()


def load():
    """Import the built reference op (needs a CUDA-capable torch process to be useful)."""
    path = built_path()
    if path is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
