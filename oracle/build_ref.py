"""TEST INFRASTRUCTURE — builds the UNMODIFIED reference kernel as a checker.

Compiles /root/reference/csrc/{pybind.cpp,gemm_cuda_quick.cu} *where they lie*
(no source is copied into this repo) into ``oracle/_ref/quick_kernels_ref.so``
with the nvcc flags of the reference's own setup.py (setup.py:60-77) plus an
explicit sm_100a gencode (the reference passes no arch at all).

The result is a torch extension exporting the reference's single symbol
``gemm_forward_cuda_quick`` (csrc/pybind.cpp:5-8).  It is only ever loaded by
``tests/`` (GPU parity), ``__graft_entry__.smoke()`` and ``bench.py --impl
reference``; the product path never touches it.  ``oracle/_ref/`` is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = "/root/reference/csrc"
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "quick_kernels_ref"


def ref_so_path() -> str | None:
    if not os.path.isdir(OUT_DIR):
        return None
    for f in os.listdir(OUT_DIR):
        if f.startswith(NAME) and f.endswith(".so"):
            return os.path.join(OUT_DIR, f)
    return None


def build(verbose: bool = False) -> str | None:
    """Build (if the reference tree is present and no .so exists yet)."""
    so = ref_so_path()
    if so is not None:
        return so
    if not os.path.isdir(REF_CSRC):
        return None  # GPU box: only prebuilt files are used
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    load(
        name=NAME,
        sources=[os.path.join(REF_CSRC, "pybind.cpp"), os.path.join(REF_CSRC, "gemm_cuda_quick.cu")],
        extra_cflags=["-O3", "-std=c++17"],
        extra_cuda_cflags=[
            "-O3", "-std=c++17",
            "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
            "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math",
            "--threads=8",
            "-gencode", "arch=compute_100a,code=sm_100a",
        ],
        build_directory=OUT_DIR,
        verbose=verbose,
        is_python_module=False,
    )
    return ref_so_path()


def load_ref():
    """Import the prebuilt reference extension (returns module or None)."""
    so = ref_so_path()
    if so is None:
        return None
    import importlib.util

    import torch  # noqa: F401  (libtorch must be loaded first)

    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("reference extension:", p)
