"""In-tree build of the native pieces (no pip install; the .so files travel with the repo snapshot).

  libquick_b200.so   CUDA kernels + C-ABI (include/quick_b200.h)           -> quick_b200/
  quick_kernels.so   torch extension, drop-in for the reference's module    -> repo root (top-level import)

``python -m quick_b200.build`` builds both; nvcc cross-compiles sm_100a without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libquick_b200.so")
EXT = os.path.join(ROOT, "quick_kernels.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
    "-Xlinker", "-rpath=/usr/local/cuda/lib64",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_lib(force: bool = False, verbose: bool = False, defines: tuple[str, ...] = (), out: str | None = None) -> str:
    """defines/out: development builds only (e.g. ("QB200_TRACE", "QB200_VARIANTS") -> libquick_b200_dev.so,
    selected with the QB200_LIB environment variable); the product library is built without defines."""
    srcs = [os.path.join(CSRC, "quick_b200.cu"), os.path.join(CSRC, "w4a16_umma.cuh"),
            os.path.join(ROOT, "include", "quick_b200.h")]
    target = out or LIB
    if not force and _newer(target, srcs):
        return target
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    tmp = target + ".tmp"      # built aside and renamed: a process that has the old file mapped keeps its inode
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", tmp, srcs[0]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    os.replace(tmp, target)
    return target


DEV_LIB = os.path.join(PKG, "libquick_b200_dev.so")       # + alternative tile variants (timing A/B)
TRACE_LIB = os.path.join(PKG, "libquick_b200_trace.so")   # + in-kernel clock64 trace (tools/trace.py)


def build_dev_libs(force: bool = False):
    """Development builds used by tools/ through QB200_LIB (never loaded by the product path)."""
    return (build_lib(force=force, defines=("QB200_VARIANTS",), out=DEV_LIB),
            build_lib(force=force, defines=("QB200_TRACE", "QB200_VARIANTS"), out=TRACE_LIB))


def build_ext(force: bool = False, verbose: bool = False) -> str:
    src = os.path.join(CSRC, "quick_kernels_ext.cpp")
    if not force and _newer(EXT, [src, os.path.join(ROOT, "include", "quick_b200.h"), LIB]):
        return EXT
    build_lib(force=False, verbose=verbose)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    bdir = os.path.join(ROOT, "build", "quick_kernels")
    os.makedirs(bdir, exist_ok=True)
    load(
        name="quick_kernels",
        sources=[src],
        extra_cflags=["-O2", "-std=c++17"],
        extra_ldflags=[f"-L{PKG}", "-lquick_b200", "-Wl,-rpath,'$$ORIGIN/quick_b200'", "-Wl,-rpath,'$$ORIGIN'",
                       f"-Wl,-rpath,{PKG}"],
        with_cuda=True,
        build_directory=bdir,
        verbose=verbose,
        is_python_module=False,
    )
    shutil.copy2(os.path.join(bdir, "quick_kernels.so"), EXT + ".tmp")
    os.replace(EXT + ".tmp", EXT)
    return EXT


def build_all(force: bool = False, verbose: bool = False):
    return build_lib(force, verbose), build_ext(force, verbose)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--dev" in sys.argv:
        print(build_dev_libs(force="--force" in sys.argv))
