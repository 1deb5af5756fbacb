"""Packaging for quick_b200 — the counterpart of the reference's setup.py (which builds the `quick_kernels`
CUDAExtension from csrc/, setup.py:82-89).  Here the native pieces are built in-tree by `quick_b200.build`
(nvcc -gencode arch=compute_100a,code=sm_100a for libquick_b200.so, torch cpp_extension for quick_kernels.so) and shipped
as package data; `pip install -e .` or `python setup.py build_py` triggers that build.  sm_100a only — no fallback."""
import os
import sys

from setuptools import find_packages, setup
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))


class BuildNative(build_py):
    def run(self):
        sys.path.insert(0, ROOT)
        from quick_b200 import build as native
        native.build_all()
        super().run()


setup(
    name="quick_b200",
    version="0.1.0",
    description="B200-native W4A16 grouped GEMM behind the AutoAWQ / QUICK plugin surface (quick_kernels, WQLinear_QUICK, AutoAWQForCausalLM)",
    packages=find_packages(include=["quick_b200", "quick_b200.*"]),
    package_data={"quick_b200": ["libquick_b200.so", "csrc/*"]},
    data_files=[("", ["quick_kernels.so"])] if os.path.exists(os.path.join(ROOT, "quick_kernels.so")) else [],
    python_requires=">=3.10",
    install_requires=["torch>=2.6", "numpy"],
    extras_require={"awq": ["transformers>=5.0", "safetensors"]},
    cmdclass={"build_py": BuildNative},
)
