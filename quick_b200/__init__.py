"""quick_b200 — B200-native W4A16 grouped GEMM behind the SqueezeBits/QUICK plugin surface.

Layout of the package (only what the hot path needs):
  csrc/            CUDA kernels (tcgen05/TMEM/TMA) + the C-ABI + the `quick_kernels` torch binding
  _lib.py, ops.py  ctypes binding of the C-ABI and torch-tensor wrappers over it
  layout.py        QUICK packed-format algebra (pack / unpack / concat / column shard) in torch
  awq/             host-side mirror of the reference's WQLinear_QUICK and QUICK_cat
  build.py         in-tree nvcc / torch-extension build
"""
__version__ = "0.1.0"
