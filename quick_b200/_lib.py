"""ctypes binding of the C-ABI declared in include/quick_b200.h.

Fails loudly when the shared library is missing: there is no CPU or eager fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QB200_LIB") or os.path.join(_PKG, "libquick_b200.so")   # QB200_LIB: experiment builds only

# every symbol include/quick_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "qb200_version", "qb200_last_error", "qb200_wq_bytes", "qb200_sz_bytes", "qb200_check_shape",
    "qb200_relayout_from_quick", "qb200_pack_quick", "qb200_awq_gemm_to_quick", "qb200_relayout_from_awq_gemm",
    "qb200_dequantize", "qb200_gemm_w4a16",
    "qb200_gemm_w4a16_cfg", "qb200_gemm_w4a16_ex", "qb200_gemm_w4a16_fused", "qb200_gemm_w4a16_norm", "qb200_attn_decode", "qb200_attn_decode_smem_bytes", "qb200_rmsnorm", "qb200_rope_kv_update",
    "qb200_silu_mul", "qb200_silu_mul_interleaved", "qb200_gemm_w4a16_allgather", "qb200_peer_barrier",
    "qb200_gemm_w4a16_tp", "qb200_rmsnorm_tp", "qb200_silu_mul_tp", "qb200_scatter_cols", "qb200_attn_decode_tp", "qb200_gemm_plan", "qb200_gemm_plan_ex", "qb200_gemm_forward_quick", "qb200_gemm_w4a16_simt",
    "qb200_linear_create", "qb200_linear_forward_host", "qb200_linear_forward_host_async", "qb200_linear_synchronize",
    "qb200_linear_forward", "qb200_linear_destroy",
    "qb200_launch_count", "qb200_debug_set_trace", "qb200_debug_set_variant",
]

QB200_OK, QB200_EINVAL, QB200_ECUDA, QB200_ENOSPC = 0, -1, -2, -3


class PeerWait(C.Structure):          # qb200_peer_wait
    _fields_ = [("epoch", C.c_void_p), ("flag_arrays", C.POINTER(C.c_void_p)), ("rank", C.c_int), ("n_peers", C.c_int)]


class PeerSignal(C.Structure):        # qb200_peer_signal
    _fields_ = [("epoch", C.c_void_p)]


_lib = None


class QuickB200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QuickB200Error(
            f"{LIB_PATH} is missing: build it with `python -m quick_b200.build` "
            "(quick_b200 has no CPU/eager fallback for the W4A16 GEMM)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    lib.qb200_version.restype = C.c_char_p
    lib.qb200_last_error.restype = C.c_char_p
    lib.qb200_wq_bytes.restype = sz
    lib.qb200_wq_bytes.argtypes = [i32, i32]
    lib.qb200_sz_bytes.restype = sz
    lib.qb200_sz_bytes.argtypes = [i32, i32, i32]
    lib.qb200_check_shape.argtypes = [i32, i32, i32, i32]
    lib.qb200_relayout_from_quick.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp]
    lib.qb200_pack_quick.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.qb200_awq_gemm_to_quick.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.qb200_relayout_from_awq_gemm.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp]
    lib.qb200_dequantize.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.qb200_gemm_w4a16.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.qb200_gemm_w4a16_cfg.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.qb200_gemm_w4a16_ex.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, C.c_uint, vp]
    lib.qb200_gemm_plan.argtypes = [i32, i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.qb200_gemm_w4a16_fused.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, C.c_uint, vp]
    lib.qb200_gemm_w4a16_allgather.argtypes = [vp, vp, vp, vp, vp, C.POINTER(vp), vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, C.c_uint, vp]
    lib.qb200_peer_barrier.argtypes = [vp, C.POINTER(vp), i32, i32, vp]
    lib.qb200_rmsnorm.argtypes = [vp, vp, vp, i32, i32, C.c_float, vp]
    lib.qb200_gemm_w4a16_tp.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(vp), vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, C.c_uint,
                                        C.POINTER(PeerWait), C.POINTER(PeerSignal), vp]
    lib.qb200_rmsnorm_tp.argtypes = [vp, vp, vp, i32, i32, C.c_float, C.POINTER(PeerWait), vp]
    lib.qb200_silu_mul_tp.argtypes = [vp, C.c_longlong, i32, C.POINTER(vp), vp, i32, i32, i32, C.POINTER(PeerSignal), vp]
    lib.qb200_attn_decode_tp.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, C.c_float, C.POINTER(vp), i32, i32, i32,
                                         C.POINTER(PeerSignal), vp]
    lib.qb200_scatter_cols.argtypes = [vp, C.c_longlong, i32, C.POINTER(vp), vp, i32, i32, i32, C.POINTER(PeerSignal), vp]
    lib.qb200_rope_kv_update.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.qb200_silu_mul.argtypes = [vp, vp, C.c_longlong, i32, vp]
    lib.qb200_silu_mul_interleaved.argtypes = [vp, vp, C.c_longlong, i32, vp]
    lib.qb200_attn_decode_smem_bytes.argtypes = [i32, i32, i32, i32]
    lib.qb200_attn_decode.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, C.c_float, vp]
    lib.qb200_gemm_plan_ex.argtypes = [i32, i32, i32, i32, i32, C.c_uint, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.qb200_gemm_forward_quick.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, sz, vp]
    lib.qb200_gemm_w4a16_simt.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.qb200_linear_create.argtypes = [C.POINTER(vp), vp, vp, vp, vp, i32, i32, i32, i32, i32]
    lib.qb200_linear_forward_host.argtypes = [vp, vp, vp, i32]
    lib.qb200_linear_forward_host_async.argtypes = [vp, vp, vp, i32]
    lib.qb200_linear_synchronize.argtypes = [vp]
    lib.qb200_linear_forward.argtypes = [vp, vp, vp, i32, vp]
    lib.qb200_linear_destroy.argtypes = [vp]
    lib.qb200_linear_destroy.restype = None
    lib.qb200_launch_count.restype = C.c_ulonglong
    lib.qb200_debug_set_trace.argtypes = [vp]
    lib.qb200_debug_set_trace.restype = None
    lib.qb200_debug_set_variant.argtypes = [i32]
    lib.qb200_debug_set_variant.restype = None
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("qb200_version",):
            pass
    _lib = lib
    return lib


def check(rc: int):
    """Map C-ABI return codes to the exceptions the reference raises (ValueError for shape errors)."""
    if rc == QB200_OK:
        return
    msg = load().qb200_last_error().decode()
    if rc == QB200_EINVAL:
        raise ValueError(msg)
    raise QuickB200Error(f"quick_b200 error {rc}: {msg}")
