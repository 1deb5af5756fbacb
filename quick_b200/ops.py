"""Thin torch-facing wrappers over the C-ABI (include/quick_b200.h).

torch is used only for device memory and the current stream; every function below passes raw
device pointers and sizes to libquick_b200.so.  CUDA tensors are required — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.QuickB200Error("quick_b200 ops need CUDA tensors (no CPU fallback exists)")


def shapes_from_quick(x_or_K, qweight, qzeros, scales):
    """Derive (K, N, G) the way the reference does (gemm_cuda_quick.cu:1468,:1477)."""
    K = int(x_or_K) if isinstance(x_or_K, int) else int(x_or_K.shape[-1])
    N = int(qweight.shape[1]) // 4 * 8
    G = K // int(scales.shape[0])
    return K, N, G


def prepack(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor, K: int | None = None):
    """QUICK layout -> B200 layout.  Returns (wq int32 flat, sz int32 flat, K, N, G)."""
    _require_cuda(qweight, qzeros, scales)
    lib = _lib.load()
    K = int(qweight.shape[0]) * 4 if K is None else K
    _, N, G = shapes_from_quick(K, qweight, qzeros, scales)
    _lib.check(lib.qb200_check_shape(1, K, N, G))
    qweight, qzeros, scales = qweight.contiguous(), qzeros.contiguous(), scales.contiguous()
    wq = torch.empty(lib.qb200_wq_bytes(K, N) // 4, dtype=torch.int32, device=qweight.device)
    sz = torch.empty(lib.qb200_sz_bytes(K, N, G) // 4, dtype=torch.int32, device=qweight.device)
    with torch.cuda.device(qweight.device):
        _lib.check(lib.qb200_relayout_from_quick(_ptr(qweight), _ptr(qzeros), _ptr(scales), K, N, G,
                                                 _ptr(wq), _ptr(sz), _stream_ptr()))
    return wq, sz, K, N, G


def pack_quick(q: torch.Tensor, z: torch.Tensor, s: torch.Tensor, G: int):
    """Logical (q uint8 [K,N], z uint8 [K/G,N], s fp16 [K/G,N]) -> QUICK-layout tensors, on the GPU."""
    _require_cuda(q, z, s)
    lib = _lib.load()
    K, N = q.shape
    _lib.check(lib.qb200_check_shape(1, K, N, G))
    q = q.to(torch.uint8).contiguous()
    z = z.to(torch.uint8).contiguous()
    s = s.to(torch.float16).contiguous()
    qweight = torch.empty((K // 4, N // 2), dtype=torch.int32, device=q.device)
    qzeros = torch.empty((K // G, N // 4), dtype=torch.int32, device=q.device)
    scales = torch.empty((K // G, 2 * N), dtype=torch.float16, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(lib.qb200_pack_quick(_ptr(q), _ptr(z), _ptr(s), K, N, G, _ptr(qweight), _ptr(qzeros), _ptr(scales),
                                        _stream_ptr()))
    return qweight, qzeros, scales


def _awq_gemm_shapes(qweight, qzeros, scales):
    """AWQ-GEMM tensors: qweight int32 [K, N/8], qzeros int32 [K/G, N/8], scales fp16 [K/G, N]."""
    K, N = int(qweight.shape[0]), int(qweight.shape[1]) * 8
    if tuple(scales.shape) != (scales.shape[0], N) or tuple(qzeros.shape) != (scales.shape[0], N // 8):
        raise ValueError("not AWQ-GEMM shaped tensors: qweight [K, N/8], qzeros [K/G, N/8], scales [K/G, N]")
    return K, N, K // int(scales.shape[0])


def awq_gemm_to_quick(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor):
    """AWQ-GEMM checkpoint tensors -> QUICK-layout (qweight, qzeros, scales), bit-exact, on the GPU."""
    _require_cuda(qweight, qzeros, scales)
    lib = _lib.load()
    K, N, G = _awq_gemm_shapes(qweight, qzeros, scales)
    _lib.check(lib.qb200_check_shape(1, K, N, G))
    qweight, qzeros, scales = qweight.contiguous(), qzeros.contiguous(), scales.to(torch.float16).contiguous()
    out_qw = torch.empty((K // 4, N // 2), dtype=torch.int32, device=qweight.device)
    out_qz = torch.empty((K // G, N // 4), dtype=torch.int32, device=qweight.device)
    out_sc = torch.empty((K // G, 2 * N), dtype=torch.float16, device=qweight.device)
    with torch.cuda.device(qweight.device):
        _lib.check(lib.qb200_awq_gemm_to_quick(_ptr(qweight), _ptr(qzeros), _ptr(scales), K, N, G, _ptr(out_qw), _ptr(out_qz),
                                               _ptr(out_sc), _stream_ptr()))
    return out_qw, out_qz, out_sc


def prepack_awq_gemm(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor):
    """AWQ-GEMM checkpoint tensors -> B200 layout.  Returns (wq int32 flat, sz int32 flat, K, N, G)."""
    _require_cuda(qweight, qzeros, scales)
    lib = _lib.load()
    K, N, G = _awq_gemm_shapes(qweight, qzeros, scales)
    _lib.check(lib.qb200_check_shape(1, K, N, G))
    qweight, qzeros, scales = qweight.contiguous(), qzeros.contiguous(), scales.to(torch.float16).contiguous()
    wq = torch.empty(lib.qb200_wq_bytes(K, N) // 4, dtype=torch.int32, device=qweight.device)
    sz = torch.empty(lib.qb200_sz_bytes(K, N, G) // 4, dtype=torch.int32, device=qweight.device)
    with torch.cuda.device(qweight.device):
        _lib.check(lib.qb200_relayout_from_awq_gemm(_ptr(qweight), _ptr(qzeros), _ptr(scales), K, N, G, _ptr(wq), _ptr(sz),
                                                    _stream_ptr()))
    return wq, sz, K, N, G


def dequantize(wq: torch.Tensor, sz: torch.Tensor, K: int, N: int, G: int) -> torch.Tensor:
    """B200 layout -> W16 [K, N] fp16 (bit-identical to the reference's in-register weights)."""
    _require_cuda(wq, sz)
    lib = _lib.load()
    W = torch.empty((K, N), dtype=torch.float16, device=wq.device)
    with torch.cuda.device(wq.device):
        _lib.check(lib.qb200_dequantize(_ptr(wq), _ptr(sz), K, N, G, _ptr(W), _stream_ptr()))
    return W


def gemm(x: torch.Tensor, wq: torch.Tensor, sz: torch.Tensor, N: int, G: int, bias: torch.Tensor | None = None,
         tok: int | None = None, split: int | None = None, out: torch.Tensor | None = None,
         independent: bool = False) -> torch.Tensor:
    """y[M,N] = x[M,K] · W (+bias) through the tcgen05 kernel. tok/split force a tile config;
    independent=True passes QB200_GEMM_INDEPENDENT (see include/quick_b200.h)."""
    _require_cuda(x, wq, sz, bias)
    lib = _lib.load()
    assert x.dim() == 2 and x.dtype == torch.float16
    x = x.contiguous()
    M, K = x.shape
    if out is None:
        out = torch.empty((M, N), dtype=torch.float16, device=x.device)
    with torch.cuda.device(x.device):
        if tok is None and split is None and not independent:
            rc = lib.qb200_gemm_w4a16(_ptr(x), _ptr(wq), _ptr(sz), _ptr(bias), _ptr(out), M, K, N, G, 0, _stream_ptr())
        else:
            rc = lib.qb200_gemm_w4a16_ex(_ptr(x), _ptr(wq), _ptr(sz), _ptr(bias), _ptr(out), M, K, N, G,
                                         tok or 0, split or 0, 1 if independent else 0, _stream_ptr())
    _lib.check(rc)
    return out


def gemm_allgather(x: torch.Tensor, wq: torch.Tensor, sz: torch.Tensor, N: int, G: int, peer_ptrs, ld_c: int, col0: int,
                   bias: torch.Tensor | None = None, residual: torch.Tensor | None = None, independent: bool = False,
                   multicast_ptr: int | None = None) -> None:
    """Fused GEMM + all-gather: this rank's [M, N] slab goes to column col0 of every peer buffer (peer_ptrs: device
    pointers of the [rows, ld_c] buffers of all ranks, mapped into this process).  See quick_b200.parallel."""
    _require_cuda(x, wq, sz, bias, residual)
    lib = _lib.load()
    assert x.dim() == 2 and x.dtype == torch.float16
    x = x.contiguous()
    M, K = x.shape
    if residual is not None:
        assert residual.is_contiguous() and residual.dtype == torch.float16 and tuple(residual.shape) == (M, ld_c)
    arr = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
    with torch.cuda.device(x.device):
        _lib.check(lib.qb200_gemm_w4a16_allgather(_ptr(x), _ptr(wq), _ptr(sz), _ptr(bias), _ptr(residual), arr, multicast_ptr or None,
                                                  len(peer_ptrs), ld_c, col0, M, K, N, G, 0, 0, 1 if independent else 0, _stream_ptr()))


def interleave_pairs(wq: torch.Tensor, sz: torch.Tensor, K: int, N: int, G: int, bias: torch.Tensor | None = None):
    """B200-layout weight of a [first | second] concatenation (gate | up) -> the same weight with its output channels
    interleaved (2i = first_i, 2i+1 = second_i): what QB200_GEMM_SILU_MUL expects.  Done once at load; N/2 must be a
    multiple of 64 (every 128-channel tile then holds 64 complete pairs)."""
    if N % 256 != 0:
        raise ValueError("interleave_pairs: N/2 must be a multiple of the 128-column tile")
    NT, KB, NG, I = N // 128, K // 64, K // G, N // 2
    idx = torch.arange(N, device=wq.device)
    src = (idx & 1) * I + (idx >> 1)                                     # new channel n' <- old column src[n']
    w = wq.view(NT, KB, 2, 128, 4).permute(0, 3, 1, 2, 4).reshape(N, KB * 8)[src]
    w = w.reshape(NT, 128, KB, 2, 4).permute(0, 2, 3, 1, 4).contiguous().view(-1)
    z = sz.view(NT, NG, 128).permute(0, 2, 1).reshape(N, NG)[src].reshape(NT, 128, NG).permute(0, 2, 1).contiguous().view(-1)
    return w, z, (None if bias is None else bias[src].contiguous())


def gemm_tp(x: torch.Tensor, wq: torch.Tensor, sz: torch.Tensor, N: int, G: int, bias=None, residual=None, out=None,
            dst=None, col0: int = 0, wait=None, silu_mul: bool = False) -> torch.Tensor | None:
    """qb200_gemm_w4a16_tp.  dst = None: local output [M, N] (returned).  dst = GatheredBuffer: this rank's [M, N] slab
    goes to column col0 of every rank's copy and the fill is published (returns None; read it through dst.rows(M) in a
    kernel that is given dst.wait).  wait = the GatheredBuffer x lives in (its rows were filled by all ranks), or None.
    residual: local [M, N] (dst None) or full-width [M, dst.width] tensor read at the slab's columns."""
    _require_cuda(x, wq, sz, bias, residual)
    lib = _lib.load()
    assert x.dim() == 2 and x.dtype == torch.float16 and x.is_contiguous()
    M, K = x.shape
    w = C.byref(wait.wait) if wait is not None else None
    fl = 2 if silu_mul else 0                  # QB200_GEMM_SILU_MUL: interleaved gate|up weight, output [M, N/2]
    with torch.cuda.device(x.device):
        if dst is None:
            if out is None:
                out = torch.empty((M, N // 2 if silu_mul else N), dtype=torch.float16, device=x.device)
            _lib.check(lib.qb200_gemm_w4a16_tp(_ptr(x), _ptr(wq), _ptr(sz), _ptr(bias), _ptr(residual), _ptr(out), None, None, 0, N, 0,
                                               M, K, N, G, 0, 0, fl, w, None, _stream_ptr()))
            return out
        if M > dst.max_rows:
            raise ValueError(f"M={M} exceeds the gathered buffer ({dst.max_rows} rows)")
        if residual is not None:
            assert residual.is_contiguous() and residual.dtype == torch.float16 and tuple(residual.shape) == (M, dst.width)
        _lib.check(lib.qb200_gemm_w4a16_tp(_ptr(x), _ptr(wq), _ptr(sz), _ptr(bias), _ptr(residual), None, dst.buf_ptrs, dst.multicast_ptr,
                                           dst.world, dst.width, col0, M, K, N, G, 0, 0, fl, w, C.byref(dst.signal), _stream_ptr()))
    return None


def rmsnorm_tp(x: torch.Tensor, weight: torch.Tensor, eps: float, wait=None) -> torch.Tensor:
    """qb200_rmsnorm_tp: RMSNorm of rows that live in a gathered buffer (wait = that GatheredBuffer, or None)."""
    _require_cuda(x, weight)
    lib = _lib.load()
    assert x.dtype == torch.float16 and x.is_contiguous()
    H = x.shape[-1]
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(lib.qb200_rmsnorm_tp(_ptr(x), _ptr(weight), _ptr(y), x.numel() // H, H, float(eps),
                                        C.byref(wait.wait) if wait is not None else None, _stream_ptr()))
    return y


def silu_mul_interleaved(gate_up: torch.Tensor) -> torch.Tensor:
    """qb200_silu_mul_interleaved: rows (g_0, u_0, g_1, u_1, ...) -> silu(g) * u, [..., I]."""
    _require_cuda(gate_up)
    lib = _lib.load()
    assert gate_up.dtype == torch.float16 and gate_up.is_contiguous()
    I = gate_up.shape[-1] // 2
    act = torch.empty(gate_up.shape[:-1] + (I,), dtype=torch.float16, device=gate_up.device)
    with torch.cuda.device(gate_up.device):
        _lib.check(lib.qb200_silu_mul_interleaved(_ptr(gate_up), _ptr(act), gate_up.numel() // (2 * I), I, _stream_ptr()))
    return act


def silu_mul_tp(gate_up: torch.Tensor, dst, col0: int) -> None:
    """qb200_silu_mul_tp: silu(g) * u of this rank's [rows, 2 I] slab -> column col0 of every rank's copy of dst, published."""
    _require_cuda(gate_up)
    lib = _lib.load()
    assert gate_up.dtype == torch.float16 and gate_up.is_contiguous()
    I = gate_up.shape[-1] // 2
    rows = gate_up.numel() // (2 * I)
    with torch.cuda.device(gate_up.device):
        _lib.check(lib.qb200_silu_mul_tp(_ptr(gate_up), rows, I, dst.buf_ptrs, dst.multicast_ptr, dst.world, dst.width, col0,
                                         C.byref(dst.signal), _stream_ptr()))


def attn_decode_tp(qkv: torch.Tensor, cos_table: torch.Tensor, sin_table: torch.Tensor, pos: torch.Tensor, cache_k: torch.Tensor,
                   cache_v: torch.Tensor, nh: int, nkv: int, dst, col0: int) -> None:
    """qb200_attn_decode_tp: one-token attention of this rank's heads (rotary + KV-cache update + attention over the cache)
    whose output [B, nh*hd] lands at column col0 of every rank's copy of dst (a GatheredBuffer)."""
    _require_cuda(qkv, cache_k, cache_v)
    lib = _lib.load()
    B, S, hd = cache_k.shape[0], cache_k.shape[2], cache_k.shape[3]
    assert qkv.is_contiguous() and qkv.dtype == torch.float16 and qkv.numel() == B * (nh + 2 * nkv) * hd
    assert cache_k.is_contiguous() and cache_v.is_contiguous() and pos.dtype == torch.long and pos.numel() == 1
    with torch.cuda.device(qkv.device):
        _lib.check(lib.qb200_attn_decode_tp(_ptr(qkv), _ptr(cos_table), _ptr(sin_table), _ptr(pos), _ptr(cache_k), _ptr(cache_v), B, nh, nkv,
                                            hd, S, float(np.float32(1.0) / np.sqrt(np.float32(hd))), dst.buf_ptrs, dst.world, dst.width, col0, C.byref(dst.signal),
                                            _stream_ptr()))


def scatter_cols(src: torch.Tensor, dst, col0: int) -> None:
    """qb200_scatter_cols: this rank's [rows, n_local] slab -> column col0 of every rank's copy of dst, published."""
    _require_cuda(src)
    lib = _lib.load()
    assert src.dtype == torch.float16 and src.is_contiguous()
    n_local = src.shape[-1]
    with torch.cuda.device(src.device):
        _lib.check(lib.qb200_scatter_cols(_ptr(src), src.numel() // n_local, n_local, dst.buf_ptrs, dst.multicast_ptr, dst.world,
                                          dst.width, col0, C.byref(dst.signal), _stream_ptr()))


def peer_barrier(epoch: torch.Tensor, flag_ptrs, rank: int) -> None:
    """qb200_peer_barrier on the current stream (epoch: 1-element int32 device tensor owned by the caller)."""
    lib = _lib.load()
    arr = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
    with torch.cuda.device(epoch.device):
        _lib.check(lib.qb200_peer_barrier(_ptr(epoch), arr, rank, len(flag_ptrs), _stream_ptr()))


def gemm_simt(x: torch.Tensor, wq: torch.Tensor, sz: torch.Tensor, N: int, G: int) -> torch.Tensor:
    """CUDA-core cross-check of the same contraction (tests only)."""
    _require_cuda(x, wq, sz)
    lib = _lib.load()
    x = x.contiguous()
    M, K = x.shape
    out = torch.empty((M, N), dtype=torch.float16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.qb200_gemm_w4a16_simt(_ptr(x), _ptr(wq), _ptr(sz), _ptr(out), M, K, N, G, _stream_ptr()))
    return out


def plan(M: int, K: int, N: int, G: int, split_hint: int = 0, independent: bool = False):
    lib = _lib.load()
    tok, split, ctas = C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.qb200_gemm_plan_ex(M, K, N, G, split_hint, 1 if independent else 0, C.byref(tok), C.byref(split), C.byref(ctas)))
    return tok.value, split.value, ctas.value


def gemm_forward_quick_stateless(x, qweight, scales, qzeros, split_k_iters: int = 8) -> torch.Tensor:
    """qb200_gemm_forward_quick: relayout into a scratch workspace + GEMM, no caching."""
    _require_cuda(x, qweight, scales, qzeros)
    lib = _lib.load()
    x = x.contiguous()
    M, K = x.shape
    _, N, G = shapes_from_quick(K, qweight, qzeros, scales)
    _lib.check(lib.qb200_check_shape(M, K, N, G))
    nbytes = lib.qb200_wq_bytes(K, N) + lib.qb200_sz_bytes(K, N, G)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    out = torch.empty((M, N), dtype=torch.float16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.qb200_gemm_forward_quick(_ptr(x), _ptr(qweight.contiguous()), _ptr(scales.contiguous()),
                                                _ptr(qzeros.contiguous()), _ptr(out), M, K, N, G, split_k_iters,
                                                _ptr(ws), nbytes, _stream_ptr()))
    return out


class HostLinear:
    """qb200_linear handle: HOST buffers in, HOST buffers out (H2D + kernel + D2H inside the call)."""

    def __init__(self, qweight, qzeros, scales, bias=None, max_m: int = 512, device: int = 0):
        lib = _lib.load()
        qweight, qzeros, scales = (t.cpu().contiguous() for t in (qweight, qzeros, scales))
        self.K = int(qweight.shape[0]) * 4
        _, self.N, self.G = shapes_from_quick(self.K, qweight, qzeros, scales)
        self.max_m = max_m
        bias = None if bias is None else bias.cpu().contiguous()
        h = C.c_void_p()
        _lib.check(lib.qb200_linear_create(C.byref(h), _ptr(qweight), _ptr(qzeros), _ptr(scales), _ptr(bias),
                                           self.K, self.N, self.G, max_m, device))
        self._h = h
        self._lib = lib

    def forward_host(self, x_host: torch.Tensor, y_host: torch.Tensor | None = None) -> torch.Tensor:
        assert not x_host.is_cuda and x_host.dtype == torch.float16 and x_host.is_contiguous()
        M = x_host.shape[0]
        if y_host is None:
            y_host = torch.empty((M, self.N), dtype=torch.float16, pin_memory=True)
        _lib.check(self._lib.qb200_linear_forward_host(self._h, _ptr(x_host), _ptr(y_host), M))
        return y_host

    def forward_host_async(self, x_host: torch.Tensor, y_host: torch.Tensor) -> torch.Tensor:
        """Enqueue H2D -> GEMM -> D2H on the handle's stream; both buffers must be pinned and stay alive
        until synchronize()."""
        assert not x_host.is_cuda and x_host.dtype == torch.float16 and x_host.is_contiguous()
        assert x_host.is_pinned() and y_host.is_pinned(), "asynchronous host calls need pinned buffers"
        _lib.check(self._lib.qb200_linear_forward_host_async(self._h, _ptr(x_host), _ptr(y_host), x_host.shape[0]))
        return y_host

    def synchronize(self):
        _lib.check(self._lib.qb200_linear_synchronize(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.qb200_linear_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
