"""Host-side QUICK layout algebra in vectorised torch (works on CPU or CUDA tensors).

This is the product's own restatement of the packed operand format the reference defines in
quick/awq/modules/linear/quick.py:52-54 (shapes) and :88-150 (interleave), used where no GPU is
involved (building modules on CPU, sharding, concatenation).  On the GPU the same transform is
``quick_b200.ops.pack_quick`` (C-ABI ``qb200_pack_quick``).  The closed form:

  qweight word  f = kt*4N + (2*ty + l//16)*N + bx*128 + (l%16)*8 + ks*4 + ch
    holds, in nibble p, q[k0 + DK[p]][c0 + DC[p]],  k0 = 32kt + 16ks + 2(l%4),  c0 = 128bx + 64ty + 16ch + l//4
  scale/zero slot x of a row -> column 128bx + 64ty + 16(m//2) + 8(m%2) + 4lh + j4
"""
from __future__ import annotations

import torch

_DK = (0, 8, 0, 8, 1, 9, 1, 9)
_DC = (0, 0, 8, 8, 0, 0, 8, 8)


def _word_coords(K: int, N: int, device):
    f = torch.arange(K * N // 8, device=device, dtype=torch.int64)
    kt = f // (4 * N)
    r = f % (4 * N)
    r4 = r // N
    c = r % N
    bx = c // 128
    w = c % 128
    lane = 16 * (r4 % 2) + w // 8
    ty = r4 // 2
    ks = (w % 8) // 4
    ch = w % 4
    k0 = 32 * kt + 16 * ks + 2 * (lane % 4)
    c0 = 128 * bx + 64 * ty + 16 * ch + lane // 4
    return k0, c0


def slot_to_column(N: int, device=None) -> torch.Tensor:
    x = torch.arange(N, device=device, dtype=torch.int64)
    nb = N // 128
    ty = x // (N // 2)
    lh = (x // (N // 4)) % 2
    bx = (x // 32) % nb
    j4 = (x % 32) // 8
    m = x % 8
    return 128 * bx + 64 * ty + 16 * (m // 2) + 8 * (m % 2) + 4 * lh + j4


def _to_i32(u: torch.Tensor) -> torch.Tensor:
    """int64 holding a uint32 bit pattern -> int32 with the same bits."""
    return torch.where(u >= 2 ** 31, u - 2 ** 32, u).to(torch.int32)


def pack_quick(q: torch.Tensor, z: torch.Tensor, s: torch.Tensor):
    """q[K,N], z[K/G,N] integer 0..15, s[K/G,N] -> (qweight, qzeros, scales) in the reference's shapes."""
    K, N = q.shape
    if N % 128 != 0:
        raise ValueError("OC is not multiple of cta_N = 128")
    if K % 32 != 0:
        raise ValueError("IC is not a multiple of 32")
    dev = q.device
    k0, c0 = _word_coords(K, N, dev)
    q64 = q.to(torch.int64)
    word = torch.zeros(K * N // 8, dtype=torch.int64, device=dev)
    for p in range(8):
        word |= (q64[k0 + _DK[p], c0 + _DC[p]] & 0xF) << (4 * p)
    qweight = _to_i32(word).reshape(K // 4, N // 2).contiguous()
    col = slot_to_column(N, dev)
    scales = s.to(torch.float16)[:, col].repeat_interleave(2, dim=1).contiguous()
    z4 = (z.to(torch.int64)[:, col] & 0xF).reshape(z.shape[0], N // 4, 4)
    zw = z4[..., 0] | (z4[..., 1] << 4) | (z4[..., 2] << 8) | (z4[..., 3] << 12)
    zw = zw | (zw << 16)
    qzeros = _to_i32(zw).contiguous()
    return qweight, qzeros, scales


def unpack_quick(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor):
    """Inverse of pack_quick -> (q int32 [K,N], z int32 [K/G,N], s fp16 [K/G,N])."""
    K, N = qweight.shape[0] * 4, qweight.shape[1] * 2
    NG = qzeros.shape[0]
    dev = qweight.device
    word = qweight.reshape(-1).to(torch.int64) & 0xFFFFFFFF
    k0, c0 = _word_coords(K, N, dev)
    q = torch.zeros((K, N), dtype=torch.int32, device=dev)
    for p in range(8):
        q[k0 + _DK[p], c0 + _DC[p]] = ((word >> (4 * p)) & 0xF).to(torch.int32)
    col = slot_to_column(N, dev)
    s = torch.zeros((NG, N), dtype=torch.float16, device=dev)
    s[:, col] = scales[:, 0::2]
    zw = qzeros.to(torch.int64) & 0xFFFFFFFF
    z = torch.zeros((NG, N), dtype=torch.int32, device=dev)
    for i in range(4):
        z[:, col[i::4]] = ((zw >> (4 * i)) & 0xF).to(torch.int32)
    return q, z, s


AWQ_ORDER = (0, 2, 4, 6, 1, 3, 5, 7)   # nibble i of an AWQ-GEMM word holds column 8c + AWQ_ORDER[i] (gemm.py:117-123)


def unpack_awq_gemm(qweight: torch.Tensor, qzeros: torch.Tensor):
    """AWQ "GEMM" checkpoint tensors (qweight int32 [K, N/8], qzeros int32 [K/G, N/8]) -> logical
    (q int32 [K, N], z int32 [K/G, N]).  Same result as the reference's unpack_awq + reverse_awq_order + mask
    (quick/awq/utils/packing_utils.py:8-39, :82-90)."""
    def unpack(t):
        w = t.to(torch.int64) & 0xFFFFFFFF
        out = torch.empty((t.shape[0], t.shape[1], 8), dtype=torch.int32, device=t.device)
        for i in range(8):
            out[:, :, AWQ_ORDER[i]] = ((w >> (4 * i)) & 0xF).to(torch.int32)
        return out.reshape(t.shape[0], -1)
    return unpack(qweight), unpack(qzeros)


def pack_awq_gemm(q: torch.Tensor, z: torch.Tensor):
    """Logical (q [K, N], z [K/G, N]) -> AWQ-GEMM (qweight int32 [K, N/8], qzeros int32 [K/G, N/8]); the
    reference packer's loops (quick/awq/modules/linear/gemm.py:108-143) in closed form."""
    def pack(t):
        t8 = (t.to(torch.int64) & 0xF).reshape(t.shape[0], -1, 8)
        w = torch.zeros(t8.shape[:2], dtype=torch.int64, device=t.device)
        for i in range(8):
            w |= t8[:, :, AWQ_ORDER[i]] << (4 * i)
        return _to_i32(w).contiguous()
    return pack(q), pack(z)


def awq_gemm_to_quick(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor):
    """AWQ-GEMM checkpoint tensors -> QUICK-layout (qweight, qzeros, scales); works on CPU or CUDA tensors
    (the GPU kernels are quick_b200.ops.awq_gemm_to_quick / prepack_awq_gemm)."""
    q, z = unpack_awq_gemm(qweight, qzeros)
    return pack_quick(q, z, scales.to(torch.float16))


def quick_cat(tensors, options: str) -> torch.Tensor:
    """N-concatenation of packed tensors (reference QUICK_cat, fused_utils.py:119-159), also for
    unequal widths (GQA k/v projections), which the reference rejects (fused_utils.py:139-142)."""
    if len(tensors) < 2:
        raise ValueError("At least two input layers are required")
    H = tensors[0].shape[0]
    for t in tensors[1:]:
        if t.shape[0] != H:
            raise ValueError("All input layers must have the same number of rows")
    if options == "qweight":
        rows = H // 2
    elif options in ("qzeros", "scales"):
        rows = H * 4
    else:
        raise ValueError("Unknown options provided or invalid reshape dimensions")
    return torch.cat([t.reshape(rows, -1) for t in tensors], dim=1).reshape(H, -1)


def slice_columns(qweight, qzeros, scales, n0: int, n1: int):
    """Output columns [n0, n1) of a packed weight as a packed weight of its own (n0, n1 multiples of the 128-column
    tile): the inverse slice of quick_cat."""
    K, N = qweight.shape[0] * 4, qweight.shape[1] * 2
    NG = qzeros.shape[0]
    if not (0 <= n0 < n1 <= N) or n0 % 128 != 0 or n1 % 128 != 0:
        raise ValueError(f"columns [{n0}, {n1}) of N={N} are not a run of 128-column tiles")
    qw = qweight.reshape(K // 8, N)[:, n0:n1].reshape(K // 4, -1).contiguous()
    sc = scales.reshape(4 * NG, N // 2)[:, n0 // 2:n1 // 2].reshape(NG, -1).contiguous()
    qz = qzeros.reshape(4 * NG, N // 16)[:, n0 // 16:n1 // 16].reshape(NG, -1).contiguous()
    return qw, qz, sc


def shard_columns(qweight, qzeros, scales, rank: int, world: int):
    """Column-parallel (N) shard `rank` of `world` equal shards of a packed weight."""
    N = qweight.shape[1] * 2
    if N % world != 0 or (N // world) % 128 != 0:
        raise ValueError(f"N={N} cannot be split into {world} shards of 128-column tiles")
    return slice_columns(qweight, qzeros, scales, rank * N // world, (rank + 1) * N // world)


def pad_columns(qweight, qzeros, scales, n_new: int):
    """Append zero-weight output columns up to N = n_new (a multiple of 128): q = z = 0 and s = 0, so the new
    channels compute exactly 0.  Used to make a width divisible by 128 x the tensor-parallel degree."""
    K, N = qweight.shape[0] * 4, qweight.shape[1] * 2
    NG = qzeros.shape[0]
    if n_new < N or n_new % 128 != 0:
        raise ValueError(f"cannot pad N={N} to {n_new}")
    if n_new == N:
        return qweight, qzeros, scales
    e = n_new - N
    zw = torch.zeros((K // 4, e // 2), dtype=qweight.dtype, device=qweight.device)
    zz = torch.zeros((NG, e // 4), dtype=qzeros.dtype, device=qzeros.device)
    zs = torch.zeros((NG, 2 * e), dtype=scales.dtype, device=scales.device)
    return (quick_cat([qweight, zw], "qweight").contiguous(), quick_cat([qzeros, zz], "qzeros").contiguous(),
            quick_cat([scales, zs], "scales").contiguous())


def pad_rows(qweight, qzeros, scales, k_new: int, G: int):
    """Append zero-weight input channels up to K = k_new (a multiple of the group size): the QUICK layout is k-tile
    major, so new k-tiles are new rows."""
    K = qweight.shape[0] * 4
    if k_new < K or k_new % G != 0 or k_new % 64 != 0:
        raise ValueError(f"cannot pad K={K} to {k_new}")
    if k_new == K:
        return qweight, qzeros, scales
    e = k_new - K
    return (torch.cat([qweight, torch.zeros((e // 4, qweight.shape[1]), dtype=qweight.dtype, device=qweight.device)], 0).contiguous(),
            torch.cat([qzeros, torch.zeros((e // G, qzeros.shape[1]), dtype=qzeros.dtype, device=qzeros.device)], 0).contiguous(),
            torch.cat([scales, torch.zeros((e // G, scales.shape[1]), dtype=scales.dtype, device=scales.device)], 0).contiguous())


def quantize_rtn(W: torch.Tensor, G: int):
    """Asymmetric uint4 round-to-nearest of W[N,K] per group of G input channels: the q/z/s semantics
    of the reference's pseudo_quantize_tensor (quick/awq/quantize/quantizer.py:46-72).
    Returns q[K,N] int32, z[K/G,N] int32, s[K/G,N] fp16."""
    N, K = W.shape
    w = W.float().reshape(N, K // G, G)
    mx, mn = w.amax(dim=2, keepdim=True), w.amin(dim=2, keepdim=True)
    s = ((mx - mn).clamp(min=1e-5) / 15).half().float()
    z = (-torch.round(mn / s)).clamp(0, 15)
    q = torch.clamp(torch.round(w / s) + z, 0, 15)
    return (q.reshape(N, K).t().contiguous().to(torch.int32), z.squeeze(2).t().contiguous().to(torch.int32),
            s.squeeze(2).t().contiguous().half())
