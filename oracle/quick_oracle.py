"""TEST INFRASTRUCTURE — CPU oracle for the QUICK W4A16 grouped GEMM.

A from-scratch numpy restatement of what the reference computes on its hot path
(SURVEY.md §8a).  Nothing here is shipped or timed as the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned two ways:
  * ``tests/golden/*.npz`` — outputs of the reference's own packer
    (quick/awq/modules/linear/quick.py:60-156) exec'd on CPU by
    ``tests/golden/make_golden.py`` (committed); ``pack_quick`` must reproduce
    them bit-for-bit (tests/test_oracle.py).
  * ``kernel_view_w16`` re-derives W16[k][n] from the *kernel's* pointer math
    (csrc/gemm_cuda_quick.cu:1353-1377, :29, :243, :52-60 and
    csrc/dequantize_quick.cuh:35-60) independently of the packer; it must agree
    with ``dequant_w16(unpack_quick(...))`` exactly.
On the GPU box the unmodified reference kernel itself (oracle/_ref) is the
final pin (tests/test_gpu_parity.py).

Notation: q[k][n] in 0..15 (k input channel, n output channel), z[g][n] in
0..15, s[g][n] fp16, g = k // G.
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------
# Packed layout (reference: quick.py:88-150; closed form SURVEY.md Appendix A)
# ---------------------------------------------------------------------------

# nibble p of a qweight word holds q[k0 + DK[p]][c0 + DC[p]]
DK = np.array([0, 8, 0, 8, 1, 9, 1, 9], dtype=np.int64)
DC = np.array([0, 0, 8, 8, 0, 0, 8, 8], dtype=np.int64)


def _qweight_index_grids(K: int, N: int):
    """(k, n) coordinates of every nibble of every qweight word.

    Returns k_idx, n_idx with shape (K*N//8, 8): flat word index f, nibble p.
    Follows the kernel's B pointer math (gemm_cuda_quick.cu:1354, :1372) for the
    word index and the mma.m16n8k16 B-fragment ownership for (k0, c0).
    """
    assert K % 32 == 0 and N % 128 == 0
    f = np.arange(K * N // 8, dtype=np.int64)
    kt, r = np.divmod(f, 4 * N)
    r4, c = np.divmod(r, N)
    bx, w = np.divmod(c, 128)
    lane = 16 * (r4 % 2) + w // 8
    ty = r4 // 2
    ks = (w % 8) // 4
    ch = w % 4
    k0 = 32 * kt + 16 * ks + 2 * (lane % 4)
    c0 = 128 * bx + 64 * ty + 16 * ch + lane // 4
    k_idx = k0[:, None] + DK[None, :]
    n_idx = c0[:, None] + DC[None, :]
    return k_idx, n_idx


def slot_to_column(N: int) -> np.ndarray:
    """Column n held by scale/zero slot x (x in 0..N-1) of a packed row.

    Reference: quick.py:125-128 (scales) and :137-140 (zeros) use the same map;
    kernel side gemm_cuda_quick.cu:1356-1358.
    """
    x = np.arange(N, dtype=np.int64)
    nb = N // 128
    ty = x // (N // 2)
    lh = (x // (N // 4)) % 2
    bx = (x // 32) % nb
    j4 = (x % 32) // 8
    m = x % 8
    return 128 * bx + 64 * ty + 16 * (m // 2) + 8 * (m % 2) + 4 * lh + j4


def pack_quick(q: np.ndarray, z: np.ndarray, s: np.ndarray):
    """(q[K,N], z[K/G,N], s[K/G,N] fp16) -> reference buffers.

    qweight int32 (K/4, N/2); qzeros int32 (K/G, N/4); scales fp16 (K/G, 2N)
    (shapes: quick.py:52-54).
    """
    K, N = q.shape
    k_idx, n_idx = _qweight_index_grids(K, N)
    nib = q[k_idx, n_idx].astype(np.uint32)
    words = np.zeros(nib.shape[0], dtype=np.uint32)
    for p in range(8):
        words |= nib[:, p] << np.uint32(4 * p)
    qweight = words.view(np.int32).reshape(K // 4, N // 2)

    col = slot_to_column(N)
    scales = np.repeat(np.asarray(s, dtype=np.float16)[:, col], 2, axis=1)
    z4 = z[:, col].astype(np.uint32).reshape(z.shape[0], N // 4, 4)
    zw = z4[:, :, 0] | (z4[:, :, 1] << 4) | (z4[:, :, 2] << 8) | (z4[:, :, 3] << 12)
    zw = zw | (zw << 16)
    qzeros = zw.astype(np.uint32).view(np.int32).reshape(z.shape[0], N // 4)
    return qweight, qzeros, np.ascontiguousarray(scales)


def unpack_quick(qweight: np.ndarray, qzeros: np.ndarray, scales: np.ndarray):
    """Inverse of ``pack_quick``. Returns q, z, s."""
    K = qweight.shape[0] * 4
    N = qweight.shape[1] * 2
    NG = qzeros.shape[0]
    words = np.ascontiguousarray(qweight).view(np.uint32).reshape(-1)
    k_idx, n_idx = _qweight_index_grids(K, N)
    q = np.zeros((K, N), dtype=np.int32)
    for p in range(8):
        q[k_idx[:, p], n_idx[:, p]] = (words >> np.uint32(4 * p)) & np.uint32(0xF)
    col = slot_to_column(N)
    s = np.zeros((NG, N), dtype=np.float16)
    s[:, col] = np.asarray(scales)[:, 0::2]
    zw = np.ascontiguousarray(qzeros).view(np.uint32).reshape(NG, N // 4)
    z = np.zeros((NG, N), dtype=np.int32)
    for i in range(4):
        z[:, col[i::4]] = (zw >> np.uint32(4 * i)) & np.uint32(0xF)
    return q, z, s


# ---------------------------------------------------------------------------
# Bit-level dequantisation (reference: csrc/dequantize_quick.cuh:15-63)
# ---------------------------------------------------------------------------

def s4_to_fp16x2_fused(word: np.ndarray) -> np.ndarray:
    """uint32 words -> (..., 8) fp16 in register order x.lo x.hi y.lo y.hi z.lo z.hi w.lo w.hi.

    Every output equals 1024 + nibble; register j (x,y,z,w) = (nibble j, nibble j+4).
    lop3 with immLut (0xf0&0xcc)|0xaa is (a & b) | c  (dequantize_quick.cuh:22,36-51);
    the top nibbles go through fma.rn.f16x2(h, 1/16, 960) (:54-60).
    """
    w = np.asarray(word, dtype=np.uint32)
    top = w >> np.uint32(8)
    h0 = (w & np.uint32(0x000F000F)) | np.uint32(0x64006400)
    h1 = (w & np.uint32(0x00F000F0)) | np.uint32(0x64006400)
    h2 = (top & np.uint32(0x000F000F)) | np.uint32(0x64006400)
    h3 = (top & np.uint32(0x00F000F0)) | np.uint32(0x64006400)

    def halves(h):
        lo = (h & np.uint32(0xFFFF)).astype(np.uint16).view(np.float16)
        hi = (h >> np.uint32(16)).astype(np.uint16).view(np.float16)
        return lo, hi

    out = np.zeros(w.shape + (8,), dtype=np.float16)
    sixteenth = np.float32(0.0625)
    for j, h in enumerate((h0, h1, h2, h3)):
        lo, hi = halves(h)
        if j in (1, 3):  # fma.rn.f16x2 (exact here: integers < 2048)
            lo = (lo.astype(np.float32) * sixteenth + np.float32(960.0)).astype(np.float16)
            hi = (hi.astype(np.float32) * sixteenth + np.float32(960.0)).astype(np.float16)
        out[..., 2 * j] = lo
        out[..., 2 * j + 1] = hi
    return out


def kernel_view_w16(qweight: np.ndarray, qzeros: np.ndarray, scales: np.ndarray, G: int) -> np.ndarray:
    """W16[k][n] exactly as the reference kernel materialises it in registers.

    Walks the kernel's own addressing, not the packer's: per (k32-tile kt, CTA
    bx, warp ty, lane l) the thread loads 8 words at
    B + kt*4N + (2ty + l/16)*N + bx*128 + (l%16)*8 (gemm_cuda_quick.cu:1354,
    :1372, :29, :243), one uint2 of zeros and two uint4 of scales for its
    ``channel`` (:1355-1358, :1373-1377), dequantises each word with
    ``s4_to_fp16x2_fused``, then sub.f16x2 / mul.rn.f16x2 (:52-60) and feeds
    registers (.x,.y) / (.z,.w) as mma.m16n8k16 B fragments of the two n8 tiles
    of n16-chunk ch (:63-97).  B fragment ownership (PTX ISA, m16n8k16 .col B):
    reg0 = rows 2*(l%4)+{0,1}, reg1 = rows 2*(l%4)+8+{0,1}, column l/4.
    """
    K = qweight.shape[0] * 4
    N = qweight.shape[1] * 2
    B = np.ascontiguousarray(qweight).view(np.uint32).reshape(-1)
    Z = np.ascontiguousarray(qzeros).view(np.uint32).reshape(-1)
    S = np.ascontiguousarray(scales).view(np.float16).reshape(-1)
    kt = np.arange(K // 32).reshape(-1, 1, 1, 1, 1, 1)
    bx = np.arange(N // 128).reshape(1, -1, 1, 1, 1, 1)
    ty = np.arange(2).reshape(1, 1, -1, 1, 1, 1)
    l = np.arange(32).reshape(1, 1, 1, -1, 1, 1)
    ks = np.arange(2).reshape(1, 1, 1, 1, -1, 1)
    ch = np.arange(4).reshape(1, 1, 1, 1, 1, -1)
    shape = np.broadcast_shapes(kt.shape, bx.shape, ty.shape, l.shape, ks.shape, ch.shape)

    channel = ty * (N // 8) * 2 + (l // 16) * (N // 8) + bx * 16 + (l % 16)
    word_idx = channel * 8 + kt * 32 * (N // 8) + ks * 4 + ch
    words = B[np.broadcast_to(word_idx, shape)]
    d = s4_to_fp16x2_fused(words)  # (..., 8): x.lo x.hi y.lo y.hi z.lo z.hi w.lo w.hi

    g = (kt * 32) // G
    zero_base = (channel // 4) * 2 + g * (N // 8) * 2          # int32 index of uint2
    scale_base = (channel // 4) * 16 + g * N * 2               # half index of first uint4
    W = np.zeros((K, N), dtype=np.float16)
    for e in range(2):  # e = 0: regs (.x,.y) -> n8 tile c0 ; e = 1: regs (.z,.w) -> c1 = c0 + 8
        m = 2 * ch + e  # half2 index among the 8 zero/scale half2s
        zw = Z[np.broadcast_to(zero_base + m // 4, shape)]
        zd = s4_to_fp16x2_fused(zw)  # half2 j = (nibble j, nibble j+4)
        mm = np.broadcast_to(m % 4, shape)
        z_lo = np.take_along_axis(zd, (2 * mm)[..., None], axis=-1)[..., 0]
        z_hi = np.take_along_axis(zd, (2 * mm + 1)[..., None], axis=-1)[..., 0]
        s_lo = S[np.broadcast_to(scale_base + 2 * m, shape)]
        s_hi = S[np.broadcast_to(scale_base + 2 * m + 1, shape)]
        col = np.broadcast_to(128 * bx + 64 * ty + 16 * ch + 8 * e + l // 4, shape)
        for reg in range(2):  # reg0: k rows +0/+1 ; reg1: +8/+9
            for half in range(2):
                v = d[..., 4 * e + 2 * reg + half]
                zz = z_lo if half == 0 else z_hi
                ss = s_lo if half == 0 else s_hi
                diff = (v.astype(np.float32) - zz.astype(np.float32)).astype(np.float16)   # sub.f16x2
                w16 = (diff.astype(np.float32) * ss.astype(np.float32)).astype(np.float16)  # mul.rn.f16x2
                krow = np.broadcast_to(32 * kt + 16 * ks + 2 * (l % 4) + 8 * reg + half, shape)
                W[krow, col] = w16
    return W


# ---------------------------------------------------------------------------
# Arithmetic
# ---------------------------------------------------------------------------

def dequant_w16(q: np.ndarray, z: np.ndarray, s: np.ndarray, G: int) -> np.ndarray:
    """W16 = fp16_rn(fp16(q - z) * s): exact difference, one rounding (…cu:53-54)."""
    zr = np.repeat(z, G, axis=0)
    sr = np.repeat(np.asarray(s, dtype=np.float16), G, axis=0)
    diff = (q.astype(np.int32) - zr.astype(np.int32)).astype(np.float16)
    return (diff.astype(np.float32) * sr.astype(np.float32)).astype(np.float16)


def gemm_exact(A16: np.ndarray, W16: np.ndarray) -> np.ndarray:
    """fp64 A·W16 — the value every implementation approximates."""
    return A16.astype(np.float64) @ W16.astype(np.float64)


def gemm_oracle(A16: np.ndarray, W16: np.ndarray) -> np.ndarray:
    """fp32-accumulate, single final fp16 rounding (what the new kernel does)."""
    return (A16.astype(np.float32) @ W16.astype(np.float32)).astype(np.float16)


def gemm_oracle_splitk(A16: np.ndarray, W16: np.ndarray, split_k: int) -> np.ndarray:
    """The reference's split-K semantics: k32-tile t -> split t % split_k
    (…cu:1366), each split rounded to fp16 (:1391-1394), then sum(0) with fp32
    accumulation and one rounding (:1515)."""
    M, K = A16.shape
    N = W16.shape[1]
    At = A16.astype(np.float32).reshape(M, K // 32, 32)
    Wt = W16.astype(np.float32).reshape(K // 32, 32, N)
    parts = []
    for i in range(split_k):
        a = At[:, i::split_k].reshape(M, -1)
        w = Wt[i::split_k].reshape(-1, N)
        parts.append((a @ w).astype(np.float16))
    return np.sum(np.stack(parts).astype(np.float32), axis=0).astype(np.float16)


def reference_output_shape(M: int, N: int, split_k: int):
    """(1,M,N) when split_k == 1, else (M,N)  (…cu:1515-1516)."""
    return (1, M, N) if split_k == 1 else (M, N)


def check_args(K: int, N: int, G: int):
    """The reference's three argument checks (…cu:1479-1484) -> ValueError."""
    if N % 128 != 0:
        raise ValueError("OC is not multiple of cta_N = 128")
    if N % 8 != 0:
        raise ValueError("OC is not multiple of pack_num = 8")
    if G % 32 != 0:
        raise ValueError("Group size should be a multiple of 32")


def forward_oracle(x16, qweight, qzeros, scales, bias=None):
    """WQLinear_QUICK.forward (quick.py:158-166) on CPU: flatten, GEMM, bias, reshape."""
    K = qweight.shape[0] * 4
    N = qweight.shape[1] * 2
    G = K // qzeros.shape[0]
    q, z, s = unpack_quick(qweight, qzeros, scales)
    out = gemm_oracle(np.asarray(x16, dtype=np.float16).reshape(-1, K), dequant_w16(q, z, s, G))
    if bias is not None:
        out = (out.astype(np.float32) + np.asarray(bias, np.float16).astype(np.float32)).astype(np.float16)
    return out.reshape(x16.shape[:-1] + (N,))


# ---------------------------------------------------------------------------
# Layout algebra (reference: fused_utils.py:119-159)
# ---------------------------------------------------------------------------

def quick_cat(tensors, options: str) -> np.ndarray:
    """N-concatenation of packed tensors (QUICK_cat, fused_utils.py:146-157),
    generalised to unequal widths (the reference rejects those, :139-142)."""
    H = tensors[0].shape[0]
    if options == "qweight":
        rows = H // 2
    elif options in ("qzeros", "scales"):
        rows = H * 4
    else:
        raise ValueError("Unknown options provided or invalid reshape dimensions")
    return np.concatenate([t.reshape(rows, -1) for t in tensors], axis=1).reshape(H, -1)


def shard_columns(qweight, qzeros, scales, rank: int, world: int):
    """Column-parallel shard = inverse of quick_cat."""
    K = qweight.shape[0] * 4
    N = qweight.shape[1] * 2
    NG = qzeros.shape[0]
    assert (N // world) % 128 == 0
    n0, n1 = rank * N // world, (rank + 1) * N // world
    qw = qweight.reshape(K // 8, N)[:, n0:n1].reshape(K // 4, -1)
    sc = scales.reshape(4 * NG, N // 2)[:, n0 // 2:n1 // 2].reshape(NG, -1)
    qz = qzeros.reshape(4 * NG, N // 16)[:, n0 // 16:n1 // 16].reshape(NG, -1)
    return np.ascontiguousarray(qw), np.ascontiguousarray(qz), np.ascontiguousarray(sc)


# ---------------------------------------------------------------------------
# Synthetic inputs
# ---------------------------------------------------------------------------

AWQ_ORDER = (0, 2, 4, 6, 1, 3, 5, 7)


def unpack_awq_gemm(qweight: np.ndarray, qzeros: np.ndarray):
    """AWQ "GEMM" layout -> logical (q [K,N], z [K/G,N]) uint8.  Restates the reference's
    unpack_awq (quick/awq/utils/packing_utils.py:8-25: nibble i of word c -> unpacked column 8c+i),
    reverse_awq_order (:28-39: column 8c+j of the result = unpacked column 8c+AWQ_REVERSE_ORDER[j]) and
    the 4-bit mask (:84-85); equivalently nibble i holds logical column 8c + AWQ_ORDER[i], the packer's
    order_map (quick/awq/modules/linear/gemm.py:117-123)."""
    def unpack(t):
        w = t.astype(np.int64) & 0xFFFFFFFF
        out = np.empty(t.shape + (8,), dtype=np.uint8)
        for i in range(8):
            out[..., AWQ_ORDER[i]] = (w >> (4 * i)) & 0xF
        return out.reshape(t.shape[0], -1)
    return unpack(qweight), unpack(qzeros)


def pack_awq_gemm(q: np.ndarray, z: np.ndarray):
    """Logical -> AWQ "GEMM" layout (reference packer loops, gemm.py:108-143, in closed form)."""
    def pack(t):
        t8 = (t.astype(np.int64) & 0xF).reshape(t.shape[0], -1, 8)
        w = np.zeros(t8.shape[:2], dtype=np.int64)
        for i in range(8):
            w |= t8[:, :, AWQ_ORDER[i]] << (4 * i)
        return w.astype(np.uint32).view(np.int32)
    return pack(q), pack(z)


def make_case(K: int, N: int, G: int, seed: int = 1234):
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 16, size=(K, N), dtype=np.int32)
    z = rng.integers(0, 16, size=(K // G, N), dtype=np.int32)
    s = (0.002 + 0.01 * rng.random((K // G, N))).astype(np.float16)
    return q, z, s


def make_activations(M: int, K: int, seed: int):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((M, K)).astype(np.float16)
