"""Minimal Llama-like decoder whose every decoder-layer linear is a WQLinear_QUICK — SURVEY §8(f1).

Only what is needed to report tokens/s the way the reference's examples/benchmark.py does
(prefill + token-by-token decode with a KV cache, benchmark.py:38-67): embedding, RMSNorm, fused QKV
(q‖k‖v concatenated in the QUICK layout — the generalised QUICK_cat, so GQA works), RoPE, a static KV
cache, torch SDPA attention, SwiGLU MLP with gate‖up fused into one GEMM, fp16 lm_head (the reference
leaves lm_head unquantised, base.py:396-405).  Reference structure mirrored: modules/fused/model.py:63-109,
block.py:39-74, attn.py:100-245, cache.py:3-58.  Weights are random-init (no checkpoints / network here).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

import quick_kernels

from ..modules.linear.quick import WQLinear_QUICK

# Decoder-layer glue (RMSNorm, rotary + KV-cache update, SiLU·up, residual add) through the library's fused kernels
# (C-ABI qb200_rmsnorm / qb200_rope_kv_update / qb200_silu_mul / qb200_gemm_w4a16_fused) instead of ~30 small torch
# kernels per layer; FUSED_GLUE = False keeps the plain torch expressions (the parity reference of the tests).
FUSED_GLUE = True
# Decode steps (one new token): rotary embedding + KV-cache update + attention over the cache in ONE kernel
# (qb200_attn_decode) instead of rope_kv_update + mask arithmetic + torch SDPA.  QB200_ATTN_DECODE=0 keeps the SDPA path.
import os as _os0
ATTN_DECODE = _os0.environ.get("QB200_ATTN_DECODE", "1") != "0"
# Measured on B200, Llama-2-7B shapes (round 2, cluster flash-decoding kernel; gpurun logs under profiles/r2b_*): against
# the rope_kv_update + SDPA path the decode step is 17 % faster at batch 1, 10 % at 32, 6 % at 64 (cache 256), 8 % faster
# at batch 1 with a 2048-position cache but 6 % slower at batch 8 there (a warp's serial chain of row batches grows with
# the cache; cuDNN's split wins once there are thousands of long rows) -> long caches keep SDPA above a few sequences.
ATTN_DECODE_MAX_BATCH = int(_os0.environ.get("QB200_ATTN_DECODE_MAX_BATCH", "64"))
# grouped-query shapes (>= 4 query heads per kv head: the fp32 score path is instruction-bound there; Mistral-7B at batch
# 64: -6 % against SDPA, +2.5 % at 32)
ATTN_DECODE_MAX_BATCH_GQA = int(_os0.environ.get("QB200_ATTN_DECODE_MAX_BATCH_GQA", "32"))
ATTN_DECODE_MAX_CACHE = int(_os0.environ.get("QB200_ATTN_DECODE_MAX_CACHE", "1024"))          # ... for batches above
ATTN_DECODE_LONG_CACHE_BATCH = int(_os0.environ.get("QB200_ATTN_DECODE_LONG_CACHE_BATCH", "4"))  # ... this many sequences
# RMSNorm folded around the GEMMs (qb200_gemm_w4a16_norm): o_proj / down_proj also emit h * gamma and the rows' sums of
# squares, q|k|v and gate|up scale their rows by 1/rms — the two RMSNorm kernels of a layer disappear (single GPU;
# the first layer's norm_1 and the final norm stay kernels).  QB200_NORM_FUSION=0 keeps the qb200_rmsnorm kernels.
NORM_FUSION = _os0.environ.get("QB200_NORM_FUSION", "1") != "0"
# Prefill through a CUDA graph once a (batch, length) shape repeats (forward()); QB200_PREFILL_GRAPH=0 keeps it eager.
PREFILL_GRAPH = _os0.environ.get("QB200_PREFILL_GRAPH", "1") != "0"
PREFILL_GRAPH_MAX_ROWS = int(_os0.environ.get("QB200_PREFILL_GRAPH_MAX_ROWS", "4096"))   # above: GPU-bound anyway
PREFILL_GRAPH_KEEP = 4
# SiLU(gate)·up inside the gate|up GEMM's epilogue (QB200_GEMM_SILU_MUL) instead of a separate qb200_silu_mul kernel.
FUSED_SILU = _os0.environ.get("QB200_FUSED_SILU", "1") != "0"


@dataclass
class LlamaLikeConfig:
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_layers: int = 32
    num_heads: int = 32
    num_kv_heads: int = 32
    vocab_size: int = 32000
    max_seq_len: int = 256
    group_size: int = 128
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0

    @property
    def head_dim(self):
        return self.hidden_size // self.num_heads


PRESETS = {
    "llama-2-7b": LlamaLikeConfig(4096, 11008, 32, 32, 32),
    "mistral-7b": LlamaLikeConfig(4096, 14336, 32, 32, 8),
    "llama-2-70b": LlamaLikeConfig(8192, 28672, 80, 64, 8),
    "tiny": LlamaLikeConfig(512, 1024, 2, 8, 4, vocab_size=1024, max_seq_len=64),
}


def random_quick_linear(in_f: int, out_f: int, G: int, dev, gen: torch.Generator, linear_impl: str = "quick_b200"):
    """Random-init packed weights straight in the QUICK layout (any nibble pattern is a valid weight; zeros and
    scales are written with the duplication the format requires, quick.py:129-130,141-150)."""
    m = WQLinear_QUICK(4, G, in_f, out_f, False, dev)
    m.qweight = torch.randint(-2 ** 31, 2 ** 31 - 1, m.qweight.shape, dtype=torch.int32, device=dev, generator=gen)
    z4 = torch.randint(0, 2 ** 16, m.qzeros.shape, dtype=torch.int32, device=dev, generator=gen)
    m.qzeros = z4 | (z4 << 16)
    s = (torch.rand((in_f // G, out_f), device=dev, generator=gen) * 0.004 + 0.001).half()
    m.scales = s.repeat_interleave(2, dim=1).contiguous()
    m.linear_impl = linear_impl
    return m


class RMSNorm(nn.Module):
    def __init__(self, dim, eps, dev):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim, dtype=torch.float16, device=dev), requires_grad=False)
        self.eps = eps

    def forward(self, x):
        if x.is_cuda and FUSED_GLUE:
            return quick_kernels.rmsnorm(x, self.weight, self.eps)      # one kernel instead of seven
        return self.forward_torch(x)

    def forward_torch(self, x):
        v = x.float()
        return (v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + self.eps)).half() * self.weight


def tp_world() -> int:
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _tp_rank() -> int:
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _from_packed(like: WQLinear_QUICK, packed, n_out: int, bias=None) -> WQLinear_QUICK:
    out = WQLinear_QUICK(like.w_bit, like.group_size, packed[0].shape[0] * 4, n_out, False, "meta", like.k_split_1, like.k_split_2)
    out.qweight, out.qzeros, out.scales = packed
    out.bias = bias
    return out


def shard_quick_linear(m: WQLinear_QUICK, rank: int, world: int) -> WQLinear_QUICK:
    """The column-parallel shard `rank` of `world` of a packed linear: output columns [rank·N/world, (rank+1)·N/world)
    as a WQLinear_QUICK of its own (layout.shard_columns: the inverse of QUICK_cat; N/world must stay a multiple of
    the 128-column tile)."""
    from ...layout import shard_columns
    if m.out_features % (128 * world) != 0:
        raise ValueError(f"N={m.out_features} does not split into {world} shards of 128-column tiles")
    n_local = m.out_features // world
    bias = None if m.bias is None else m.bias[rank * n_local:(rank + 1) * n_local].clone()
    return _from_packed(m, shard_columns(m.qweight, m.qzeros, m.scales, rank, world), n_local, bias)


def tp_shard_qkv(m: WQLinear_QUICK, nh: int, nkv: int, hd: int, rank: int, world: int) -> WQLinear_QUICK:
    """Head-parallel shard of a fused q‖k‖v projection: rank r keeps [q heads of r | k heads of r | v heads of r], so
    attention is local to the rank and q‖k‖v needs no gather at all ('qkv -> attention is head-local')."""
    from ...layout import quick_cat, slice_columns
    if nh % world or nkv % world:
        raise ValueError(f"{nh} query / {nkv} key-value heads do not split over {world} ranks")
    nh_l, nkv_l = nh // world, nkv // world
    runs = [(rank * nh_l * hd, nh_l * hd), (nh * hd + rank * nkv_l * hd, nkv_l * hd), ((nh + nkv) * hd + rank * nkv_l * hd, nkv_l * hd)]
    if any(w % 128 or s0 % 128 for s0, w in runs):
        raise ValueError(f"heads per rank x head_dim ({nh_l}x{hd}, {nkv_l}x{hd}) must be multiples of the 128-column tile")
    parts = [slice_columns(m.qweight, m.qzeros, m.scales, s0, s0 + w) for s0, w in runs]
    packed = tuple(quick_cat([p[i] for p in parts], opt).contiguous() for i, opt in enumerate(("qweight", "qzeros", "scales")))
    bias = None if m.bias is None else torch.cat([m.bias[s0:s0 + w] for s0, w in runs]).clone()
    return _from_packed(m, packed, sum(w for _, w in runs), bias)


def tp_pad_intermediate(I: int, world: int) -> int:
    """The MLP width a tensor-parallel model runs with: the next multiple of 128 x world (zero-weight channels: exact)."""
    step = 128 * world
    return (I + step - 1) // step * step


def tp_shard_gate_up(m: WQLinear_QUICK, I: int, rank: int, world: int) -> WQLinear_QUICK:
    """Rank r's [gate columns of r | up columns of r] of a fused gate‖up projection (both halves padded with zero-weight
    channels to a multiple of 128 x world first): SiLU(gate)·up is then local to the rank, and only the activation —
    half as wide as gate‖up — is gathered for the down projection."""
    from ...layout import pad_columns, quick_cat, slice_columns
    Ip = tp_pad_intermediate(I, world)
    I_l = Ip // world
    halves = []
    for h in range(2):
        part = slice_columns(m.qweight, m.qzeros, m.scales, h * I, (h + 1) * I) if I % 128 == 0 else None
        if part is None:
            raise ValueError(f"intermediate size {I} is not a multiple of the 128-column tile")
        part = pad_columns(*part, Ip)
        halves.append(slice_columns(*part, rank * I_l, (rank + 1) * I_l))
    packed = tuple(quick_cat([p[i] for p in halves], opt).contiguous() for i, opt in enumerate(("qweight", "qzeros", "scales")))
    bias = None
    if m.bias is not None:
        pad = torch.zeros(Ip - I, dtype=m.bias.dtype, device=m.bias.device)
        bias = torch.cat([torch.cat([m.bias[h * I:(h + 1) * I], pad])[rank * I_l:(rank + 1) * I_l] for h in range(2)]).clone()
    return _from_packed(m, packed, 2 * I_l, bias)


def tp_shard_down(m: WQLinear_QUICK, I: int, rank: int, world: int) -> WQLinear_QUICK:
    """Column-parallel shard of the down projection whose input channels are padded like tp_shard_gate_up pads them."""
    from ...layout import pad_rows
    Ip = tp_pad_intermediate(I, world)
    padded = _from_packed(m, pad_rows(m.qweight, m.qzeros, m.scales, Ip, m.group_size), m.out_features, m.bias)
    return shard_quick_linear(padded, rank, world)


# Tensor-parallel exchange: "peer" = every producer stores its column slab into all ranks' buffers over NVLink and the
# hand-over rides in the kernels themselves (quick_b200.parallel.GatheredBuffer, include/quick_b200.h qb200_peer_*);
# "nccl" = the same dataflow with torch.distributed all-gathers (the baseline, and what runs on CPU / gloo).
import os as _os
TP_MODE = _os.environ.get("QB200_TP_MODE", "peer")


class TensorParallel:
    """Per-model tensor-parallel state: rank / world and, in peer mode, the four gathered buffers a decoder layer
    alternates — attention output, hidden state after the attention block (B), MLP activation, hidden state after the
    MLP (A).  Between two fills of one buffer every rank fills the three others, which is what makes reuse safe
    (include/quick_b200.h)."""

    def __init__(self, cfg, batch: int, dev):
        import torch.distributed as dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.mode = TP_MODE if torch.device(dev).type == "cuda" else "nccl"
        self.I_pad = tp_pad_intermediate(cfg.intermediate_size, self.world)
        if self.mode == "peer":
            from ...parallel import GatheredBuffer
            rows = batch * cfg.max_seq_len
            self.attn = GatheredBuffer(rows, cfg.num_heads * cfg.head_dim)
            self.hid_b = GatheredBuffer(rows, cfg.hidden_size)
            self.act = GatheredBuffer(rows, self.I_pad)
            self.hid_a = GatheredBuffer(rows, cfg.hidden_size)

    def all_gather_cols(self, local: torch.Tensor) -> torch.Tensor:
        """(M, n/R) slabs of every rank -> (M, n) in rank order: ONE all-gather (NCCL over NVLink, or gloo)."""
        import torch.distributed as dist
        flat = local.contiguous()
        gathered = torch.empty((self.world * flat.shape[0], flat.shape[1]), dtype=flat.dtype, device=flat.device)
        dist.all_gather_into_tensor(gathered, flat)
        return gathered.view(self.world, flat.shape[0], flat.shape[1]).permute(1, 0, 2).reshape(flat.shape[0], -1)


def _linear(m: WQLinear_QUICK, x, ref_mod=None, residual=None):
    """Route through the B200 kernel, or (baseline runs only) through the unmodified reference kernel.
    residual: returns residual + linear(x) (fused into the GEMM epilogue)."""
    if ref_mod is None:
        return m(x, residual if FUSED_GLUE else None) if (residual is None or FUSED_GLUE) else residual + m(x)
    split = m.k_split_1 if m.out_features > m.in_features else m.k_split_2      # quick.py:161-164
    out = ref_mod.gemm_forward_cuda_quick(x.reshape(-1, x.shape[-1]), m.qweight, m.scales, m.qzeros, split)
    out = out.reshape(x.shape[:-1] + (m.out_features,))
    return out if residual is None else residual + out


class NormCarry:
    """What a layer's down_proj hands to the next layer when the RMSNorm is folded around the GEMMs: the hidden state
    scaled by the next norm_1 weight and its rows' per-tile sums of squares (WQLinear_QUICK.forward_norm_out)."""
    __slots__ = ("scaled", "ssq")

    def __init__(self, scaled, ssq):
        self.scaled, self.ssq = scaled, ssq


class Block(nn.Module):
    next_norm = None      # norm_1 of the following layer (set by the model; None for the last layer)

    def __init__(self, cfg: LlamaLikeConfig, dev, gen, batch: int, tp: Optional["TensorParallel"] = None, parts=None):
        """parts: {"qkv_proj", "o_proj", "gate_up_proj", "down_proj": WQLinear_QUICK, "norm_1", "norm_2": fp16 weight}
        taken from a loaded checkpoint (fuse_hf_model); without it the weights are random-init.
        tp: tensor-parallel state (None on one GPU).  Under tensor parallelism this rank holds: q‖k‖v of ITS heads
        (attention and the KV cache are head-sharded, nothing is gathered), N/R output columns of o_proj and down_proj
        (column-parallel, slabs gathered), and [gate | up] of ITS slice of the MLP width (SiLU·up is local, the
        activation is gathered for down_proj)."""
        super().__init__()
        self.cfg, self.tp = cfg, tp
        R = tp.world if tp is not None else 1
        rank = tp.rank if tp is not None else 0
        hd, nh, nkv = cfg.head_dim, cfg.num_heads, cfg.num_kv_heads
        H, I = cfg.hidden_size, cfg.intermediate_size
        if nh % R or nkv % R:
            raise ValueError(f"{nh} query / {nkv} key-value heads do not split over {R} ranks")
        self.nh_l, self.nkv_l = nh // R, nkv // R
        self.I_l = (tp.I_pad if tp is not None else I) // R
        self.norm_1 = RMSNorm(H, cfg.rms_eps, dev)
        self.norm_2 = RMSNorm(H, cfg.rms_eps, dev)
        if parts is not None:
            self.norm_1.weight.data = parts["norm_1"].detach().to(dev, torch.float16).contiguous()
            self.norm_2.weight.data = parts["norm_2"].detach().to(dev, torch.float16).contiguous()
            expect = {"qkv_proj": (H, (nh + 2 * nkv) * hd), "o_proj": (nh * hd, H), "gate_up_proj": (H, 2 * I), "down_proj": (I, H)}
            for name, (k, n) in expect.items():
                m = parts[name]
                if (m.in_features, m.out_features) != (k, n):
                    raise ValueError(f"{name}: ({m.in_features} -> {m.out_features}) does not match the config ({k} -> {n})")
                if R > 1:       # this rank's shard; the full tensors are freed by the caller
                    m = {"qkv_proj": lambda: tp_shard_qkv(m, nh, nkv, hd, rank, R), "o_proj": lambda: shard_quick_linear(m, rank, R),
                         "gate_up_proj": lambda: tp_shard_gate_up(m, I, rank, R), "down_proj": lambda: tp_shard_down(m, I, rank, R)}[name]()
                m.tp_sharded = R > 1
                setattr(self, name, m)
        else:
            # random-init: the rank's shard is generated directly (N/R must stay a multiple of the 128-column tile)
            def lin(in_f, out_f):
                assert out_f % 128 == 0, f"N={out_f} is not a multiple of the 128-column tile"
                m = random_quick_linear(in_f, out_f, cfg.group_size, dev, gen)
                m.tp_sharded = R > 1
                return m

            self.qkv_proj = lin(H, (self.nh_l + 2 * self.nkv_l) * hd)
            self.o_proj = lin(nh * hd, H // R)
            self.gate_up_proj = lin(H, 2 * self.I_l)
            self.down_proj = lin(self.I_l * R, H // R)
        self.register_buffer("cache_k", torch.zeros(batch, self.nkv_l, cfg.max_seq_len, hd, dtype=torch.float16, device=dev), persistent=False)
        self.register_buffer("cache_v", torch.zeros(batch, self.nkv_l, cfg.max_seq_len, hd, dtype=torch.float16, device=dev), persistent=False)

    def _attention(self, qkv, cos, sin, pos_idx, attn_mask, rope, fused_decode):
        """qkv (B, T, (nh_l + 2 nkv_l) hd) of this rank's heads -> attention output (B, T, nh_l hd); updates the cache."""
        B, T, _ = qkv.shape
        hd, nh, nkv = self.cfg.head_dim, self.nh_l, self.nkv_l
        if fused_decode:
            # decode step on the fused path (the model decided: see LlamaLikeQuickModel._fused_decode_ok)
            return quick_kernels.attn_decode(qkv, rope[0], rope[1], pos_idx, self.cache_k, self.cache_v, nh, nkv)
        if FUSED_GLUE and qkv.is_cuda:
            # rotary embedding of q and k + KV-cache update in one kernel (rope tables are indexed by position)
            q = quick_kernels.rope_kv_update(qkv, rope[0], rope[1], pos_idx, self.cache_k, self.cache_v, nh, nkv)
        else:
            q, k, v = qkv.split([nh * hd, nkv * hd, nkv * hd], dim=-1)
            q = q.view(B, T, nh, hd).transpose(1, 2)
            k = k.view(B, T, nkv, hd).transpose(1, 2)
            v = v.view(B, T, nkv, hd).transpose(1, 2)
            q, k = _rope(q, cos, sin), _rope(k, cos, sin)
            self.cache_k.index_copy_(2, pos_idx, k)
            self.cache_v.index_copy_(2, pos_idx, v)
        o = F.scaled_dot_product_attention(q, self.cache_k, self.cache_v, attn_mask=attn_mask, enable_gqa=(nkv != nh))
        return o.transpose(1, 2).reshape(B, T, nh * hd)

    def forward(self, x, cos, sin, pos_idx, attn_mask, ref_mod=None, rope=None, x_src=None):
        """Returns (hidden state, the gathered buffer it lives in or None)."""
        if self.tp is not None:
            return self.forward_tp(x, x_src, cos, sin, pos_idx, attn_mask, rope)
        cfg = self.cfg
        B, T, _ = x.shape
        fused_decode = T == 1 and attn_mask is None
        # RMSNorm folded around the GEMMs: x_src carries (h * gamma_1, sums of squares) from the previous layer's down_proj
        norm_fusion = NORM_FUSION and FUSED_GLUE and FUSED_SILU and x.is_cuda and ref_mod is None
        if isinstance(x_src, NormCarry):
            qkv = self.qkv_proj.forward_normed(x_src.scaled, x_src.ssq, self.norm_1.eps)
        else:
            qkv = _linear(self.qkv_proj, self.norm_1(x), ref_mod)
        o = self._attention(qkv, cos, sin, pos_idx, attn_mask, rope, fused_decode)
        if norm_fusion:
            x, xg, ssq = self.o_proj.forward_norm_out(o, x, self.norm_2.weight)
            act = self.gate_up_proj.enable_silu_mul().forward_silu_mul(xg, ssq, self.norm_2.eps)
            if self.next_norm is not None:
                x, xg, ssq = self.down_proj.forward_norm_out(act, x, self.next_norm.weight)
                return x, NormCarry(xg, ssq)
            return _linear(self.down_proj, act, ref_mod, residual=x), None
        x = _linear(self.o_proj, o, ref_mod, residual=x)
        if FUSED_GLUE and FUSED_SILU and x.is_cuda and ref_mod is None:
            # SiLU(gate)·up inside the gate|up GEMM's epilogue (gate / up channels interleaved once, at first use)
            act = self.gate_up_proj.enable_silu_mul().forward_silu_mul(self.norm_2(x))
        else:
            gu = _linear(self.gate_up_proj, self.norm_2(x), ref_mod)
            if FUSED_GLUE and gu.is_cuda:
                act = quick_kernels.silu_mul(gu)
            else:
                g, u = gu.split(cfg.intermediate_size, dim=-1)
                act = F.silu(g) * u
        return _linear(self.down_proj, act, ref_mod, residual=x), None

    def forward_tp(self, x, x_src, cos, sin, pos_idx, attn_mask, rope):
        """One decoder layer on this rank's shards.  Peer mode: 8 kernels (7 for a decode step), none of them a barrier — the two column-parallel
        GEMMs store their slabs into every rank's hidden-state buffer, the attention output and the MLP activation are
        scattered the same way, and each consumer (RMSNorm rows, GEMM activations) meets the producers in its own
        prologue.  NCCL mode: the same dataflow with four all-gathers."""
        tp, cfg = self.tp, self.cfg
        B, T, H = x.shape
        M = B * T
        hd = cfg.head_dim
        fused_decode = T == 1 and attn_mask is None
        x2d = x.reshape(M, H)
        col_h = tp.rank * (H // tp.world)
        if tp.mode == "peer":
            from ... import ops
            xn = ops.rmsnorm_tp(x2d, self.norm_1.weight, self.norm_1.eps, wait=x_src)
            wq, sz = self.qkv_proj._prepacked()
            qkv = ops.gemm_tp(xn, wq, sz, self.qkv_proj.out_features, self.qkv_proj.group_size, bias=self.qkv_proj.bias)
            if fused_decode:
                # decode step: rotary + KV-cache update + attention of the local heads, output stored straight into
                # every rank's attention buffer (no separate scatter kernel)
                ops.attn_decode_tp(qkv, rope[0], rope[1], pos_idx, self.cache_k, self.cache_v, self.nh_l, self.nkv_l, tp.attn,
                                   tp.rank * self.nh_l * hd)
            else:
                o = self._attention(qkv.view(B, T, -1), cos, sin, pos_idx, attn_mask, rope, fused_decode)
                ops.scatter_cols(o.reshape(M, self.nh_l * hd).contiguous(), tp.attn, tp.rank * self.nh_l * hd)
            wq, sz = self.o_proj._prepacked()
            ops.gemm_tp(tp.attn.rows(M), wq, sz, self.o_proj.out_features, self.o_proj.group_size, bias=self.o_proj.bias,
                        residual=x2d, dst=tp.hid_b, col0=col_h, wait=tp.attn)
            x2 = tp.hid_b.rows(M)
            xn2 = ops.rmsnorm_tp(x2, self.norm_2.weight, self.norm_2.eps, wait=tp.hid_b)
            gup = self.gate_up_proj.enable_silu_mul()      # [gate_r | up_r] -> interleaved channels, once
            wq, sz = gup._b200
            # SiLU(gate)·up in the epilogue, this rank's slice of the activation stored straight into every rank's buffer
            ops.gemm_tp(xn2, wq, sz, gup.out_features, gup.group_size, bias=gup._pair_bias, dst=tp.act, col0=tp.rank * self.I_l, silu_mul=True)
            wq, sz = self.down_proj._prepacked()
            ops.gemm_tp(tp.act.rows(M), wq, sz, self.down_proj.out_features, self.down_proj.group_size, bias=self.down_proj.bias,
                        residual=x2, dst=tp.hid_a, col0=col_h, wait=tp.act)
            return tp.hid_a.rows(M).view(B, T, H), tp.hid_a
        qkv = self.qkv_proj(self.norm_1(x2d))
        o = self._attention(qkv.view(B, T, -1), cos, sin, pos_idx, attn_mask, rope, fused_decode)
        o_full = tp.all_gather_cols(o.reshape(M, self.nh_l * hd))
        x2 = x2d + tp.all_gather_cols(self.o_proj(o_full))
        gu = self.gate_up_proj(self.norm_2(x2))
        if FUSED_GLUE and gu.is_cuda:
            act = quick_kernels.silu_mul(gu)
        else:
            g, u = gu.split(self.I_l, dim=-1)
            act = F.silu(g) * u
        out = x2 + tp.all_gather_cols(self.down_proj(tp.all_gather_cols(act)))
        return out.view(B, T, H), None


def _rope(t, cos, sin):
    t1, t2 = t[..., : t.shape[-1] // 2], t[..., t.shape[-1] // 2:]
    return (t * cos + torch.cat((-t2, t1), dim=-1) * sin).to(t.dtype)


class LlamaLikeQuickModel(nn.Module):
    def __init__(self, cfg: LlamaLikeConfig, batch: int, dev="cuda", seed: int = 0, parts=None):
        """parts: {"embed": nn.Embedding, "blocks": [Block parts], "norm": fp16 weight, "lm_head": nn.Linear} from a
        loaded checkpoint (fuse_hf_model); without it everything is random-init (benchmarks)."""
        super().__init__()
        self.cfg, self.batch = cfg, batch
        self._decode_graph = None
        self._prefill_graphs = {}         # (shape, all_logits) -> CUDA graph of a multi-token forward, see forward()
        self._attn_decode_supported = None
        self.start_pos = 0        # next free cache slot for the stateful HF-style calls (fuse_hf_model) and generate()
        import torch.distributed as dist
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        gen = torch.Generator(device=dev); gen.manual_seed(seed + 1000 * rank)   # every rank draws its own column slabs
        self.embed = parts["embed"] if parts is not None else nn.Embedding(cfg.vocab_size, cfg.hidden_size, device=dev, dtype=torch.float16)
        # tensor parallel (a process group exists): head-sharded attention, column-parallel o / down, local SiLU·up
        self.tp = TensorParallel(cfg, batch, dev) if tp_world() > 1 else None
        if parts is None:
            self.blocks = nn.ModuleList([Block(cfg, dev, gen, batch, self.tp) for _ in range(cfg.num_layers)])
            self.norm = RMSNorm(cfg.hidden_size, cfg.rms_eps, dev)
            self.lm_head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False, device=dev, dtype=torch.float16)
        else:
            self.blocks = nn.ModuleList([Block(cfg, dev, gen, batch, self.tp, parts=p) for p in parts["blocks"]])
            self.norm = RMSNorm(cfg.hidden_size, cfg.rms_eps, dev)
            self.norm.weight.data = parts["norm"].detach().to(dev, torch.float16).contiguous()
            self.lm_head = parts["lm_head"]
        for blk, nxt in zip(self.blocks[:-1], self.blocks[1:]):
            blk.__dict__["next_norm"] = nxt.norm_1      # plain reference (not a registered submodule): norm folded around the GEMMs
        inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, device=dev).float() / cfg.head_dim))
        ang = torch.outer(torch.arange(cfg.max_seq_len, device=dev).float(), inv)
        ang = torch.cat((ang, ang), dim=-1)
        self.register_buffer("rope_cos", ang.cos().half(), persistent=False)
        self.register_buffer("rope_sin", ang.sin().half(), persistent=False)
        self.ref_mod = None   # set to the oracle/_ref module to time the reference kernel inside the same runner

    def _fused_decode_ok(self, x) -> bool:
        """One new token per sequence, CUDA, fused glue on, whole batch present, a batch small enough for it to pay
        off, and a (heads, cache length) the single-kernel decode attention supports; anything else takes the
        rope_kv_update + SDPA path."""
        cfg = self.cfg
        if not (ATTN_DECODE and FUSED_GLUE and x.is_cuda and x.shape[1] == 1 and x.shape[0] == self.batch
                and self.batch <= (ATTN_DECODE_MAX_BATCH if cfg.num_heads // cfg.num_kv_heads < 4 else ATTN_DECODE_MAX_BATCH_GQA)
                and (cfg.max_seq_len <= ATTN_DECODE_MAX_CACHE or self.batch <= ATTN_DECODE_LONG_CACHE_BATCH)):
            return False
        if self._attn_decode_supported is None:
            R = self.tp.world if self.tp is not None else 1      # tensor parallel: this rank's heads only
            self._attn_decode_supported = bool(quick_kernels.attn_decode_supported(cfg.num_heads // R, cfg.num_kv_heads // R,
                                                                                    cfg.head_dim, cfg.max_seq_len))
        return self._attn_decode_supported

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, pos_idx: torch.Tensor, all_logits: bool = False):
        """input_ids (B, T); pos_idx (T,) int64 device tensor of the cache positions being written.  Returns the logits
        of the last position (B, 1, V), or of every position with all_logits (perplexity-style evaluation).
        Multi-token calls (prefill) of up to PREFILL_GRAPH_MAX_ROWS rows are host-bound when run eagerly (≈ 270 launches
        for a 7B model); a (batch, length) shape seen for the second time is captured into a CUDA graph and replayed from
        then on (a few shapes are kept)."""
        if (PREFILL_GRAPH and input_ids.is_cuda and input_ids.dim() == 2 and input_ids.shape[1] > 1 and self.tp is None
                and self.ref_mod is None and input_ids.numel() <= PREFILL_GRAPH_MAX_ROWS and pos_idx.is_cuda
                and not torch.cuda.is_current_stream_capturing()):
            return self._forward_prefill_graphed(input_ids, pos_idx, all_logits)
        return self._forward(input_ids, pos_idx, all_logits)

    def _forward_prefill_graphed(self, input_ids, pos_idx, all_logits):
        # the module-level switches (and a monkey-patched _linear: the tests' dense reference) select different kernels
        key = (tuple(input_ids.shape), bool(all_logits), FUSED_GLUE, FUSED_SILU, NORM_FUSION, id(_linear))
        st = self._prefill_graphs.get(key)
        if st is None:                                   # first sighting: run eagerly, remember the shape
            self._prefill_graphs[key] = {"graph": None}
            while len(self._prefill_graphs) > PREFILL_GRAPH_KEEP:
                self._prefill_graphs.pop(next(iter(self._prefill_graphs)))
            return self._forward(input_ids, pos_idx, all_logits)
        if st["graph"] is None:                          # second sighting: capture
            dev = input_ids.device
            st["ids"], st["pos"] = input_ids.clone(), pos_idx.clone()
            torch.cuda.synchronize(dev)
            side, graph = torch.cuda.Stream(dev), torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                self._forward(st["ids"], st["pos"], all_logits)          # warm-up on the capture stream (idempotent)
                torch.cuda.synchronize(dev)
                with torch.cuda.graph(graph, stream=side):
                    st["out"] = self._forward(st["ids"], st["pos"], all_logits)
            torch.cuda.synchronize(dev)
            st["graph"] = graph
        st["ids"].copy_(input_ids)
        st["pos"].copy_(pos_idx)
        st["graph"].replay()
        return st["out"].clone()                         # the graph's output buffer is reused by the next replay

    def _forward(self, input_ids: torch.Tensor, pos_idx: torch.Tensor, all_logits: bool = False):
        cfg = self.cfg
        x = self.embed(input_ids)
        if self._fused_decode_ok(x):
            cos = sin = attn_mask = None      # qb200_attn_decode indexes the rotary tables and the cache by position itself
        else:
            cos = self.rope_cos.index_select(0, pos_idx)[None, None]
            sin = self.rope_sin.index_select(0, pos_idx)[None, None]
            # causal mask over the static cache: key j visible to query at position p iff j <= p
            keys = torch.arange(cfg.max_seq_len, device=x.device)
            visible = keys[None, :] <= pos_idx[:, None]
            # additive form, built once per forward: a boolean mask makes SDPA convert it in every layer (fill + where)
            attn_mask = torch.zeros(visible.shape, dtype=x.dtype, device=x.device).masked_fill_(~visible, float("-inf")) \
                if x.is_cuda else visible
        src = None        # tensor parallel, peer mode: the gathered buffer the hidden state lives in
        for blk in self.blocks:
            x, src = blk(x, cos, sin, pos_idx, attn_mask, self.ref_mod, (self.rope_cos, self.rope_sin), src)
        if src is not None and not isinstance(src, NormCarry):
            # the final norm is the consumer of the last gathered hidden state: it meets the ranks inside the kernel, so
            # it runs on all rows (no torch op may touch gathered rows before a wait) and the last position is sliced after
            from ... import ops
            x = ops.rmsnorm_tp(x.reshape(-1, x.shape[-1]), self.norm.weight, self.norm.eps, wait=src).view(x.shape)
            return self.lm_head(x if all_logits else x[:, -1:, :])
        return self.lm_head(self.norm(x if all_logits else x[:, -1:, :]))

    # ---- generation on the static cache (what the reference gets from HF generate over its fused blocks,
    # base.py:88-90 + modules/fused/attn.py:187-245: a start_pos that advances with every call)
    def _decode_step(self, tok: torch.Tensor, pos: int, use_graph: bool):
        dev = tok.device
        if self._decode_graph is None:
            self._decode_graph = {"tok": torch.zeros(self.batch, 1, dtype=torch.long, device=dev),
                                  "pos": torch.zeros(1, dtype=torch.long, device=dev), "graph": None, "out": None}
        st = self._decode_graph
        st["tok"].copy_(tok)
        st["pos"].fill_(pos)
        if not (use_graph and dev.type == "cuda" and self.ref_mod is None):
            return self(st["tok"], st["pos"])
        if st["graph"] is None:
            for _ in range(2):                      # idempotent warm-up: same token, same cache slot
                self(st["tok"], st["pos"])
            torch.cuda.synchronize(dev)
            side, graph = torch.cuda.Stream(dev), torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    st["out"] = self(st["tok"], st["pos"])
            torch.cuda.synchronize(dev)
            st["graph"] = graph
        st["graph"].replay()
        return st["out"]

    @staticmethod
    def _pick(logits, do_sample, temperature, top_k, top_p, generator):
        logits = logits[:, -1, :].float()
        if not do_sample:
            return logits.argmax(-1, keepdim=True)
        logits = logits / max(float(temperature), 1e-5)
        if top_k and top_k > 0:
            kth = logits.topk(min(int(top_k), logits.shape[-1]), dim=-1).values[:, -1:]
            logits = logits.masked_fill(logits < kth, float("-inf"))
        if top_p is not None and top_p < 1.0:
            srt, idx = logits.sort(dim=-1, descending=True)
            cum = srt.softmax(-1).cumsum(-1)
            drop = cum - srt.softmax(-1) > top_p          # keep the first token that crosses top_p
            srt = srt.masked_fill(drop, float("-inf"))
            logits = torch.full_like(logits, float("-inf")).scatter(-1, idx, srt)
        tok = torch.multinomial(logits.softmax(-1), 1, generator=generator)
        if tp_world() > 1:          # tensor parallel: every rank must feed the same token back — rank 0's draw wins
            import torch.distributed as dist
            dist.broadcast(tok, src=0)
        return tok

    @torch.no_grad()
    def generate(self, input_ids=None, max_new_tokens: Optional[int] = None, do_sample: bool = False,
                 temperature: float = 1.0, top_k: int = 0, top_p: float = 1.0, eos_token_id=None, pad_token_id=None,
                 attention_mask=None, use_graph: bool = True, generator=None, inputs=None, **unused):
        """Greedy / sampled continuation: prefill of the prompt, then one CUDA-graph replay per token.  Returns
        (B, prompt + new) token ids like HF ``generate``.  The batch must equal the cache batch the model was built
        with (``from_quantized(batch_size=…)``, the reference's AWQ_BATCH_SIZE); prompts must be unpadded."""
        cfg = self.cfg
        ids = input_ids if input_ids is not None else inputs
        if ids is None or ids.dim() != 2:
            raise ValueError("generate needs input_ids of shape (batch, prompt_len)")
        B, T = ids.shape
        if B != self.batch:
            raise ValueError(f"batch {B} != the KV-cache batch {self.batch} this model was built with (batch_size=…)")
        if attention_mask is not None and not bool(torch.as_tensor(attention_mask).bool().all()):
            raise NotImplementedError("padded prompts are not supported by the static-cache runner")
        max_new = int(max_new_tokens) if max_new_tokens is not None else cfg.max_seq_len - T
        if T < 1 or max_new < 1 or T + max_new > cfg.max_seq_len:
            raise ValueError(f"prompt {T} + new {max_new} tokens exceed the cache length {cfg.max_seq_len}")
        dev = self.embed.weight.device
        ids = ids.to(dev)
        eos = None
        if eos_token_id is not None:
            eos = torch.as_tensor(eos_token_id, device=dev).reshape(-1)
            pad = int(pad_token_id) if pad_token_id is not None else int(eos[0])
        logits = self(ids, torch.arange(T, device=dev))
        done = torch.zeros(B, 1, dtype=torch.bool, device=dev)
        out = [ids]
        for i in range(max_new):
            tok = self._pick(logits, do_sample, temperature, top_k, top_p, generator)
            if eos is not None:
                tok = torch.where(done, torch.full_like(tok, pad), tok)
                done = done | (tok[..., None] == eos).any(-1)
            out.append(tok)
            if i + 1 == max_new or (eos is not None and bool(done.all())):
                break
            logits = self._decode_step(tok, T + i, use_graph)
        self.start_pos = T + len(out) - 2          # cache slots written: the prompt and every token fed back
        return torch.cat(out, dim=1)

    @torch.no_grad()
    def forward_stateful(self, input_ids=None, use_cache=True, **kwargs):
        """The calling convention of the reference's fused model (modules/fused/model.py:76-109, attn.py:111-233,
        fused_utils.py:17-29): the cache position is module state — a multi-token call is a new prefill at position 0,
        a single-token call appends at the current position (one CUDA-graph replay).  ``out[0]`` / ``out.logits`` are
        the logits of every input position, as examples/benchmark.py:47-60 expects."""
        from transformers.modeling_outputs import CausalLMOutputWithPast
        dev = self.embed.weight.device
        ids = torch.as_tensor(input_ids, device=dev)
        B, T = ids.shape
        if B != self.batch:
            raise ValueError(f"batch {B} != the KV-cache batch {self.batch} this model was built with (batch_size=…)")
        start = self.start_pos if (T == 1 and use_cache) else 0
        if T == 1 and start + 1 > self.cfg.max_seq_len:
            # the reference's policy when decoding runs past the cache (fused_utils.py:26-28, cache.py:37-50): roll the
            # oldest min(100, cache length) positions out, zero the freed tail, continue at the reduced position
            n = min(100, self.cfg.max_seq_len)
            for blk in self.blocks:
                for c in (blk.cache_k, blk.cache_v):
                    if n < self.cfg.max_seq_len:
                        c.copy_(torch.roll(c, shifts=-n, dims=2))
                    c[:, :, -n:].zero_()
            start -= n
        if start + T > self.cfg.max_seq_len:
            raise ValueError(f"position {start} + {T} tokens exceed the cache length {self.cfg.max_seq_len} (max_new_tokens=…)")
        if T == 1:
            logits = self._decode_step(ids, start, use_graph=True).clone()
        else:
            logits = self(ids, torch.arange(start, start + T, device=dev), all_logits=True)
        self.start_pos = start + T
        return CausalLMOutputWithPast(logits=logits)

    def release_quick_buffers(self, drop=False):
        """Inference-only: every decoder linear keeps its B200-layout copy on the GPU and moves the QUICK-layout
        checkpoint tensors to host memory (WQLinear_QUICK.release_quick_buffers; drop=True discards them:
        random-init benchmark models) — the device then holds each weight once instead of twice.  Not for runs that
        route through the reference kernel (ref_mod)."""
        for blk in self.blocks:
            for m in (blk.qkv_proj, blk.o_proj, blk.gate_up_proj, blk.down_proj):
                m.release_quick_buffers(drop=drop)
        return self

    def weight_bytes(self):
        n = 0
        for blk in self.blocks:
            for m in (blk.qkv_proj, blk.o_proj, blk.gate_up_proj, blk.down_proj):
                n += m.in_features * m.out_features // 2 + (m.in_features // m.group_size) * m.out_features * 4
        return n + self.lm_head.weight.numel() * 2


@torch.no_grad()
def benchmark_generation(model: LlamaLikeQuickModel, n_context: int, n_generate: int, use_graph: bool = True):
    """Reference methodology (examples/benchmark.py:38-67,127-129): prefill tokens/s = ctx*batch / prefill
    seconds, decode tokens/s = batch / median(decode step seconds); CUDA events."""
    dev = model.embed.weight.device
    B = model.batch
    ids = torch.randint(0, model.cfg.vocab_size, (B, n_context), device=dev)
    pos = torch.arange(n_context, device=dev)
    for _ in range(2):
        model(ids, pos)       # warm-up (also builds the B200 weight copies)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); logits = model(ids, pos); e1.record(); torch.cuda.synchronize()
    prefill_s = e0.elapsed_time(e1) * 1e-3

    tok = logits.argmax(-1).view(B, 1)
    step_pos = torch.tensor([n_context], device=dev)
    static_tok, static_pos = tok.clone(), step_pos.clone()
    graph = None
    if use_graph and model.ref_mod is None:   # the reference kernel launches on the legacy stream: not capturable
        for _ in range(2):
            model(static_tok, static_pos)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                static_out = model(static_tok, static_pos)
        torch.cuda.synchronize()
    times = []
    for i in range(n_generate):
        static_pos.fill_(n_context + i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if graph is not None:
            graph.replay(); out = static_out
        else:
            out = model(static_tok, static_pos)
        b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b) * 1e-3)
        static_tok.copy_(out.argmax(-1).view(B, 1))
    times.sort()
    med = times[len(times) // 2]
    return {"batch": B, "prefill_len": n_context, "decode_len": n_generate, "prefill_tokens_per_s": n_context * B / prefill_s,
            "decode_tokens_per_s": B / med, "decode_ms_per_step": med * 1e3, "cuda_graph": graph is not None}


def fuse_hf_model(model, batch_size: int = 1, max_seq_len: Optional[int] = None):
    """Swap a loaded HF Llama/Mistral ``…ForCausalLM`` whose decoder linears are WQLinear_QUICK for the fused runner,
    in place (the job of the reference's LlamaFuser / MistralFuser, models/llama.py:79-126): q‖k‖v and gate‖up are
    concatenated in the QUICK layout (also for grouped-query attention, which the reference's QUICK_cat rejects,
    fused_utils.py:139-142), norms / embedding / lm_head are taken over, ``model.model`` becomes the runner and
    ``model.forward`` / ``model.generate`` route to it.  max_seq_len defaults to config.max_new_tokens like the
    reference (llama.py:115), capped at a sliding window if the family has one."""
    from ..utils.fused_utils import fuse_quick_linears
    hc = model.config
    layers = model.model.layers
    attn0 = layers[0].self_attn
    for name in ("q_proj", "k_proj", "v_proj", "o_proj"):
        if not isinstance(getattr(attn0, name), WQLinear_QUICK):
            raise TypeError(f"self_attn.{name} is {type(getattr(attn0, name)).__name__}, not WQLinear_QUICK — "
                            "fuse_layers needs every decoder linear quantized (modules_to_not_convert must be empty)")
    rope = getattr(hc, "rope_parameters", None) or {}
    if rope.get("rope_type", "default") != "default" or getattr(hc, "rope_scaling", None) not in (None, {}, rope):
        raise NotImplementedError(f"rotary embedding variant {rope or hc.rope_scaling} is not implemented in the fused runner")
    theta = rope.get("rope_theta") or getattr(hc, "rope_theta", None) or 10000.0
    nh, nkv = hc.num_attention_heads, getattr(hc, "num_key_value_heads", None) or hc.num_attention_heads
    if getattr(hc, "head_dim", None) not in (None, hc.hidden_size // nh):
        raise NotImplementedError("head_dim != hidden_size / num_attention_heads")
    seq = int(max_seq_len or getattr(hc, "max_new_tokens", None) or 2048)
    window = getattr(hc, "sliding_window", None) if getattr(hc, "use_sliding_window", True) else None
    if window:
        seq = min(seq, int(window))
    cfg = LlamaLikeConfig(hc.hidden_size, hc.intermediate_size, len(layers), nh, nkv, hc.vocab_size, seq,
                          attn0.q_proj.group_size, float(hc.rms_norm_eps), float(theta))
    dev = attn0.q_proj.qweight.device
    blocks = []
    for layer in layers:
        a, mlp = layer.self_attn, layer.mlp
        blocks.append({"qkv_proj": fuse_quick_linears(a.q_proj, a.k_proj, a.v_proj), "o_proj": a.o_proj,
                       "gate_up_proj": fuse_quick_linears(mlp.gate_proj, mlp.up_proj), "down_proj": mlp.down_proj,
                       "norm_1": layer.input_layernorm.weight, "norm_2": layer.post_attention_layernorm.weight})
        for mod, names in ((a, ("q_proj", "k_proj", "v_proj")), (mlp, ("gate_proj", "up_proj"))):
            for n in names:
                delattr(mod, n)          # the fused copies replace them: free the memory layer by layer
    runner = LlamaLikeQuickModel(cfg, batch_size, dev, parts={"embed": model.model.embed_tokens, "blocks": blocks,
                                                              "norm": model.model.norm.weight, "lm_head": model.lm_head})
    if dev.type == "cuda" and _os.environ.get("QB200_KEEP_QUICK_BUFFERS", "0") != "1":
        runner.release_quick_buffers()      # the device holds every weight once (B200 layout); checkpoint tensors go to the host
    model.model = runner
    model.qb200_fused = True

    model.forward = runner.forward_stateful
    model.generate = runner.generate
    return model
