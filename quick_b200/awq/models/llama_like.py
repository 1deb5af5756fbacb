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
# Measured on B200, Llama-2-7B shapes, cache length 192 (profiles/r1e_*): +1.6 % decode tok/s at batch 1, +2.5 % at 8,
# -4 % at 64 (one CTA per (kv head, sequence) re-reads nothing but also shares nothing; cuDNN's kernel wins once there
# are thousands of rows) -> larger batches keep the SDPA path.
ATTN_DECODE_MAX_BATCH = int(_os0.environ.get("QB200_ATTN_DECODE_MAX_BATCH", "16"))
# One CTA streams the whole cache of its (kv head, sequence): measured at a cache of 192 positions only; long caches
# want the positions split over several CTAs (flash-decoding), which SDPA does -> keep it for caches beyond this.
ATTN_DECODE_MAX_CACHE = int(_os0.environ.get("QB200_ATTN_DECODE_MAX_CACHE", "1024"))


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


def shard_quick_linear(m: WQLinear_QUICK, rank: int, world: int) -> WQLinear_QUICK:
    """The column-parallel shard `rank` of `world` of a packed linear: output columns [rank·N/world, (rank+1)·N/world)
    as a WQLinear_QUICK of its own (layout.shard_columns: the inverse of QUICK_cat; N/world must stay a multiple of
    the 128-column tile)."""
    from ...layout import shard_columns
    if m.out_features % (128 * world) != 0:
        raise ValueError(f"N={m.out_features} does not split into {world} shards of 128-column tiles")
    n_local = m.out_features // world
    out = WQLinear_QUICK(m.w_bit, m.group_size, m.in_features, n_local, False, "meta", m.k_split_1, m.k_split_2)
    out.qweight, out.qzeros, out.scales = shard_columns(m.qweight, m.qzeros, m.scales, rank, world)
    out.bias = None if m.bias is None else m.bias[rank * n_local:(rank + 1) * n_local].clone()
    return out


def _gather_columns(local: torch.Tensor) -> torch.Tensor:
    """Tensor-parallel linears (BASELINE config 5): every rank holds N/R output columns of each
    weight and computes its (.., N/R) slab with the same kernel; ONE all-gather (NCCL over NVLink) rebuilds the
    full width.  Same reassembly as quick_b200.parallel.ColumnParallelQuickLinear."""
    import torch.distributed as dist
    R = dist.get_world_size()
    lead, n_local = local.shape[:-1], local.shape[-1]
    flat = local.reshape(-1, n_local).contiguous()
    gathered = torch.empty((R * flat.shape[0], n_local), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(gathered, flat)
    return gathered.view(R, flat.shape[0], n_local).permute(1, 0, 2).reshape(lead + (R * n_local,))


# Tensor-parallel exchange: "peer" = fused GEMM + all-gather over peer memory (quick_b200.parallel.PeerGatherWorkspace),
# "nccl" = kernel + one all_gather_into_tensor + re-layout copy (the baseline).  QB200_TP_MODE selects.
import os as _os
TP_MODE = _os.environ.get("QB200_TP_MODE", "peer")


def _linear(m: WQLinear_QUICK, x, ref_mod=None, residual=None):
    """Route through the B200 kernel, or (baseline runs only) through the unmodified reference kernel.
    residual: returns residual + linear(x) (fused into the GEMM epilogue)."""
    if ref_mod is None:
        ws = getattr(m, "peer_ws", None)
        if ws is not None:
            wq, sz = m._prepacked()
            x2d = x.reshape(-1, x.shape[-1])
            res2d = None if residual is None else residual.reshape(-1, ws.n_total)
            y = ws.gemm(x2d, wq, sz, m.out_features, m.group_size, bias=m.bias, residual=res2d)
            return y.reshape(x.shape[:-1] + (ws.n_total,))
        if getattr(m, "tp_sharded", False):
            y = _gather_columns(m(x))
            return y if residual is None else residual + y
        return m(x, residual if FUSED_GLUE else None) if (residual is None or FUSED_GLUE) else residual + m(x)
    split = m.k_split_1 if m.out_features > m.in_features else m.k_split_2      # quick.py:161-164
    out = ref_mod.gemm_forward_cuda_quick(x.reshape(-1, x.shape[-1]), m.qweight, m.scales, m.qzeros, split)
    out = out.reshape(x.shape[:-1] + (m.out_features,))
    return out if residual is None else residual + out


class Block(nn.Module):
    def __init__(self, cfg: LlamaLikeConfig, dev, gen, batch: int, peer_ws=None, parts=None):
        """parts: {"qkv_proj", "o_proj", "gate_up_proj", "down_proj": WQLinear_QUICK, "norm_1", "norm_2": fp16 weight}
        taken from a loaded checkpoint (fuse_hf_model); without it the weights are random-init."""
        super().__init__()
        self.cfg = cfg
        hd, nh, nkv = cfg.head_dim, cfg.num_heads, cfg.num_kv_heads
        self.norm_1 = RMSNorm(cfg.hidden_size, cfg.rms_eps, dev)
        self.norm_2 = RMSNorm(cfg.hidden_size, cfg.rms_eps, dev)
        if parts is not None:
            R = tp_world()
            self.norm_1.weight.data = parts["norm_1"].detach().to(dev, torch.float16).contiguous()
            self.norm_2.weight.data = parts["norm_2"].detach().to(dev, torch.float16).contiguous()
            expect = {"qkv_proj": (cfg.hidden_size, (nh + 2 * nkv) * hd), "o_proj": (nh * hd, cfg.hidden_size),
                      "gate_up_proj": (cfg.hidden_size, 2 * cfg.intermediate_size),
                      "down_proj": (cfg.intermediate_size, cfg.hidden_size)}
            for name, (k, n) in expect.items():
                m = parts[name]
                if (m.in_features, m.out_features) != (k, n):
                    raise ValueError(f"{name}: ({m.in_features} -> {m.out_features}) does not match the config ({k} -> {n})")
                if R > 1:       # tensor parallel: this rank keeps its N/R output columns, the rest is freed
                    m = shard_quick_linear(m, _tp_rank(), R)
                    if TP_MODE == "peer":
                        m.peer_ws = peer_ws(n)
                m.tp_sharded = R > 1
                setattr(self, name, m)
            self.register_buffer("cache_k", torch.zeros(batch, nkv, cfg.max_seq_len, hd, dtype=torch.float16, device=dev), persistent=False)
            self.register_buffer("cache_v", torch.zeros(batch, nkv, cfg.max_seq_len, hd, dtype=torch.float16, device=dev), persistent=False)
            return
        # Under torch.distributed every linear is column-parallel: this rank's module holds N/R output columns
        # (random-init, so the shard is generated directly instead of slicing a full weight with
        # layout.shard_columns) and _linear() all-gathers the slabs.  N/R must stay a multiple of 128.
        R = tp_world()

        def lin(in_f, out_f):
            assert out_f % (128 * R) == 0, f"N={out_f} does not split into {R} shards of 128-column tiles"
            m = random_quick_linear(in_f, out_f // R, cfg.group_size, dev, gen)
            m.tp_sharded = R > 1
            if R > 1 and TP_MODE == "peer":
                m.peer_ws = peer_ws(out_f)       # one workspace per projection width, shared by all layers
            return m

        self.qkv_proj = lin(cfg.hidden_size, (nh + 2 * nkv) * hd)
        self.o_proj = lin(cfg.hidden_size, cfg.hidden_size)
        self.gate_up_proj = lin(cfg.hidden_size, 2 * cfg.intermediate_size)
        self.down_proj = lin(cfg.intermediate_size, cfg.hidden_size)
        self.register_buffer("cache_k", torch.zeros(batch, nkv, cfg.max_seq_len, hd, dtype=torch.float16, device=dev), persistent=False)
        self.register_buffer("cache_v", torch.zeros(batch, nkv, cfg.max_seq_len, hd, dtype=torch.float16, device=dev), persistent=False)

    def forward(self, x, cos, sin, pos_idx, attn_mask, ref_mod=None, rope=None):
        cfg = self.cfg
        B, T, _ = x.shape
        hd, nh, nkv = cfg.head_dim, cfg.num_heads, cfg.num_kv_heads
        qkv = _linear(self.qkv_proj, self.norm_1(x), ref_mod)
        if T == 1 and attn_mask is None:
            # decode step on the fused path (the model decided: see LlamaLikeQuickModel._fused_decode_ok)
            o = quick_kernels.attn_decode(qkv, rope[0], rope[1], pos_idx, self.cache_k, self.cache_v, nh, nkv)
            x = _linear(self.o_proj, o, ref_mod, residual=x)
            gu = _linear(self.gate_up_proj, self.norm_2(x), ref_mod)
            return _linear(self.down_proj, quick_kernels.silu_mul(gu), ref_mod, residual=x)
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
        x = _linear(self.o_proj, o.transpose(1, 2).reshape(B, T, nh * hd), ref_mod, residual=x)
        gu = _linear(self.gate_up_proj, self.norm_2(x), ref_mod)
        if FUSED_GLUE and gu.is_cuda:
            act = quick_kernels.silu_mul(gu)
        else:
            g, u = gu.split(cfg.intermediate_size, dim=-1)
            act = F.silu(g) * u
        return _linear(self.down_proj, act, ref_mod, residual=x)


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
        self._attn_decode_supported = None
        self.start_pos = 0        # next free cache slot for the stateful HF-style calls (fuse_hf_model) and generate()
        import torch.distributed as dist
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        gen = torch.Generator(device=dev); gen.manual_seed(seed + 1000 * rank)   # every rank draws its own column slabs
        self.embed = parts["embed"] if parts is not None else nn.Embedding(cfg.vocab_size, cfg.hidden_size, device=dev, dtype=torch.float16)
        # tensor parallel, peer mode: one symmetric full-width output buffer per projection width (rows = the largest
        # token count a forward can carry), shared by all layers — consecutive uses are separated by other barriers
        self._peer_ws = {}

        def peer_ws(n_total):
            if n_total not in self._peer_ws:
                from ...parallel import PeerGatherWorkspace
                self._peer_ws[n_total] = PeerGatherWorkspace(batch * cfg.max_seq_len, n_total)
            return self._peer_ws[n_total]

        if parts is None:
            self.blocks = nn.ModuleList([Block(cfg, dev, gen, batch, peer_ws) for _ in range(cfg.num_layers)])
            self.norm = RMSNorm(cfg.hidden_size, cfg.rms_eps, dev)
            self.lm_head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False, device=dev, dtype=torch.float16)
        else:
            self.blocks = nn.ModuleList([Block(cfg, dev, gen, batch, peer_ws, parts=p) for p in parts["blocks"]])
            self.norm = RMSNorm(cfg.hidden_size, cfg.rms_eps, dev)
            self.norm.weight.data = parts["norm"].detach().to(dev, torch.float16).contiguous()
            self.lm_head = parts["lm_head"]
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
                and self.batch <= ATTN_DECODE_MAX_BATCH and cfg.max_seq_len <= ATTN_DECODE_MAX_CACHE
                and tp_world() == 1):     # not yet validated on several GPUs
            return False
        if self._attn_decode_supported is None:
            self._attn_decode_supported = bool(quick_kernels.attn_decode_supported(cfg.num_heads, cfg.num_kv_heads, cfg.head_dim,
                                                                                    cfg.max_seq_len))
        return self._attn_decode_supported

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, pos_idx: torch.Tensor, all_logits: bool = False):
        """input_ids (B, T); pos_idx (T,) int64 device tensor of the cache positions being written.  Returns the logits
        of the last position (B, 1, V), or of every position with all_logits (perplexity-style evaluation)."""
        cfg = self.cfg
        x = self.embed(input_ids)
        if self._fused_decode_ok(x):
            cos = sin = attn_mask = None      # qb200_attn_decode indexes the rotary tables and the cache by position itself
        else:
            cos = self.rope_cos.index_select(0, pos_idx)[None, None]
            sin = self.rope_sin.index_select(0, pos_idx)[None, None]
            # causal mask over the static cache: key j visible to query at position p iff j <= p
            keys = torch.arange(cfg.max_seq_len, device=x.device)
            attn_mask = keys[None, :] <= pos_idx[:, None]
        for blk in self.blocks:
            x = blk(x, cos, sin, pos_idx, attn_mask, self.ref_mod, (self.rope_cos, self.rope_sin))
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
