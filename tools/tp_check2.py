"""torchrun --nproc-per-node R tools/tp_check2.py — tensor parallelism with the hand-over inside the kernels
(qb200_gemm_w4a16_tp / qb200_rmsnorm_tp / qb200_silu_mul_tp / qb200_scatter_cols over quick_b200.parallel.GatheredBuffer)
against the same dataflow with NCCL all-gathers: op-level bit-identity, model-level logits (eager, graph replay,
repeated), then tokens/s of both modes.  Development / evidence tool; prints one JSON line per check on rank 0."""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)

from quick_b200 import ops
from quick_b200.parallel import GatheredBuffer
from quick_b200.awq.models import llama_like as ll


def say(**kw):
    ok = torch.tensor([1 if kw.get("ok", True) else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    kw["ok_all_ranks"] = bool(ok.item())
    if rank == 0:
        print(json.dumps(kw), flush=True)
    return kw["ok_all_ranks"]


def rand_b200(K, N, G, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    wq = torch.randint(-2 ** 31, 2 ** 31 - 1, (K * N // 8,), device=dev, dtype=torch.int32, generator=g)
    s = (torch.rand(K // G * N, device=dev, generator=g) * 0.01 + 0.002).half().view(torch.int16).to(torch.int32) & 0xFFFF
    z = torch.randint(0, 16, (K // G * N,), device=dev, generator=g, dtype=torch.int32)
    return wq, (s | ((0x6400 + z) << 16)).to(torch.int32)


def gather_cols(local_t):
    flat = local_t.contiguous()
    out = torch.empty((world * flat.shape[0], flat.shape[1]), dtype=flat.dtype, device=dev)
    dist.all_gather_into_tensor(out, flat)
    return out.view(world, flat.shape[0], flat.shape[1]).permute(1, 0, 2).reshape(flat.shape[0], -1)


all_ok = True
# ---- 1. op level: column-parallel GEMM -> gathered buffer -> consumers that wait inside the kernel ----
K, G = 4096, 128
N_l = 1024
wq, sz = rand_b200(K, N_l, G, 100 + rank)
wq2, sz2 = rand_b200(N_l * world, 512, G, 200 + rank)
bufA = GatheredBuffer(512, N_l * world)
bufB = GatheredBuffer(512, 512 * world)
norm_w = torch.ones(N_l * world, device=dev, dtype=torch.float16)
for M in (1, 16, 64, 300):
    x = torch.randn(M, K, device=dev, generator=torch.Generator(device=dev).manual_seed(M)).half()
    res = torch.randn(M, N_l * world, device=dev, generator=torch.Generator(device=dev).manual_seed(1000 + M)).half()
    for it in range(3):     # repeated fills of the same buffers (alternating A, B keeps the reuse rule)
        ops.gemm_tp(x, wq, sz, N_l, G, residual=res, dst=bufA, col0=rank * N_l)
        y = ops.rmsnorm_tp(bufA.rows(M), norm_w, 1e-5, wait=bufA)                       # consumer 1: RMSNorm rows
        ops.gemm_tp(bufA.rows(M), wq2, sz2, 512, G, dst=bufB, col0=rank * 512, wait=bufA)   # consumer 2: GEMM activations
        z2 = ops.rmsnorm_tp(bufB.rows(M), torch.ones(512 * world, device=dev, dtype=torch.float16), 1e-5, wait=bufB)
    torch.cuda.synchronize()
    want_full = res + gather_cols(ops.gemm(x, wq, sz, N_l, G))
    want_y = ops.rmsnorm_tp(want_full.contiguous(), norm_w, 1e-5)
    want_b = gather_cols(ops.gemm(want_full.contiguous(), wq2, sz2, 512, G))
    ok = torch.equal(bufA.rows(M), want_full) and torch.equal(y, want_y) and torch.equal(bufB.rows(M), want_b)
    all_ok &= say(check="gemm_tp -> rmsnorm_tp / gemm_tp(wait)", M=M, ok=bool(ok), multicast=bufA.multicast_ptr is not None)
# scatter_cols and silu_mul_tp
for M in (1, 33, 256):
    src = torch.randn(M, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(7 * M + rank)).half()
    gu = torch.randn(M, 2 * 256, device=dev, generator=torch.Generator(device=dev).manual_seed(9 * M + rank)).half()
    bufC = bufB if world * 256 == bufB.width else None
    if bufC is None:
        bufC = GatheredBuffer(512, 256 * world)
    ops.scatter_cols(src, bufC, rank * 256)
    got1 = ops.rmsnorm_tp(bufC.rows(M), torch.ones(256 * world, device=dev, dtype=torch.float16), 1e-5, wait=bufC).clone()
    ops.gemm_tp(torch.zeros(M, K, device=dev, dtype=torch.float16), wq, sz, N_l, G, dst=bufA, col0=rank * N_l)   # another buffer's fill in between
    ops.rmsnorm_tp(bufA.rows(M), norm_w, 1e-5, wait=bufA)
    ops.silu_mul_tp(gu, bufC, rank * 256)
    got2 = ops.rmsnorm_tp(bufC.rows(M), torch.ones(256 * world, device=dev, dtype=torch.float16), 1e-5, wait=bufC)
    torch.cuda.synchronize()
    import quick_kernels
    w1 = ops.rmsnorm_tp(gather_cols(src).contiguous(), torch.ones(256 * world, device=dev, dtype=torch.float16), 1e-5)
    w2 = ops.rmsnorm_tp(gather_cols(quick_kernels.silu_mul(gu)).contiguous(), torch.ones(256 * world, device=dev, dtype=torch.float16), 1e-5)
    all_ok &= say(check="scatter_cols / silu_mul_tp", M=M, ok=bool(torch.equal(got1, w1) and torch.equal(got2, w2)))
    ops.gemm_tp(torch.zeros(M, K, device=dev, dtype=torch.float16), wq, sz, N_l, G, dst=bufA, col0=rank * N_l)
    ops.rmsnorm_tp(bufA.rows(M), norm_w, 1e-5, wait=bufA)
del bufA, bufB

# ---- 1b. what one hand-over costs: chain of [column-parallel GEMM -> RMSNorm of the gathered rows], graph replay ----
def time_graph(fn, n):
    fn(); torch.cuda.synchronize()
    st = torch.cuda.Stream(); g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            fn()
    g.replay(); torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        g.replay()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 10 / n * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(t.item(), 3)

H = 4096
bufs = [GatheredBuffer(64, H), GatheredBuffer(64, H)]
nw = torch.ones(H, device=dev, dtype=torch.float16)
wsets = [rand_b200(H, H // world, G, 300 + i) for i in range(8)]
wfull = [rand_b200(H, H, G, 400 + i) for i in range(8)]
for M in (1, 64):
    x0 = torch.randn(M, H, device=dev).half()

    def chain_tp():
        x = x0
        for i in range(16):
            ops.gemm_tp(x, *wsets[i % 8], H // world, G, dst=bufs[i & 1], col0=rank * (H // world))
            x = ops.rmsnorm_tp(bufs[i & 1].rows(M), nw, 1e-5, wait=bufs[i & 1])

    def chain_local():
        x = x0
        for i in range(16):
            y = ops.gemm(x, *wfull[i % 8], H, G)
            x = ops.rmsnorm_tp(y, nw, 1e-5)

    def chain_local_shard():     # the same per-rank GEMM work as the TP chain, no hand-over (not a valid computation)
        x = x0
        for i in range(16):
            y = ops.gemm(x, *wsets[i % 8], H // world, G)
            x = ops.rmsnorm_tp(x, nw, 1e-5)
    say(check="hand-over cost", M=M, us_per_pair_tp=time_graph(chain_tp, 16), us_per_pair_one_gpu=time_graph(chain_local, 16),
        us_per_pair_shard_no_handover=time_graph(chain_local_shard, 16))
# ---- 1c. stress: thousands of hand-overs at the sizes where a late slab would show (tiny kernels), every chain verified ----
for M in (1, 8):
    bad = 0
    for it in range(150):
        x0 = torch.randn(M, H, device=dev, generator=torch.Generator(device=dev).manual_seed(1000 * M + it)).half()
        x = x0
        for i in range(16):       # GEMM fills a gathered buffer, the RMSNorm of its rows is the waiting reader (values stay bounded)
            ops.gemm_tp(x, *wsets[(i + it) % 8], H // world, G, dst=bufs[i & 1], col0=rank * (H // world))
            x = ops.rmsnorm_tp(bufs[i & 1].rows(M), nw, 1e-5, wait=bufs[i & 1])
        got = x.clone()
        x = x0
        for i in range(16):
            x = ops.rmsnorm_tp(gather_cols(ops.gemm(x, *wsets[(i + it) % 8], H // world, G)).contiguous(), nw, 1e-5)
        bad += int(not torch.equal(got, x))
    all_ok &= say(check="stress: 150 chains x 16 hand-overs, each verified against the NCCL chain", M=M, mismatching_chains=bad, ok=bad == 0)
del bufs

if os.environ.get('TP_ONLY_HANDOVER') == '1':
    dist.destroy_process_group(); sys.exit(0)
# ---- 2. model level: peer mode vs NCCL mode, same weights ----
cfg = ll.LlamaLikeConfig(1024, 2816, 4, 8, 8 if world > 4 else 4 if world > 2 else 2, vocab_size=2048, max_seq_len=96)
if cfg.num_kv_heads % world:
    cfg.num_kv_heads = world
B = 4
models = {}
for mode in ("nccl", "peer"):
    ll.TP_MODE = mode
    torch.manual_seed(0)
    models[mode] = ll.LlamaLikeQuickModel(cfg, B, dev, seed=3)
ids = torch.randint(0, cfg.vocab_size, (B, 24), device=dev, generator=torch.Generator(device=dev).manual_seed(11))
dist.broadcast(ids, 0)
pos = torch.arange(24, device=dev)
a, b = models["nccl"](ids, pos, all_logits=True), models["peer"](ids, pos, all_logits=True)
torch.cuda.synchronize()
rms = a.float().pow(2).mean().sqrt().item()
all_ok &= say(check="prefill logits peer vs nccl", bit_identical=bool(torch.equal(a, b)), max_err_over_rms=float((a.float() - b.float()).abs().max().item() / rms),
              ok=bool((a.float() - b.float()).abs().max().item() <= 2e-2 * rms))
sa, sb = models["nccl"].generate(ids, max_new_tokens=24), models["peer"].generate(ids, max_new_tokens=24)   # CUDA-graph decode, fused attention
all_ok &= say(check="generate() tokens peer vs nccl (graph decode)", ok=bool(torch.equal(sa, sb)), tokens=int(sa.shape[1]))
sc = models["peer"].generate(ids, max_new_tokens=24)
all_ok &= say(check="second generate on the same graphs", ok=bool(torch.equal(sb, sc)))
del models
torch.cuda.empty_cache()

# ---- 3. tokens/s: 7B shapes, both modes ----
rows = []
for name, batches in (("llama-2-7b", (1, 64)),) + ((("llama-2-70b", (1, 8)),) if os.environ.get("TP_70B", "1") == "1" else ()):
    for mode in ("peer", "nccl"):
        ll.TP_MODE = mode
        for bs in batches:
            try:
                c = copy.deepcopy(ll.PRESETS[name]); c.max_seq_len = 256
                m = ll.LlamaLikeQuickModel(c, bs, dev)
                m.release_quick_buffers(drop=True)
                r = ll.benchmark_generation(m, 128, 128)
                rows.append({"model": name, "mode": mode, "batch": bs, "decode_tok_s": round(r["decode_tokens_per_s"], 1),
                             "prefill_tok_s": round(r["prefill_tokens_per_s"], 1), "decode_ms": round(r["decode_ms_per_step"], 3)})
                del m
            except Exception as e:
                rows.append({"model": name, "mode": mode, "batch": bs, "error": f"{type(e).__name__}: {e}"[:200]})
            torch.cuda.empty_cache()
            if rank == 0:
                print(json.dumps(rows[-1]), flush=True)
say(check="ALL", ok=bool(all_ok))
dist.destroy_process_group()
