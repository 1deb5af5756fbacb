"""Timing sweep of the tcgen05 kernel over (tok, split) per M — CUDA-graph replay, cold (rotating
weight sets > L2) and hot.  Development tool; writes gpurun_out/tune.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from quick_b200 import _lib, ops

dev = "cuda"
VARIANTS = [int(v) for v in os.environ.get("VARS", "0").split(",")]   # needs QB200_LIB=libquick_b200_dev.so for 1
SPLITS = [int(v) for v in os.environ.get("SPLITS", "1,2,4,8").split(",")]
K = int(os.environ.get("K", 4096)); N = int(os.environ.get("N", 4096)); G = 128
NSETS = 40
INDEP = os.environ.get("INDEP", "0") == "1"   # QB200_GEMM_INDEPENDENT launches (distinct output buffer per weight set)
Ms = [int(m) for m in os.environ.get("MS", "1,8,16,32,64,128,256,512,1024,2048").split(",")]


def rand_weights(seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    wq = torch.randint(-2**31, 2**31 - 1, (K * N // 8,), device=dev, dtype=torch.int32, generator=g)
    s = (torch.rand(K // G * N, device=dev, generator=g) * 0.01 + 0.002).half().view(torch.int16).to(torch.int32) & 0xFFFF
    z = torch.randint(0, 16, (K // G * N,), device=dev, generator=g, dtype=torch.int32)
    sz = (s | ((0x6400 + z) << 16)).to(torch.int32)
    return wq, sz


sets = [rand_weights(i) for i in range(NSETS)]


def time_graph(fn_i, n_launch, reps=20):
    # warm-up (also sets func attributes outside capture)
    for i in range(min(n_launch, 3)): fn_i(i)
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for i in range(n_launch): fn_i(i)
    g.replay(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        a.record()
        for _ in range(reps): g.replay()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3 / (reps * n_launch))
    return best


res = []
for M in Ms:
    x = torch.randn(M, K, device=dev).half()
    outs = [torch.empty(M, N, device=dev, dtype=torch.float16) for _ in range(NSETS)]
    flop = 2.0 * M * K * N
    alg_bytes = K * N / 2 + (K // G) * N * 2.5 + M * K * 2 + M * N * 2
    toks = [t for t in (16, 32, 64, 128, 256) if t >= min(M, 256) or t == 256]
    toks = [t for t in toks if t <= max(16, 4 * M)] if M <= 64 else [t for t in (64, 128, 256) if t <= max(64, M)]
    auto = ops.plan(M, K, N, G)
    for var, tok in [(v, t) for v in VARIANTS for t in toks]:
        _lib.load().qb200_debug_set_variant(var)
        for split in SPLITS:
            if split == 8 and tok > 32: continue
            try:
                cold = time_graph(lambda i: ops.gemm(x, sets[i % NSETS][0], sets[i % NSETS][1], N, G, tok=tok, split=split, out=outs[i % NSETS], independent=INDEP), NSETS)
                hot = time_graph(lambda i: ops.gemm(x, sets[0][0], sets[0][1], N, G, tok=tok, split=split, out=outs[i % NSETS], independent=INDEP), NSETS)
            except Exception as e:
                print("ERR", M, tok, split, str(e)[:200], flush=True); continue
            rec = {"M": M, "var": var, "tok": tok, "split": split, "cold_us": cold * 1e6, "hot_us": hot * 1e6,
                   "cold_TOPS": flop / cold / 1e12, "cold_GBs": alg_bytes / cold / 1e9, "indep": INDEP, "auto": [tok, split] == list(auto[:2])}
            res.append(rec)
            print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in rec.items()}), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", os.environ.get("OUT", "tune.json")), "w"), indent=1)
