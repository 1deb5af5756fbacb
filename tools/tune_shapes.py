"""Per-shape (K,N) split sweep at small M for the 7B/70B layer shapes (planner tuning)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from quick_b200 import ops
dev = "cuda"; G = 128
def rand_weights(K, N, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    wq = torch.randint(-2**31, 2**31 - 1, (K * N // 8,), device=dev, dtype=torch.int32, generator=g)
    sz = torch.full((K // G * N,), 0x64081c00, device=dev, dtype=torch.int32)
    return wq, sz
def time_graph(fn_i, n_launch, reps=10):
    for i in range(3): fn_i(i)
    torch.cuda.synchronize()
    st = torch.cuda.Stream(); g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for i in range(n_launch): fn_i(i)
    g.replay(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); best = 1e9
    for _ in range(3):
        a.record()
        for _ in range(reps): g.replay()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3 / (reps * n_launch))
    return best
shapes = [(4096, 12288), (4096, 22016), (11008, 4096), (4096, 4096), (8192, 1280), (8192, 7168), (28672, 1024)]
if os.environ.get("SHAPES") == "mistral":     # Mistral-7B: q|k|v (GQA), gate|up, down
    shapes = [(4096, 6144), (4096, 28672), (14336, 4096)]
if os.environ.get("SHAPES") == "70b":     # Llama-2-70B on one GPU and its 2 / 4-GPU shards (q|k|v, o, gate|up, down)
    shapes = [(8192, 10240), (8192, 8192), (8192, 57344), (28672, 8192), (8192, 5120), (8192, 4096), (8192, 28672), (28672, 4096),
              (8192, 2560), (8192, 2048), (8192, 14336), (28672, 2048)]
for (K, N) in shapes:
    nsets = max(4, int(400e6 // (K * N // 2)))
    sets = [rand_weights(K, N, i) for i in range(nsets)]
    for M in (1, 16, 64):
        x = torch.randn(M, K, device=dev).half(); out = torch.empty(M, N, device=dev, dtype=torch.float16)
        tok = 16 if M <= 16 else 64
        auto = ops.plan(M, K, N, G)
        res = {}
        for split in (1, 2, 4, 8):
            if split == 8 and tok > 32: continue
            try:
                t = time_graph(lambda i: ops.gemm(x, sets[i % nsets][0], sets[i % nsets][1], N, G, tok=tok, split=split, out=out), nsets)
                res[split] = round(t * 1e6, 2)
            except Exception as e:
                res[split] = "ERR"
        byts = K * N / 2 + K // G * N * 2.5
        print(json.dumps({"K": K, "N": N, "M": M, "tiles": N // 128, "auto": auto, "us_by_split": res, "hbm_roof_us": round(byts / 6538e3, 2)}), flush=True)
    del sets
