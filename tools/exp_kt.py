"""Experiment: per-stage / fixed cost vs TMEM ring depth (QB200_KT env) and K."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from quick_b200 import ops
dev = "cuda"; G = 128; NSETS = 24

def rand_weights(K, N, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    wq = torch.randint(-2**31, 2**31 - 1, (K * N // 8,), device=dev, dtype=torch.int32, generator=g)
    s = (torch.rand(K // G * N, device=dev, generator=g) * 0.01 + 0.002).half().view(torch.int16).to(torch.int32) & 0xFFFF
    z = torch.randint(0, 16, (K // G * N,), device=dev, generator=g, dtype=torch.int32)
    return wq, (s | ((0x6400 + z) << 16)).to(torch.int32)

def time_graph(fn_i, n_launch, reps=20):
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

kt = os.environ.get("QB200_KT", "default")
tiny = torch.zeros(32, device=dev)
print(json.dumps({"kt": kt, "empty_node_us": round(time_graph(lambda i: tiny.add_(1.0), 40) * 1e6, 2)}), flush=True)
cases = [(16, 1, 1, 256, 4096), (16, 1, 1, 1024, 4096), (16, 1, 1, 4096, 4096), (16, 4, 1, 1024, 4096), (16, 4, 1, 4096, 4096),
         (256, 1, 256, 1024, 4096), (256, 1, 256, 4096, 4096), (256, 1, 1024, 4096, 4096)]
for (tok, split, M, K, N) in cases:
    sets = [rand_weights(K, N, i) for i in range(NSETS)]
    x = torch.randn(M, K, device=dev).half(); out = torch.empty(M, N, device=dev, dtype=torch.float16)
    t = time_graph(lambda i: ops.gemm(x, sets[i % NSETS][0], sets[i % NSETS][1], N, G, tok=tok, split=split, out=out), NSETS)
    print(json.dumps({"kt": kt, "tok": tok, "split": split, "M": M, "K": K, "stages_per_cta": K // 64 // split, "us": round(t * 1e6, 2)}), flush=True)
    del sets
