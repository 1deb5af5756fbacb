"""Timing of the decode-sized GEMV path (tok = 1) against the tcgen05 kernel's ordered plan, same protocol as bench.py:
CUDA-graph replay of NSETS GEMMs over rotating weight sets (cold weights), CUDA events, best of 5 replays.
usage: python tools/gemv_check.py [--out file.json]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from quick_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--sets", type=int, default=40)
args = ap.parse_args()
dev = torch.device("cuda")
HBM = 6458.4
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
G = 128


def weights(K, N, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    wq = torch.randint(-2 ** 31, 2 ** 31 - 1, (K * N // 8,), device=dev, dtype=torch.int32, generator=g)
    s = (torch.rand(K // G * N, device=dev, generator=g) * 0.01 + 0.002).half().view(torch.int16).to(torch.int32) & 0xFFFF
    z = torch.randint(0, 16, (K // G * N,), device=dev, generator=g, dtype=torch.int32)
    return wq, (s | ((0x6400 + z) << 16)).to(torch.int32)


def timed(fn, n):
    fn(); torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / n)
    return best


rows = []
for (K, N) in ((4096, 4096), (4096, 12288), (4096, 22016), (11008, 4096), (8192, 8192)):
    nsets = max(8, min(args.sets, int(3.0e9 // (K * N // 2))))
    sets = [weights(K, N, i) for i in range(nsets)]
    for M in (1, 2, 4):
        x = torch.randn(M, K, device=dev).half()
        outs = [torch.empty(M, N, device=dev, dtype=torch.float16) for _ in range(nsets)]
        res = {}
        for name, kw in (("tcgen05", {}), ("gemv", {"tok": 1, "split": 1})):
            def fn():
                for i in range(nsets):
                    ops.gemm(x, sets[i][0], sets[i][1], N, G, out=outs[i], **kw)
            res[name] = timed(fn, nsets)
            res[name + "_out"] = outs[0].clone()
        rms = res["tcgen05_out"].float().pow(2).mean().sqrt().item()
        err = (res["gemv_out"].float() - res["tcgen05_out"].float()).abs().max().item()
        nbytes = K * N // 2 + (K // G) * N * 2 + (K // G) * N // 2 + 2 * M * K + 2 * M * N
        row = {"K": K, "N": N, "M": M, "sets": nsets, "tcgen05_us": round(res["tcgen05"], 3), "gemv_us": round(res["gemv"], 3),
               "gemv_hbm_frac": round(nbytes / (res["gemv"] * 1e-6) / 1e9 / HBM, 4),
               "tcgen05_hbm_frac": round(nbytes / (res["tcgen05"] * 1e-6) / 1e9 / HBM, 4), "max_abs_diff_over_rms": round(err / rms, 5)}
        rows.append(row)
        print(json.dumps(row), flush=True)
    del sets
    torch.cuda.empty_cache()
if args.out:
    json.dump(rows, open(args.out, "w"), indent=1)
