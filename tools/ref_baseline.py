"""Measure the UNMODIFIED reference kernel (oracle/_ref, built by oracle/build_ref.py) on this GPU:
parity of the oracle definition against it, then a hot/cold timing sweep at K=N=4096 g=128.
Development/measurement tool; writes gpurun_out/ref_baseline.json (summarised into BASELINE.md)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import quick_oracle as qo
from oracle.build_ref import load_ref
from quick_b200 import ops

qk = load_ref()
assert qk is not None, "oracle/_ref/quick_kernels_ref.so missing"
dev = "cuda"
res = {"gpu": torch.cuda.get_device_name(), "torch": torch.__version__}


def make(K, N, G, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    q = torch.randint(0, 16, (K, N), device=dev, generator=g, dtype=torch.int32)
    z = torch.randint(0, 16, (K // G, N), device=dev, generator=g, dtype=torch.int32)
    s = (torch.rand(K // G, N, device=dev, generator=g) * 0.01 + 0.002).half()
    W = ((q - z.repeat_interleave(G, 0)).half() * s.repeat_interleave(G, 0))
    return q, z, s, W


def bench(fn, iters=200, warm=20):
    for i in range(warm): fn(i)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        a.record()
        for i in range(iters): fn(i)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3 / iters)
    return best


res["parity"] = []
for (K, N, G, sk) in [(512, 512, 128, 8), (4096, 4096, 128, 8), (4096, 11008, 128, 2), (11008, 4096, 128, 8), (256, 256, 32, 1), (512, 768, 64, 2)]:
    q, z, s, W = make(K, N, G, 1234)
    qw, qz, sc = ops.pack_quick(q, z, s, G)
    for M in [1, 2, 8, 16, 17, 32, 33, 64, 100, 128, 256]:
        g = torch.Generator(device=dev); g.manual_seed(M)
        A = torch.randn(M, K, device=dev, generator=g).half()
        out = qk.gemm_forward_cuda_quick(A, qw, sc, qz, sk); torch.cuda.synchronize()
        ref64 = A.double() @ W.double()
        rms = ref64.pow(2).mean().sqrt().item()
        e = (out.reshape(M, N).double() - ref64).abs()
        # identity probe: rows of I select rows of dequant(W) -> bit-level check of W16
        res["parity"].append({"K": K, "N": N, "G": G, "split_k": sk, "M": M, "shape": list(out.shape),
                              "max_abs_over_rms": e.max().item() / rms,
                              "allclose": bool(torch.allclose(out.reshape(M, N).double(), ref64, rtol=1e-2, atol=1e-2 * rms))})
    # identity probe (split_k=1 keeps one fp16 rounding): A = I[:64] rows k0..k0+63
    Mi = 64
    for k0 in (0, K - 64):
        A = torch.zeros(Mi, K, device=dev, dtype=torch.float16); A[torch.arange(Mi), k0 + torch.arange(Mi)] = 1
        out = qk.gemm_forward_cuda_quick(A, qw, sc, qz, 1).reshape(Mi, N)
        res["parity"].append({"K": K, "N": N, "G": G, "identity_probe_k0": k0, "w16_bitexact": bool(torch.equal(out, W[k0:k0 + Mi]))})
    print("parity", K, N, G, sk, [r for r in res["parity"] if r["K"] == K and r["N"] == N][-1], flush=True)

K = N = 4096; G = 128
NCOPY = 40
sets = []
for c in range(NCOPY):
    q, z, s, W = make(K, N, G, c); sets.append(ops.pack_quick(q, z, s, G)); del q, z, s, W
Wd = [torch.randn(K, N, device=dev).half() * 0.02 for _ in range(8)]
res["sweep"] = []
for M in [1, 8, 16, 32, 64, 128, 256, 512, 1024, 2048]:
    A = torch.randn(M, K, device=dev).half()
    row = {"M": M, "K": K, "N": N, "G": G}
    flop = 2.0 * M * K * N
    row["alg_bytes"] = K * N / 2 + (K // G) * N * 2.5 + M * K * 2 + M * N * 2
    for sk in [8, 1, 2, 4, 16]:
        t_hot = bench(lambda i: qk.gemm_forward_cuda_quick(A, sets[0][0], sets[0][2], sets[0][1], sk))
        t_cold = bench(lambda i: qk.gemm_forward_cuda_quick(A, sets[i % NCOPY][0], sets[i % NCOPY][2], sets[i % NCOPY][1], sk))
        row[f"ref_sk{sk}_hot_us"] = t_hot * 1e6; row[f"ref_sk{sk}_cold_us"] = t_cold * 1e6
        row[f"ref_sk{sk}_cold_TOPS"] = flop / t_cold / 1e12
    t16 = bench(lambda i: torch.matmul(A, Wd[i % 8]))
    row["fp16_matmul_us"] = t16 * 1e6; row["fp16_matmul_TOPS"] = flop / t16 / 1e12
    res["sweep"].append(row)
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in row.items()}), flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ref_baseline.json"), "w"), indent=1)
print("WROTE ref_baseline.json")
