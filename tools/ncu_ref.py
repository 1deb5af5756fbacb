"""One call of the unmodified reference kernel per M (cold weights) for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle.build_ref import load_ref
from quick_b200 import ops
ref = load_ref(); dev = "cuda"; K = N = 4096; G = 128
sets = []
for i in range(16):
    g = torch.Generator(device=dev); g.manual_seed(i)
    q = torch.randint(0, 16, (K, N), device=dev, generator=g, dtype=torch.int32)
    z = torch.randint(0, 16, (K // G, N), device=dev, generator=g, dtype=torch.int32)
    s = (torch.rand(K // G, N, device=dev, generator=g) * 0.01 + 0.002).half()
    sets.append(ops.pack_quick(q, z, s, G))
torch.cuda.synchronize()
i = 0
for M in (1, 8, 16, 64, 128, 256, 512):
    x = torch.randn(M, K, device=dev).half()
    for rep in range(3):
        qw, qz, sc = sets[i % 16]; i += 1
        ref.gemm_forward_cuda_quick(x, qw, sc, qz, 8)
torch.cuda.synchronize(); print("done")
