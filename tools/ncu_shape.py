import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from quick_b200 import ops
M, K, N = map(int, sys.argv[1:4]); G = 128
dev = "cuda"
sets = []
for i in range(6):
    g = torch.Generator(device=dev); g.manual_seed(i)
    wq = torch.randint(-2**31, 2**31 - 1, (K * N // 8,), device=dev, dtype=torch.int32, generator=g)
    sets.append((wq, torch.full((K // G * N,), 0x64081c00, device=dev, dtype=torch.int32)))
x = torch.randn(M, K, device=dev).half()
for i in range(12):
    ops.gemm(x, sets[i % 6][0], sets[i % 6][1], N, G)
torch.cuda.synchronize()
print("done", ops.plan(M, K, N, G))
