"""Launch one configuration a few times (for `ncu --set full -k regex:umma`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from quick_b200 import ops
M = int(sys.argv[1]); K = N = 4096; G = 128
INDEP = len(sys.argv) > 2 and sys.argv[2] == "indep"   # QB200_GEMM_INDEPENDENT plan (ncu serialises the launches anyway)
dev = "cuda"
sets = []
for i in range(20):   # 20 x 9 MB > L2: cold weights
    g = torch.Generator(device=dev); g.manual_seed(i)
    wq = torch.randint(-2**31, 2**31 - 1, (K * N // 8,), device=dev, dtype=torch.int32, generator=g)
    s = (torch.rand(K // G * N, device=dev, generator=g) * 0.01 + 0.002).half().view(torch.int16).to(torch.int32) & 0xFFFF
    z = torch.randint(0, 16, (K // G * N,), device=dev, generator=g, dtype=torch.int32)
    sets.append((wq, (s | ((0x6400 + z) << 16)).to(torch.int32)))
x = torch.randn(M, K, device=dev).half()
for i in range(20):
    ops.gemm(x, sets[i][0], sets[i][1], N, G, independent=INDEP)
torch.cuda.synchronize()
print("done", M, ops.plan(M, K, N, G, independent=INDEP))
