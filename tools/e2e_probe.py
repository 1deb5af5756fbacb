"""Host-buffer end-to-end probe (development tool): time to ENQUEUE the sweep through qb200_linear_forward_host_async vs time
until every result is back on the host, for NH handles / CONN hardware connections."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", os.environ.get("CONN", "32"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from quick_b200 import ops
sys.argv = ["bench.py"]
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py")); bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
K = N = 4096; MS = [1, 8, 16, 64, 128, 256, 512]; nh = int(os.environ.get("NH", 16))
handles = []
for i in range(nh):
    qw, qz, sc = bench.rand_quick_weights(5000 + i, "cuda")
    handles.append(ops.HostLinear(qw, qz, sc, max_m=512, device=0))
hx = {M: torch.randn(M, K).half().pin_memory() for M in MS}
hy = {M: [torch.empty(M, N, dtype=torch.float16).pin_memory() for _ in range(nh)] for M in MS}
def step():
    t0 = time.perf_counter()
    for M in sorted(MS, reverse=True):
        for i in range(nh):
            handles[i].forward_host_async(hx[M], hy[M][i])
    t1 = time.perf_counter()
    for h in handles: h.synchronize()
    t2 = time.perf_counter()
    return t1 - t0, t2 - t0
step(); step()
r = [step() for _ in range(10)]
iss = sorted(a for a, b in r)[5]; tot = sorted(b for a, b in r)[5]
flops = sum(2.0 * M * K * N for M in MS) * nh
print(f"nh={nh} issue {iss*1e3:.2f} ms  total {tot*1e3:.2f} ms  -> {flops/tot/1e12:.1f} TOPS; bytes each way {sum(2*M*K for M in MS)*nh/1e6:.0f} MB -> {sum(2*M*K for M in MS)*nh/tot/1e9:.1f} GB/s")
