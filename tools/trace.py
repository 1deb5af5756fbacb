"""Dump the in-kernel clock64 trace of CTA (0,0,0) for one launch (debug tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from quick_b200 import ops, _lib
tok, split, M, K, N = map(int, sys.argv[1:6]); G = 128
var = int(sys.argv[6]) if len(sys.argv) > 6 else 0
_lib.load().qb200_debug_set_variant(var)
dev = "cuda"
wq = torch.randint(-2**31, 2**31 - 1, (K * N // 8,), device=dev, dtype=torch.int32)
sz = torch.full((K // G * N,), 0x64082000, device=dev, dtype=torch.int32)
x = torch.randn(M, K, device=dev).half()
for _ in range(3): ops.gemm(x, wq, sz, N, G, tok=tok, split=split)
torch.cuda.synchronize()
tr = torch.zeros(6 * 256 * 4 + 8 + 256 * 4 + 256 * 8, dtype=torch.int64, device=dev)
lib = _lib.load(); lib.qb200_debug_set_trace(tr.data_ptr())
ops.gemm(x, wq, sz, N, G, tok=tok, split=split); torch.cuda.synchronize()
lib.qb200_debug_set_trace(None)
t = tr.cpu()[:6 * 256 * 4].view(6, 256, 4)
t0 = int(t[3, 0, 0])
rel = lambda v: int(v) - t0 if int(v) else None
nkb = (K // 64 // split + 1) // 2
print("cfg", tok, split, M, K, N, "var", var, "stages", nkb, "w_prefetch_issued", rel(t[0, 0, 2]))
print("setup_done", rel(t[3, 0, 1]), "accum_seen", rel(t[3, 0, 2]), "cluster_bar1", rel(t[3, 2, 0]), "scatter_done", rel(t[3, 2, 1]), "cluster_bar2", rel(t[3, 2, 2]), "tile_staged", rel(t[3, 2, 3]), "epi_done", rel(t[3, 0, 3]), "dealloc", rel(t[3, 1, 0]))
print("it | prod: x_slot_free issued | mma: loop_top tfull_seen ready mmas_issued committed | deq: w_landed lds+consts tmem_free st_half st_issued st_done handed_off")
for it in range(min(nkb, 40)):
    print(it, "|", rel(t[0, it, 0]), rel(t[0, it, 1]), "|", rel(t[5, it, 0]), rel(t[5, it, 1]), rel(t[1, it, 0]), rel(t[1, it, 1]), rel(t[1, it, 2]), "|",
          rel(t[2, it, 0]), rel(t[4, it, 0]), rel(t[4, it, 1]), rel(t[4, it, 2]), rel(t[2, it, 1]), rel(t[4, it, 3]), rel(t[2, it, 2]), "| handoff q0 q1 q3:", rel(t[5, it, 1]), rel(t[5, it, 2]), rel(t[5, it, 3]))
