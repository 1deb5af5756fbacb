"""In-kernel clock64 trace of CTA (0,0,0) of the LAST launch of a PDL chain (steady state of back-to-back GEMMs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from quick_b200 import ops, _lib
tok, split, M, K, N = map(int, sys.argv[1:6]); G = 128
INDEP = os.environ.get("INDEP", "0") == "1"
dev = "cuda"; NL = 12
lib = _lib.load()
lib.qb200_debug_set_variant(int(os.environ.get("VAR", "-1")))
sets = [(torch.randint(-2**31, 2**31 - 1, (K * N // 8,), device=dev, dtype=torch.int32),
         torch.full((K // G * N,), 0x64082000, device=dev, dtype=torch.int32)) for _ in range(NL)]
x = torch.randn(M, K, device=dev).half()
outs = [torch.empty(M, N, device=dev, dtype=torch.float16) for _ in range(NL)]
def run():
    for i in range(NL): ops.gemm(x, sets[i][0], sets[i][1], N, G, tok=tok or None, split=split or None, out=outs[i], independent=INDEP)
run(); torch.cuda.synchronize()
tr = torch.zeros(6 * 256 * 4 + 8 + 256 * 4 + 256 * 8, dtype=torch.int64, device=dev)
lib.qb200_debug_set_trace(tr.data_ptr())
st = torch.cuda.Stream(); g = torch.cuda.CUDAGraph()
with torch.cuda.stream(st):
    with torch.cuda.graph(g, stream=st):
        run()
lib.qb200_debug_set_trace(None)
for _ in range(3): g.replay()
torch.cuda.synchronize(); tr.zero_(); torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
t = tr.cpu()[:6 * 256 * 4].view(6, 256, 4)
t0 = int(t[3, 0, 0])
rel = lambda v: int(v) - t0 if int(v) else None
nkb = (K // 64 // (split or 1) + 1) // 2
print("CHAIN cfg", tok, split, M, K, N, "indep", INDEP, "stages", nkb, "w_prefetch_issued", rel(t[0, 0, 2]))
print("setup_done", rel(t[3, 0, 1]), "accum_seen", rel(t[3, 0, 2]), "cluster_bar1", rel(t[3, 2, 0]), "scatter_done", rel(t[3, 2, 1]), "cluster_bar2", rel(t[3, 2, 2]), "tile_staged", rel(t[3, 2, 3]), "epi_done", rel(t[3, 0, 3]), "dealloc", rel(t[3, 1, 0]))
print("it | prod: x_slot_free issued | mma: loop_top ready mmas_issued committed | deq: w_landed lds+consts tmem_free st_half st_issued st_done handed_off | handoff other quads")
for it in range(min(nkb, 40)):
    print(it, "|", rel(t[0, it, 0]), rel(t[0, it, 1]), "|", rel(t[5, it, 0]), rel(t[1, it, 0]), rel(t[1, it, 1]), rel(t[1, it, 2]), "|",
          rel(t[2, it, 0]), rel(t[4, it, 0]), rel(t[4, it, 1]), rel(t[4, it, 2]), rel(t[2, it, 1]), rel(t[4, it, 3]), rel(t[2, it, 2]), "|", rel(t[5, it, 1]), rel(t[5, it, 2]), rel(t[5, it, 3]))
