"""Cross-launch timeline (QB200_TRACE build): 40 back-to-back GEMMs replayed from a CUDA graph; CTA (0,0,0) of
every launch stamps %globaltimer at start / after griddepcontrol.wait / accumulator complete / exit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from quick_b200 import ops, _lib
M, K, N = map(int, sys.argv[1:4]); G = 128
tok = int(sys.argv[4]) if len(sys.argv) > 4 else 0
split = int(sys.argv[5]) if len(sys.argv) > 5 else 0
NL = 40
INDEP = os.environ.get("INDEP", "0") == "1"
dev = "cuda"
lib = _lib.load()
sets = []
for i in range(NL):
    wq = torch.randint(-2**31, 2**31 - 1, (K * N // 8,), device=dev, dtype=torch.int32)
    sz = torch.full((K // G * N,), 0x64082000, device=dev, dtype=torch.int32)
    sets.append((wq, sz))
x = torch.randn(M, K, device=dev).half()
outs = [torch.empty(M, N, device=dev, dtype=torch.float16) for _ in range(NL)]
TL = 6 * 256 * 4
TLALL = TL + 8 + 256 * 4
tr = torch.zeros(TLALL + 256 * 8, dtype=torch.int64, device=dev)
def reset():
    tr.zero_()
    tr[TLALL:].view(256, 8)[:, 0::2] = 2 ** 62     # atomicMin slots
def run():
    for i in range(NL): ops.gemm(x, sets[i][0], sets[i][1], N, G, tok=tok or None, split=split or None, out=outs[i], independent=INDEP)
run(); torch.cuda.synchronize()
lib.qb200_debug_set_trace(tr.data_ptr())
st = torch.cuda.Stream(); g = torch.cuda.CUDAGraph()
with torch.cuda.stream(st):
    with torch.cuda.graph(g, stream=st):
        run()
lib.qb200_debug_set_trace(None)
for _ in range(3): g.replay()
torch.cuda.synchronize()
reset(); torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); g.replay(); b.record(); torch.cuda.synchronize()
t = tr.cpu()
n = int(t[TL]); rows = t[TL + 8: TL + 8 + 4 * n].view(n, 4)
t0 = int(rows[0, 0])
print(f"indep={INDEP} M={M} K={K} N={N} plan={ops.plan(M, K, N, G)} forced=({tok},{split}) launches={n} graph_us_per_gemm={a.elapsed_time(b) * 1e3 / NL:.2f}")
print("seq | start  post_wait  accum  exit   (ns, rel. to first start) | start-to-start  exit-to-next-postwait")
for i in range(n):
    r = [int(v) - t0 if int(v) else None for v in rows[i]]
    d = (int(rows[i, 0]) - int(rows[i - 1, 0])) if i else 0
    e = (int(rows[i, 1]) - int(rows[i - 1, 3])) if i and int(rows[i, 1]) and int(rows[i - 1, 3]) else None
    if i < 12 or i >= n - 3: print(i, "|", *r, "|", d, e)
# all-CTA view: rows indexed by the host launch counter (consecutive for the 40 launches of the graph)
allr = t[TLALL:].view(256, 8)
used = [i for i in range(256) if int(allr[i, 1]) > 0]
used.sort(key=lambda i: int(allr[i, 0]))
print("launch | CTA start min..max | wait-returned min..max | exit min..max  (ns rel.) | grid done -> next wait-returned(min)")
prev_exit = None
for j, i in enumerate(used):
    r = [int(v) - t0 for v in allr[i, :6]]
    gap = (int(allr[i, 2]) - prev_exit) if prev_exit is not None else None
    if j < 12 or j >= len(used) - 2: print(j, "|", r[0], r[1], "|", r[2], r[3], "|", r[4], r[5], "|", gap)
    prev_exit = int(allr[i, 5])
