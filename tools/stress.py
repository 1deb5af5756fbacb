"""Repeated-launch stress: same config many times, results compared against the first launch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from quick_b200 import ops
tok, split, M, K, N, iters = map(int, sys.argv[1:7]); G = 128
dev = "cuda"
nsets = 8
sets = []
for i in range(nsets):
    g = torch.Generator(device=dev); g.manual_seed(i)
    wq = torch.randint(-2**31, 2**31 - 1, (K * N // 8,), device=dev, dtype=torch.int32, generator=g)
    s = (torch.rand(K // G * N, device=dev, generator=g) * 0.01 + 0.002).half().view(torch.int16).to(torch.int32) & 0xFFFF
    z = torch.randint(0, 16, (K // G * N,), device=dev, generator=g, dtype=torch.int32)
    sets.append((wq, (s | ((0x6400 + z) << 16)).to(torch.int32)))
x = torch.randn(M, K, device=dev).half()
from quick_b200 import _lib
rep = torch.zeros(4096, dtype=torch.int64).pin_memory()
_lib.load().qb200_debug_set_trace(rep.data_ptr())
_lib.load().qb200_debug_set_variant(int(os.environ.get('VAR', '-1')))

def print_report():
    r = rep.tolist()
    blk = r[0] - 1
    print("timeout report from block (%d,%d,%d)" % (blk >> 20, (blk >> 10) & 1023, blk & 1023))
    names = {1: "producer<-cons", 2: "mma<-tfull", 3: "dequant<-full", 4: "dequant<-cons", 5: "epilogue<-accum", 6: "producer tail<-cons", 7: "dequant<-full(prev stage)"}
    for w in range(10):
        row = r[8 + w * 8: 16 + w * 8]
        if row[5]:
            print("  warp %d: waiting %s iter=%d bar=0x%x parity=%d state=0x%016x" % (w, names.get(row[0], row[0]), row[1], row[2] & 0xffff, row[3], row[4] & 0xffffffffffffffff))

def main():
    global bad
    ref = [ops.gemm(x, w, z_, N, G, tok=tok or None, split=split or None).clone() for (w, z_) in sets]
    torch.cuda.synchronize()
    bad = 0
    t0 = time.time()
    for it in range(iters):
        w, z_ = sets[it % nsets]
        out = ops.gemm(x, w, z_, N, G, tok=tok or None, split=split or None)
        if it % int(os.environ.get('SYNC_EVERY', '50')) == int(os.environ.get('SYNC_EVERY', '50')) - 1:
            try:
                torch.cuda.synchronize()
            except Exception as e:
                print("FAIL at iter", it, str(e)[:80])
                r = rep.tolist()
                blk = r[0] - 1
                print("timeout report from block (%d,%d,%d)" % (blk >> 20, (blk >> 10) & 1023, blk & 1023))
                names = {1: "producer<-cons", 2: "mma<-tfull", 3: "dequant<-full", 4: "dequant<-cons", 5: "epilogue<-accum", 6: "producer tail<-cons", 7: "dequant<-full(prev stage)"}
                for w in range(10):
                    row = r[8 + w * 8: 16 + w * 8]
                    if row[5]:
                        print("  warp %d: waiting %s iter=%d bar=0x%x parity=%d state=0x%016x" % (w, names.get(row[0], row[0]), row[1], row[2] & 0xffff, row[3], row[4] & 0xffffffffffffffff))
                sys.exit(2)
            if not torch.equal(out, ref[it % nsets]):
                bad += 1
    print("cfg", tok, split, M, K, N, "iters", iters, "mismatches", bad, "time %.1fs" % (time.time() - t0))

bad = 0
try:
    main()
except Exception as e:
    print("FAIL:", str(e)[:80]); print_report(); sys.exit(2)
