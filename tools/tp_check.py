"""torchrun --nproc-per-node R tools/tp_check.py — fused GEMM + all-gather over peer memory against the NCCL
all-gather path (quick_b200.parallel) on identical shards, eager and under CUDA-graph replay."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
os.environ.setdefault("NCCL_DEBUG", "NONE")
local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
from quick_b200 import ops, layout
from quick_b200.parallel import ColumnParallelQuickLinear, PeerGatherWorkspace
ok = True
keep = []     # symmetric-memory objects must not be freed while a stream is capturing
for (K, N, G, Ms) in ((1024, 1024, 128, (1, 16, 100)), (4096, 4096, 128, (1, 64, 256, 300))):
    g = torch.Generator(device="cuda"); g.manual_seed(7)          # same full weight on every rank
    q = torch.randint(0, 16, (K, N), device="cuda", generator=g, dtype=torch.int32)
    z = torch.randint(0, 16, (K // G, N), device="cuda", generator=g, dtype=torch.int32)
    s = (torch.rand(K // G, N, device="cuda", generator=g) * 0.01 + 0.002).half()
    qw, qz, sc = ops.pack_quick(q, z, s, G)
    bias = torch.randn(N, device="cuda", generator=g).half()
    lin = ColumnParallelQuickLinear(qw, qz, sc, bias)
    sh = lin.shard
    wq, sz, *_ = ops.prepack(sh.qweight, sh.qzeros, sh.scales)
    ws = PeerGatherWorkspace(max(Ms), N)
    keep += [lin, ws]
    if rank == 0: print("multicast:", ws.multicast_ptr is not None, flush=True)
    for M in Ms:
        x = torch.randn(M, K, device="cuda", generator=g).half()
        res = torch.randn(M, N, device="cuda", generator=g).half()
        want = lin(x)                                   # kernel + NCCL all-gather + re-layout
        got = ws.gemm(x, wq, sz, sh.n_local, G, bias=sh.bias).clone()
        ws.release()                                    # back-to-back reuse of one workspace: everybody is done reading
        got_res = ws.gemm(x, wq, sz, sh.n_local, G, bias=sh.bias, residual=res).clone()
        ws.release()
        torch.cuda.synchronize()
        e1, e2 = torch.equal(got, want), torch.equal(got_res, res + want)
        ok = ok and e1 and e2
        if rank == 0: print(f"K={K} N={N} M={M}: peer==nccl {e1}, residual {e2}", flush=True)
    # CUDA-graph replay: the device-side epoch must advance on every replay
    x = torch.randn(16, K, device="cuda", generator=g).half()
    want = lin(x)
    st = torch.cuda.Stream(); gr = torch.cuda.CUDAGraph()
    ws.gemm(x, wq, sz, sh.n_local, G, bias=sh.bias); torch.cuda.synchronize()
    with torch.cuda.stream(st):
        with torch.cuda.graph(gr, stream=st):
            out = ws.gemm(x, wq, sz, sh.n_local, G, bias=sh.bias)
    for i in range(200):
        ws.buf.zero_()
        gr.replay()
        ws.release()
        if i % 50 == 49:
            torch.cuda.synchronize()
            e = torch.equal(out, want); ok = ok and e
            if rank == 0: print(f"  graph replay {i + 1}: {e}", flush=True)
    dist.barrier()
t = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("TP_CHECK", "PASS" if t.item() == 1 else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1 else 1)
