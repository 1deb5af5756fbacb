"""In-process correctness sweep of the tcgen05 kernel over (tok, split, M, K, N, G) and tile variants
against a torch fp32 matmul of the oracle-defined W16 (development tool; the tests are the gate)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import quick_oracle as qo
from quick_b200 import _lib, ops

variants = [int(v) for v in os.environ.get("VARS", "-1").split(",")]
lib = _lib.load()
cases = {}


def case(K, N, G):
    if (K, N, G) not in cases:
        q, z, s = qo.make_case(K, N, G, 1234)
        tq, tz, ts = torch.from_numpy(q).cuda(), torch.from_numpy(z).cuda(), torch.from_numpy(s).cuda()
        qw, qz, sc = ops.pack_quick(tq, tz, ts, G)
        wq, sz, *_ = ops.prepack(qw, qz, sc)
        W = ((tq - tz.repeat_interleave(G, 0)).half() * ts.repeat_interleave(G, 0)).float()
        cases[(K, N, G)] = (wq, sz, W)
    return cases[(K, N, G)]


cfgs = []
cfgs += [(16, 1, 1, 64, 128, 64), (16, 1, 16, 128, 128, 128), (16, 1, 16, 512, 512, 128)]
cfgs += [(t, 1, t, 512, 512, 128) for t in (32, 64, 128, 256)]
cfgs += [(16, s, m, 2048, 256, 128) for s in (2, 4, 8) for m in (1, 3, 7, 16)]
cfgs += [(32, s, m, 2048, 256, 64) for s in (2, 4, 8) for m in (5, 20, 32)]
cfgs += [(64, s, m, 1024, 256, 32) for s in (2, 4) for m in (9, 33, 64)]
cfgs += [(64, 4, 150, 1280, 384, 128)]
cfgs += [(128, s, 100, 1024, 256, 128) for s in (2, 4)]
cfgs += [(256, s, 300, 1024, 256, 128) for s in (2, 4)]
cfgs += [(16, 4, 2, 256, 128, 64), (16, 2, 5, 320, 128, 32), (32, 2, 17, 11008, 128, 128), (16, 2, 2, 192, 128, 64)]
cfgs += [(0, 0, m, 4096, 4096, 128) for m in (1, 8, 16, 64, 128, 256, 512, 1000)]
nbad = 0
for var in variants:
    lib.qb200_debug_set_variant(var)
    for (tok, split, M, K, N, G) in cfgs:
        wq, sz, W = case(K, N, G)
        A = torch.from_numpy(qo.make_activations(M, K, seed=M)).cuda()
        ref = A.float() @ W
        bias = torch.randn(N, device="cuda").half() if (M % 2 == 1) else None
        if bias is not None:
            ref = ref + bias.float()
        if not tok:
            tok, split, _ = ops.plan(M, K, N, G)
        out = ops.gemm(A, wq, sz, N, G, bias=bias, tok=tok or None, split=split or None)
        out2 = ops.gemm(A, wq, sz, N, G, bias=bias, tok=tok or None, split=split or None, independent=True)
        torch.cuda.synchronize()
        rms = ref.pow(2).mean().sqrt().item()
        ok = bool(torch.allclose(out.float(), ref, rtol=1e-2, atol=1e-2 * rms)) and bool(torch.equal(out, out2))
        nbad += (not ok)
        print(json.dumps({"var": var, "cfg": [tok, split, M, K, N, G], "ok": ok,
                          "max_err_over_rms": round((out.float() - ref).abs().max().item() / rms, 5)}), flush=True)
print("BAD", nbad)
sys.exit(1 if nbad else 0)
