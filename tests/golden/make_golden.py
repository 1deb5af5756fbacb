"""Generate golden vectors from the reference's own Python code (run in the build
container only: needs /root/reference; the GPU box and the tests never run this).

  python tests/golden/make_golden.py

* ``pack_*.npz``  — WQLinear_QUICK.from_linear (quick/awq/modules/linear/quick.py:60-156)
  executed on CPU: the module text is exec'd with a stub ``quick_kernels`` and the
  hard-coded device 'cuda' (quick.py:95,101,147) replaced by 'cpu'.  Inputs are
  integer q/z and fp16 s; the fp32 weight handed to from_linear is (q - z)·s so
  that the reference's ``round((W + z·s)/s)`` (quick.py:76-81) recovers q exactly.
* ``cat_*.npz``   — QUICK_cat (quick/awq/utils/fused_utils.py:119-159), function
  source extracted with ``ast`` (the module itself needs awq_ext to import).
"""
import ast
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_ref_quick_module():
    src = open(f"{REF}/quick/awq/modules/linear/quick.py").read().replace("'cuda'", "'cpu'")
    stub = types.ModuleType("quick_kernels")
    stub.gemm_forward_cuda_quick = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub"))
    sys.modules["quick_kernels"] = stub
    mod = types.ModuleType("ref_quick")
    exec(compile(src, "ref_quick.py", "exec"), mod.__dict__)
    return mod


def load_ref_quick_cat():
    src = open(f"{REF}/quick/awq/utils/fused_utils.py").read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "QUICK_cat"][0]
    ns = {"torch": torch}
    exec("from typing import Optional, Tuple\n" + ast.get_source_segment(src, fn), ns)
    return ns["QUICK_cat"]


def main():
    ref = load_ref_quick_module()
    cat = load_ref_quick_cat()
    shapes = [(128, 128, 128), (256, 256, 128), (128, 512, 32), (256, 768, 64), (128, 512, 128), (512, 512, 128)]
    packed = {}
    for (K, N, G) in shapes:
        rng = np.random.default_rng(K * 7 + N * 3 + G)
        q = rng.integers(0, 16, size=(K, N), dtype=np.int32)
        z = rng.integers(0, 16, size=(K // G, N), dtype=np.int32)
        s = (0.002 + 0.01 * rng.random((K // G, N))).astype(np.float16)
        tq, tz, ts = torch.from_numpy(q), torch.from_numpy(z), torch.from_numpy(s).float()
        W = ((tq - tz.repeat_interleave(G, 0)).float() * ts.repeat_interleave(G, 0)).t().contiguous()  # (N, K)
        lin = torch.nn.Linear(K, N, bias=False)
        lin.weight.data = W
        mod = ref.WQLinear_QUICK.from_linear(lin, 4, G, False, scales=ts.t().contiguous(), zeros=tz.t().contiguous().float())
        out = dict(q=q.astype(np.uint8), z=z.astype(np.uint8), s=s, G=np.int32(G),
                   qweight=mod.qweight.numpy(), qzeros=mod.qzeros.numpy(), scales=mod.scales.numpy())
        np.savez_compressed(os.path.join(HERE, f"pack_K{K}_N{N}_G{G}.npz"), **out)
        packed[(K, N, G)] = mod
        print("packed", K, N, G, out["qweight"].shape, out["qzeros"].shape, out["scales"].shape)

    # QUICK_cat goldens: three equal-shape layers (the only case the reference accepts)
    K, N, G = 256, 256, 128
    mods = []
    for i in range(3):
        rng = np.random.default_rng(100 + i)
        q = rng.integers(0, 16, size=(K, N), dtype=np.int32)
        z = rng.integers(0, 16, size=(K // G, N), dtype=np.int32)
        s = (0.002 + 0.01 * rng.random((K // G, N))).astype(np.float16)
        tq, tz, ts = torch.from_numpy(q), torch.from_numpy(z), torch.from_numpy(s).float()
        W = ((tq - tz.repeat_interleave(G, 0)).float() * ts.repeat_interleave(G, 0)).t().contiguous()
        lin = torch.nn.Linear(K, N, bias=False)
        lin.weight.data = W
        m = ref.WQLinear_QUICK.from_linear(lin, 4, G, False, scales=ts.t().contiguous(), zeros=tz.t().contiguous().float())
        mods.append((q, z, s, m))
    out = {}
    for i, (q, z, s, m) in enumerate(mods):
        out[f"q{i}"] = q.astype(np.uint8); out[f"z{i}"] = z.astype(np.uint8); out[f"s{i}"] = s
        out[f"qweight{i}"] = m.qweight.numpy(); out[f"qzeros{i}"] = m.qzeros.numpy(); out[f"scales{i}"] = m.scales.numpy()
    out["cat_qweight"] = cat(*[m.qweight for *_, m in mods], options="qweight").numpy()
    out["cat_qzeros"] = cat(*[m.qzeros for *_, m in mods], options="qzeros").numpy()
    out["cat_scales"] = cat(*[m.scales for *_, m in mods], options="scales").numpy()
    out["G"] = np.int32(G)
    np.savez_compressed(os.path.join(HERE, "cat_K256_3xN256_G128.npz"), **out)
    print("cat", out["cat_qweight"].shape, out["cat_qzeros"].shape, out["cat_scales"].shape)


if __name__ == "__main__":
    main()
