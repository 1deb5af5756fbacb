"""GPU bring-up / diagnosis driver (development tool, not part of the product path).

  python tools/gpu_bringup.py layout            # pack / relayout / dequant bit-exactness
  python tools/gpu_bringup.py gemm TOK SPLIT M K N G
  python tools/gpu_bringup.py sweep             # many configs, each in a subprocess with a timeout
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from oracle import quick_oracle as qo
from quick_b200 import layout, ops


def case(K, N, G, seed=1234):
    q, z, s = qo.make_case(K, N, G, seed)
    W16 = qo.dequant_w16(q, z, s, G)
    tq, tz, ts = torch.from_numpy(q), torch.from_numpy(z), torch.from_numpy(s)
    return q, z, s, W16, tq, tz, ts


def cmd_layout():
    for (K, N, G) in [(128, 128, 128), (512, 512, 128), (256, 768, 64), (128, 512, 32), (4096, 4096, 128)]:
        q, z, s, W16, tq, tz, ts = case(K, N, G)
        qw_c, qz_c, sc_c = layout.pack_quick(tq, tz, ts)
        qw_g, qz_g, sc_g = ops.pack_quick(tq.cuda(), tz.cuda(), ts.cuda(), G)
        ok_pack = torch.equal(qw_g.cpu(), qw_c) and torch.equal(qz_g.cpu(), qz_c) and torch.equal(sc_g.cpu().view(torch.int16), sc_c.view(torch.int16))
        wq, sz, K2, N2, G2 = ops.prepack(qw_g, qz_g, sc_g)
        Wd = ops.dequantize(wq, sz, K, N, G).cpu().numpy()
        ok_deq = np.array_equal(Wd.view(np.uint16), W16.view(np.uint16))
        print(json.dumps({"case": [K, N, G], "gpu_pack_eq_host_pack": ok_pack, "relayout_dequant_bitexact": ok_deq}), flush=True)


def run_gemm(tok, split, M, K, N, G, simt=False, verbose=True):
    q, z, s, W16, tq, tz, ts = case(K, N, G)
    A = qo.make_activations(M, K, seed=M)
    qw, qz, sc = ops.pack_quick(tq.cuda(), tz.cuda(), ts.cuda(), G)
    wq, sz, *_ = ops.prepack(qw, qz, sc)
    x = torch.from_numpy(A).cuda()
    ref = torch.from_numpy(A.astype(np.float32)).cuda() @ torch.from_numpy(W16.astype(np.float32)).cuda()
    if simt:
        out = ops.gemm_simt(x, wq, sz, N, G)
    else:
        out = ops.gemm(x, wq, sz, N, G, tok=tok or None, split=split or None)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs()
    rms = ref.pow(2).mean().sqrt().item()
    rec = {"tok": tok, "split": split, "M": M, "K": K, "N": N, "G": G, "simt": simt,
           "max_abs": err.max().item(), "rms_ref": rms, "max_abs_over_rms": err.max().item() / rms,
           "ok": bool(torch.allclose(out.float(), ref, rtol=1e-2, atol=1e-2 * rms))}
    if verbose and not rec["ok"]:
        bad = (err > 1e-2 * rms + 1e-2 * ref.abs())
        rec["bad_frac"] = bad.float().mean().item()
        rec["bad_rows"] = bad.any(1).nonzero().flatten()[:16].tolist()
        rec["bad_cols"] = bad.any(0).nonzero().flatten()[:32].tolist()
        rec["out_sample"] = out[0, :8].float().tolist()
        rec["ref_sample"] = ref[0, :8].tolist()
        rec["out_nan"] = bool(torch.isnan(out).any().item())
        # does out match the reference with some simple structure?  ratio statistics
        ratio = (out.float() / ref)[ref.abs() > 0.5 * rms]
        rec["ratio_median"] = ratio.median().item() if ratio.numel() else None
    print(json.dumps(rec), flush=True)
    return rec["ok"]


def cmd_sweep():
    cfgs = []
    # first the simplest: one k-block chain, no split
    cfgs += [(16, 1, 1, 64, 128, 64), (16, 1, 16, 128, 128, 128), (16, 1, 16, 512, 512, 128)]
    cfgs += [(t, 1, t, 512, 512, 128) for t in (32, 64, 128, 256)]
    cfgs += [(16, s, 7, 2048, 256, 128) for s in (2, 4, 8)]
    cfgs += [(32, s, 20, 2048, 256, 64) for s in (2, 4, 8)]
    cfgs += [(64, s, 64, 1024, 256, 32) for s in (2, 4)]
    cfgs += [(128, s, 100, 1024, 256, 128) for s in (2, 4)]
    cfgs += [(256, s, 300, 1024, 256, 128) for s in (2, 4)]
    cfgs += [(0, 0, m, 4096, 4096, 128) for m in (1, 8, 16, 64, 128, 256, 512)]
    results = []
    for c in cfgs:
        t0 = time.time()
        p = subprocess.run([sys.executable, __file__, "gemm", *map(str, c)], capture_output=True, text=True, timeout=300)
        line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else ""
        print(f"cfg={c} rc={p.returncode} t={time.time()-t0:.1f}s {line}", flush=True)
        if p.returncode != 0:
            print("  stderr:", p.stderr.strip()[-600:], flush=True)
        results.append({"cfg": c, "rc": p.returncode, "out": line})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "bringup_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "layout":
        cmd_layout()
    elif what == "gemm":
        tok, split, M, K, N, G = map(int, sys.argv[2:8])
        ok = run_gemm(tok, split, M, K, N, G)
        sys.exit(0 if ok else 3)
    elif what == "simt":
        M, K, N, G = map(int, sys.argv[2:6])
        sys.exit(0 if run_gemm(0, 0, M, K, N, G, simt=True) else 3)
    elif what == "sweep":
        cmd_sweep()
