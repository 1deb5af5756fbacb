"""Decode-attention cluster-size sweep (development tool): µs per qb200_attn_decode call in a chain of 32 calls over
distinct KV caches (one per "layer"), for forced cluster sizes.  python tools/tune_attn.py --out gpurun_out/tune_attn.json"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import quick_kernels

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--layers", type=int, default=32)
args = ap.parse_args()
rows = []
for (nh, nkv, hd, name) in ((32, 32, 128, "7b"), (64, 8, 128, "70b")):
    for S, p in ((256, 192), (2048, 1984), (8192, 8000)):
        ang = torch.zeros(S, hd, device="cuda", dtype=torch.float16)
        cos, sin = ang + 1, ang
        for B in (1, 8, 64):
            if B * nkv * S * hd * 2 * 2 * args.layers > 40e9:
                continue
            caches = [(torch.randn(B, nkv, S, hd, device="cuda", dtype=torch.float16), torch.randn(B, nkv, S, hd, device="cuda", dtype=torch.float16))
                      for _ in range(args.layers)]
            qkv = torch.randn(B, 1, (nh + 2 * nkv) * hd, device="cuda", dtype=torch.float16)
            pos = torch.tensor([p], device="cuda")
            for split in (0, 1, 2, 4, 8):
                if split:
                    os.environ["QB200_ATTN_SPLIT"] = str(split)
                else:
                    os.environ.pop("QB200_ATTN_SPLIT", None)
                def run():
                    for ck, cv in caches:
                        quick_kernels.attn_decode(qkv, cos, sin, pos, ck, cv, nh, nkv)
                run(); torch.cuda.synchronize()
                side, graph = torch.cuda.Stream(), torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(graph, stream=side):
                        run()
                torch.cuda.synchronize()
                ts = []
                for _ in range(20):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); graph.replay(); b.record(); torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b) * 1e3 / args.layers)
                ts.sort()
                kv_mb = B * nkv * (p + 1) * hd * 2 * 2 / 1e6
                r = {"model": name, "S": S, "pos": p, "B": B, "split": split or "auto", "us": round(ts[len(ts) // 2], 2),
                     "kv_MB": round(kv_mb, 1), "GBs": round(kv_mb / ts[len(ts) // 2] * 1e3, 0)}
                rows.append(r); print(json.dumps(r), flush=True)
            del caches; torch.cuda.empty_cache()
os.environ.pop("QB200_ATTN_SPLIT", None)
if args.out:
    json.dump(rows, open(args.out, "w"), indent=1)
