"""Where a decode step's time goes (development tool): CUDA graphs of the decoder-layer kernel sequence of a random-init
model with kernels knocked out, over all layers' (cold) weights.  µs per layer = graph time / layers.
  python tools/layer_breakdown.py --model llama-2-7b --batch 1 --out gpurun_out/layer_breakdown.json"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import quick_kernels
from quick_b200.awq.models.llama_like import PRESETS, LlamaLikeQuickModel

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="llama-2-7b", choices=list(PRESETS))
ap.add_argument("--batch", type=int, nargs="+", default=[1])
ap.add_argument("--ctx", type=int, default=192)
ap.add_argument("--out", default="")
args = ap.parse_args()
cfg = PRESETS[args.model]
cfg.max_seq_len = 256
rows = []
for bs in args.batch:
    torch.manual_seed(0)
    model = LlamaLikeQuickModel(cfg, bs)
    dev = model.embed.weight.device
    tok = torch.randint(0, cfg.vocab_size, (bs, 1), device=dev)
    pos = torch.tensor([args.ctx], device=dev)
    model(tok, pos); model(tok, pos)          # builds the B200 copies / interleaved gate|up
    torch.cuda.synchronize()
    H = cfg.hidden_size
    x0 = torch.randn(bs, 1, H, device=dev, dtype=torch.float16) * 0.1
    qkv_w = (cfg.num_heads + 2 * cfg.num_kv_heads) * cfg.head_dim
    o_in = torch.randn(bs, 1, cfg.num_heads * cfg.head_dim, device=dev, dtype=torch.float16) * 0.1
    act_in = torch.randn(bs, 1, cfg.intermediate_size, device=dev, dtype=torch.float16) * 0.1
    rope = (model.rope_cos, model.rope_sin)

    def layers(norm=True, attn=True, qkv=True, o=True, gu=True, down=True):
        x = x0
        for blk in model.blocks:
            xn = blk.norm_1(x) if norm else x
            q = blk.qkv_proj(xn) if qkv else None
            if attn:
                if q is None:
                    q = torch.zeros(bs, 1, qkv_w, device=dev, dtype=torch.float16)
                a = quick_kernels.attn_decode(q, rope[0], rope[1], pos, blk.cache_k, blk.cache_v, cfg.num_heads, cfg.num_kv_heads)
            else:
                a = o_in
            x = blk.o_proj(a, x) if o else x
            xn = blk.norm_2(x) if norm else x
            act = blk.gate_up_proj.forward_silu_mul(xn) if gu else act_in
            x = blk.down_proj(act, x) if down else x
        return x

    variants = {
        "full layer (norm, qkv, attn, o, norm, gate|up, down)": dict(),
        "without attention": dict(attn=False),
        "without the two RMSNorms": dict(norm=False),
        "GEMMs only": dict(norm=False, attn=False),
        "qkv only": dict(norm=False, attn=False, o=False, gu=False, down=False),
        "o only": dict(norm=False, attn=False, qkv=False, gu=False, down=False),
        "gate|up only": dict(norm=False, attn=False, qkv=False, o=False, down=False),
        "down only": dict(norm=False, attn=False, qkv=False, o=False, gu=False),
        "norms only": dict(attn=False, qkv=False, o=False, gu=False, down=False),
        "attention only": dict(norm=False, qkv=False, o=False, gu=False, down=False),
        "norm + qkv": dict(attn=False, o=False, gu=False, down=False),
        "norm + gate|up": dict(attn=False, qkv=False, o=False, down=False),
    }
    res = {}
    for name, kw in variants.items():
        for _ in range(2):
            layers(**kw)
        torch.cuda.synchronize()
        side, graph = torch.cuda.Stream(), torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                layers(**kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(30):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); graph.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        res[name] = round(ts[len(ts) // 2] / len(model.blocks), 2)
        print(json.dumps({"model": args.model, "batch": bs, "variant": name, "us_per_layer": res[name]}), flush=True)
    # the whole decode step for reference
    step = []
    model._decode_graph = None
    for i in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); model._decode_step(tok, args.ctx, True); b.record(); torch.cuda.synchronize()
        step.append(a.elapsed_time(b) * 1e3)
    step.sort()
    res["whole decode step (us)"] = round(step[len(step) // 2], 1)
    print(json.dumps({"model": args.model, "batch": bs, "decode_step_us": res["whole decode step (us)"]}), flush=True)
    rows.append({"model": args.model, "batch": bs, "layers": len(model.blocks), "us_per_layer": res})
    del model; torch.cuda.empty_cache()
if args.out:
    json.dump(rows, open(args.out, "w"), indent=1)
