"""End to end through the plugin surface at a non-toy size, no network: random-init HF Llama (GQA) → fp16 checkpoint
on disk → AutoAWQForCausalLM.from_pretrained → quantize (AWQ search on the GPU, random-token calibration) →
save_quantized → examples/benchmark.py on the saved checkpoint (from_quantized + fused runner).  Prints one JSON line
with the stage timings; the benchmark rows go to --out."""
import argparse, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--hidden", type=int, default=2048)
ap.add_argument("--inter", type=int, default=5632)
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--heads", type=int, default=16)
ap.add_argument("--kv_heads", type=int, default=4)
ap.add_argument("--vocab", type=int, default=32000)
ap.add_argument("--family", default="llama", choices=["llama", "mistral"])
ap.add_argument("--batch", type=int, nargs="+", default=[1, 8])
ap.add_argument("--out", default="")
args = ap.parse_args()

import transformers
from quick_b200.awq import AutoAWQForCausalLM

kw = dict(hidden_size=args.hidden, intermediate_size=args.inter, num_hidden_layers=args.layers, num_attention_heads=args.heads,
          num_key_value_heads=args.kv_heads, vocab_size=args.vocab, max_position_embeddings=4096)
with tempfile.TemporaryDirectory() as tmp:
    t = {}
    torch.manual_seed(0)
    t0 = time.perf_counter()
    if args.family == "llama":
        hf = transformers.LlamaForCausalLM(transformers.LlamaConfig(**kw)).half()
    else:
        hf = transformers.MistralForCausalLM(transformers.MistralConfig(sliding_window=4096, **kw)).half()
    n_params = sum(p.numel() for p in hf.parameters())
    hf.save_pretrained(os.path.join(tmp, "fp16")); del hf
    t["build_fp16_s"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    model = AutoAWQForCausalLM.from_pretrained(os.path.join(tmp, "fp16"), device_map="cuda")
    calib = torch.randint(0, args.vocab, (8, 512), generator=torch.Generator().manual_seed(1))
    torch.cuda.synchronize(); t["load_fp16_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    model.quantize(None, quant_config={"zero_point": True, "q_group_size": 128, "w_bit": 4, "version": "QUICK"}, calib_data=calib)
    torch.cuda.synchronize(); t["quantize_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    model.save_quantized(os.path.join(tmp, "quick")); del model; torch.cuda.empty_cache()
    t["save_s"] = time.perf_counter() - t0
    size = sum(os.path.getsize(os.path.join(tmp, "quick", f)) for f in os.listdir(os.path.join(tmp, "quick")))

    rows_path = os.path.join(tmp, "rows.json")
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "benchmark.py"), "--model_path", os.path.join(tmp, "quick"),
                        "--batch_size", *map(str, args.batch), "--out", rows_path], capture_output=True, text=True, timeout=600)
    t["benchmark_s"] = time.perf_counter() - t0
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-3000:]); sys.exit(1)
    rows = json.load(open(rows_path))
res = {"family": args.family, "params_M": round(n_params / 1e6, 1), "checkpoint_MB": round(size / 1e6, 1),
       **{k: round(v, 2) for k, v in t.items()}, "rows": rows}
print(json.dumps(res))
if args.out:
    json.dump(res, open(args.out, "w"), indent=1)
