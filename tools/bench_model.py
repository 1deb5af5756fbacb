"""tokens/s of a random-init Llama-like AWQ-QUICK model (BASELINE configs[2..3]; reference methodology of
examples/benchmark.py: prefill = decode = 128).  --impl reference runs the same runner with every linear
routed to the unmodified reference kernel (oracle/_ref)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from quick_b200.awq.models.llama_like import PRESETS, LlamaLikeQuickModel, benchmark_generation

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="llama-2-7b", choices=list(PRESETS))
ap.add_argument("--batch", type=int, nargs="+", default=[1, 8, 32, 64])
ap.add_argument("--ctx", type=int, default=128)
ap.add_argument("--gen", type=int, default=128)
ap.add_argument("--layers", type=int, default=0, help="override layer count (0 = preset)")
ap.add_argument("--impl", default="quick_b200", choices=["quick_b200", "reference"])
ap.add_argument("--generate", action="store_true", help="also time model.generate() (the user-facing call of the plugin "
                "surface: prefill + graph decode + token pick + host loop), wall clock around the call")
ap.add_argument("--out", default="")
args = ap.parse_args()
# tensor parallel: `python -m torch.distributed.run --nproc-per-node R tools/bench_model.py --model llama-2-70b`
# (one process per GPU; every linear column-parallel + one NCCL all-gather, SURVEY §8e / BASELINE config 5)
world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0))
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG", "NONE")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
cfg = PRESETS[args.model]
if args.layers:
    cfg.num_layers = args.layers
cfg.max_seq_len = args.ctx + args.gen
rows = []
torch.manual_seed(1234)   # replicated parts (embedding, lm_head) and the prompt are identical on every rank
for bs in args.batch:
    model = LlamaLikeQuickModel(cfg, bs)
    if args.impl == "reference":
        from oracle.build_ref import load_ref
        model.ref_mod = load_ref(); assert model.ref_mod is not None
    r = benchmark_generation(model, args.ctx, args.gen)
    if args.generate and world == 1 and args.impl == "quick_b200":
        import time
        ids = torch.randint(0, cfg.vocab_size, (bs, args.ctx), device="cuda")
        model.generate(ids, max_new_tokens=4)          # builds the decode graph
        torch.cuda.synchronize(); t0 = time.perf_counter()
        seq = model.generate(ids, max_new_tokens=args.gen)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        assert seq.shape == (bs, args.ctx + args.gen)
        r["generate_tokens_per_s"] = round(bs * args.gen / dt, 1)     # includes the prefill of ctx tokens
        r["generate_s"] = round(dt, 4)
    r.update({"model": args.model, "impl": args.impl, "layers": cfg.num_layers, "weight_GB_per_rank": round(model.weight_bytes() / 1e9, 2),
              "mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2), "tp": world})
    rows.append(r)
    if rank == 0:
        print(json.dumps(r), flush=True)
    del model; torch.cuda.empty_cache()
if args.out and rank == 0:
    json.dump(rows, open(args.out, "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
