"""tokens/s of an AWQ-QUICK model through the plugin surface — the counterpart of the reference's examples/benchmark.py
(same command line, same methodology: prefill tokens/s = context·batch / prefill seconds, decode tokens/s = batch /
median(decode-step seconds), benchmark.py:38-67,127-129; CUDA events around every model call).

  python examples/benchmark.py --model_path /path/to/quick-checkpoint --batch_size 8
  python examples/benchmark.py --random_init llama-2-7b --batch_size 1 8 64        # no checkpoint: random-init weights

``--generator torch`` drives the model token by token through its stateful forward (the reference's ``generate_torch``);
``--generator hf`` calls ``model.generate`` (the user-facing call, wall clock).  ``--pretrained`` (fp16 HF model) is the
reference's FP16 baseline row and needs an fp16 checkpoint directory.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def generate_torch(model, input_ids, n_generate):
    context_time, generate_time = 0.0, []
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    token = None
    with torch.inference_mode():
        for i in range(n_generate):
            inputs = input_ids if i == 0 else token
            start.record()
            out = model(inputs, use_cache=True)
            end.record()
            torch.cuda.synchronize()
            token = out[0][:, -1].max(1)[1].unsqueeze(1)
            if i == 0:
                context_time = start.elapsed_time(end) * 1e-3
            else:
                generate_time.append(start.elapsed_time(end) * 1e-3)
    return context_time, generate_time


def generate_hf(model, input_ids, n_generate):
    """Whole-call wall clock of model.generate: returns (0, [per-token average]) — the fused runner has no per-token hook."""
    model.generate(input_ids, max_new_tokens=2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.generate(input_ids, max_new_tokens=n_generate)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return 0.0, [dt / n_generate] * n_generate


def load(args, batch_size, n_context, n_generate):
    if args.random_init:
        import copy
        from quick_b200.awq.models.llama_like import PRESETS, LlamaLikeQuickModel
        cfg = copy.deepcopy(PRESETS[args.random_init])
        cfg.max_seq_len = n_context + n_generate
        runner = LlamaLikeQuickModel(cfg, batch_size)

        class Stateful(torch.nn.Module):                   # the same two calls a loaded checkpoint offers
            def __init__(self):
                super().__init__()
                self.runner = runner
            forward = staticmethod(runner.forward_stateful)
            generate = staticmethod(runner.generate)
        return Stateful(), cfg.vocab_size
    from quick_b200.awq import AutoAWQForCausalLM
    if args.pretrained:
        m = AutoAWQForCausalLM.from_pretrained(args.model_path, safetensors=not args.no_safetensors, device_map="cuda",
                                               torch_dtype=torch.float16)
    else:
        m = AutoAWQForCausalLM.from_quantized(args.model_path, args.quant_file, max_new_tokens=n_context + n_generate,
                                              batch_size=batch_size, safetensors=not args.no_safetensors)
    return m, m.config.vocab_size


def run_round(args, batch_size, n_context, n_generate):
    torch.cuda.reset_peak_memory_stats()
    model, vocab = load(args, batch_size, n_context, n_generate)
    input_ids = torch.randint(0, vocab, (batch_size, n_context), device="cuda")
    gen = generate_torch if args.generator == "torch" else generate_hf
    gen(model, input_ids, min(n_generate, 4))               # warm-up: B200 weight copies, decode graph
    row = {"Batch Size": batch_size, "Prefill Length": n_context, "Decode Length": n_generate}
    try:
        context_time, generate_time = gen(model, input_ids, n_generate)
        generate_time.sort()
        row["Prefill tokens/s"] = round(n_context * batch_size / context_time, 2) if context_time else None
        row["Decode tokens/s"] = round(batch_size / generate_time[len(generate_time) // 2], 2)
    except torch.cuda.OutOfMemoryError:
        row["Prefill tokens/s"] = row["Decode tokens/s"] = "OOM"
    row["Memory (VRAM)"] = f"{torch.cuda.max_memory_allocated() / 2 ** 30:.2f} GB"
    row["Version"] = "FP16" if args.pretrained else "QUICK (quick_b200)"
    del model
    torch.cuda.empty_cache()
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model_path", type=str, default="", help="path to the (quantized) model directory")
    ap.add_argument("--quant_file", type=str, default="", help="weights filename inside model_path")
    ap.add_argument("--random_init", type=str, default="", help="llama-2-7b | mistral-7b | llama-2-70b | tiny instead of a checkpoint")
    ap.add_argument("--batch_size", type=int, nargs="+", default=[1])
    ap.add_argument("--no_safetensors", default=False, action="store_true")
    ap.add_argument("--generator", type=str, default="torch", choices=["torch", "hf"])
    ap.add_argument("--pretrained", default=False, action="store_true", help="measure the fp16 model instead")
    ap.add_argument("--lengths", type=int, nargs="+", default=[128], help="prefill = decode lengths (reference: 32 … 2048)")
    ap.add_argument("--out", type=str, default="")
    args = ap.parse_args()
    if not args.model_path and not args.random_init:
        ap.error("give --model_path or --random_init")
    rows = [run_round(args, bs, n, n) for n in args.lengths for bs in args.batch_size]
    print(f"GPU: {torch.cuda.get_device_name()}")
    print(f"Model: {args.model_path or args.random_init + ' (random-init)'}")
    for r in rows:
        print(json.dumps(r))
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
