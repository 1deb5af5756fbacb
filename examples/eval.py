"""Perplexity of an AWQ-QUICK (or fp16) model over a local text file — the ``--tasks wikitext`` leg of the reference's
examples/eval.py with caller-supplied text (no dataset download).

  python examples/eval.py --model_path /path/to/quick-checkpoint --text_file wiki.test.raw
  python examples/eval.py --model_path /path/to/fp16 --use_pretrained --text_file wiki.test.raw
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from quick_b200.awq import AutoAWQForCausalLM
from quick_b200.awq.evaluation import evaluate_perplexity


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model_path", required=True)
    ap.add_argument("--quant_file", default="")
    ap.add_argument("--text_file", required=True)
    ap.add_argument("--use_pretrained", default=False, action="store_true", help="evaluate the fp16 model instead")
    ap.add_argument("--seqlen", type=int, default=2048)
    args = ap.parse_args()
    from transformers import AutoTokenizer
    tokenizer = AutoTokenizer.from_pretrained(args.model_path, trust_remote_code=True)
    if args.use_pretrained:
        model = AutoAWQForCausalLM.from_pretrained(args.model_path, device_map="cuda")
    else:
        model = AutoAWQForCausalLM.from_quantized(args.model_path, args.quant_file, max_new_tokens=args.seqlen, batch_size=1)
    with open(args.text_file, encoding="utf-8") as f:
        text = f.read()
    print(f"Perplexity: {evaluate_perplexity(model, tokenizer, text, seqlen=args.seqlen):.3f}")


if __name__ == "__main__":
    main()
