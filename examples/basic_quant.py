"""Quantize an fp16 Llama / Mistral checkpoint to AWQ-QUICK and save it — the counterpart of the reference's
examples/basic_quant.py.  Calibration text comes from a local file (one sample per line): there is no dataset
download in this library.

  python examples/basic_quant.py --model_path /path/to/fp16 --quant_path /path/to/out --calib_file calib.txt
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from quick_b200.awq import AutoAWQForCausalLM


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model_path", required=True)
    ap.add_argument("--quant_path", required=True)
    ap.add_argument("--calib_file", required=True, help="text file, one calibration sample per line")
    ap.add_argument("--q_group_size", type=int, default=128)
    ap.add_argument("--n_samples", type=int, default=128)
    args = ap.parse_args()
    from transformers import AutoTokenizer
    quant_config = {"zero_point": True, "q_group_size": args.q_group_size, "w_bit": 4, "version": "QUICK"}
    model = AutoAWQForCausalLM.from_pretrained(args.model_path, device_map="cuda")
    tokenizer = AutoTokenizer.from_pretrained(args.model_path, trust_remote_code=True)
    with open(args.calib_file, encoding="utf-8") as f:
        calib = [line for line in f if line.strip()]
    model.quantize(tokenizer, quant_config=quant_config, calib_data=calib, n_samples=args.n_samples)
    model.save_quantized(args.quant_path)
    tokenizer.save_pretrained(args.quant_path)
    print(f'Model is quantized and saved at "{args.quant_path}"')


if __name__ == "__main__":
    main()
