"""One decode step of the Llama-like runner (few layers) for an ncu launch list: which kernels make up a token."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import copy, torch
from quick_b200.awq.models.llama_like import PRESETS, LlamaLikeQuickModel
name, layers, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = copy.deepcopy(PRESETS[name]); cfg.num_layers = layers; cfg.max_seq_len = 256
torch.manual_seed(0)
m = LlamaLikeQuickModel(cfg, B)
ids = torch.randint(0, cfg.vocab_size, (B, 128), device="cuda"); pos = torch.arange(128, device="cuda")
m(ids, pos); torch.cuda.synchronize()
tok = torch.zeros(B, 1, dtype=torch.long, device="cuda"); p1 = torch.tensor([128], device="cuda")
for _ in range(3): m(tok, p1)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("decode_step")
m(tok, p1)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done")
