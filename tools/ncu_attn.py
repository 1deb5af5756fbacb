"""Launch the decode-attention kernel a few times over cold caches (for `ncu --set full -k regex:attn_decode`).
usage: python tools/ncu_attn.py <batch> [nh nkv hd S pos]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import quick_kernels
B = int(sys.argv[1])
nh, nkv, hd, S, p = (int(v) for v in sys.argv[2:7]) if len(sys.argv) > 6 else (32, 32, 128, 256, 192)
cos = torch.ones(S, hd, device="cuda", dtype=torch.float16); sin = torch.zeros_like(cos)
pos = torch.tensor([p], device="cuda")
qkv = torch.randn(B, 1, (nh + 2 * nkv) * hd, device="cuda", dtype=torch.float16)
n = max(4, int(300e6 // (B * nkv * S * hd * 4)))          # > L2 worth of distinct caches
caches = [(torch.randn(B, nkv, S, hd, device="cuda", dtype=torch.float16), torch.randn(B, nkv, S, hd, device="cuda", dtype=torch.float16))
          for _ in range(min(n, 64))]
for ck, cv in caches:
    quick_kernels.attn_decode(qkv, cos, sin, pos, ck, cv, nh, nkv)
torch.cuda.synchronize()
print("done", B, nh, nkv, hd, S, p, len(caches))
