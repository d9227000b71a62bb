"""attention kernel alone (cfg-2 per-layer shape at reduced batch) for ncu"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eraxvif5tts_b200 import ops
B, H, n = 8, 16, 1875
n_pad = (n + 7) // 8 * 8
dev = "cuda"
q = torch.randn(B, H, n, 64, device=dev).to(torch.bfloat16)
k = torch.randn(B, H, n, 64, device=dev).to(torch.bfloat16)
vt = torch.randn(B, H, n, 64, device=dev).to(torch.bfloat16)
out = torch.empty(B * n, H * 64, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.attn_fwd(q, k, vt, out, None, 0, B, H, n, n_pad)
torch.cuda.synchronize()
print("ok")
