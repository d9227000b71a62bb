"""attention kernel alone (cfg-2 per-layer shape at reduced batch) for ncu"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eraxvif5tts_b200 import ops
B, H, n = 32, 16, 1875
dev = "cuda"
D = H * 64
qkv = torch.randn(B * n, 3 * D, device=dev).to(torch.bfloat16)
out = torch.empty(B * n, H * 64, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.attn_fwd(qkv, qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, None, 0, B, H, n)
torch.cuda.synchronize()
print("ok")
