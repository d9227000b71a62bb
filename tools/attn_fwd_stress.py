"""Determinism stress of the attention forward with large score variance (frequent lazy-rescale events)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eraxvif5tts_b200 import ops  # noqa: E402

B, H, n = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 16, 1200)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 100
std = float(sys.argv[5]) if len(sys.argv) > 5 else 1.2
dev = torch.device("cuda", 0)
D = H * 64
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(B * n, 3 * D, generator=g) * std).to(dev).bfloat16()
q, k, v = (qkv[:, i * D:(i + 1) * D].float().reshape(B, n, H, 64) for i in range(3))
s = torch.einsum("bqhd,bkhd->bhqk", q, k) * 0.125
ref = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), v).reshape(B * n, D)
first, bad = None, 0
for i in range(reps):
    out = torch.full((B * n, D), float("nan"), dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, H, n, dtype=torch.float32, device=dev)
    ops.attn_fwd_lse(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, lse, None, 0, B, H, n)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs()
    if first is None:
        first = out.clone()
    same = torch.equal(out, first)
    if not same or err.max().item() > 2e-2:
        bad += 1
        rows = torch.nonzero(err.amax(dim=1) > 2e-2).flatten()
        cols = torch.nonzero(err.amax(dim=0) > 2e-2).flatten()
        print(f"run {i}: identical_to_first={same} max err {err.max().item():.4f} bad rows {rows[:10].tolist()} (n={rows.numel()}) heads {sorted(set((cols // 64).tolist()))}")
print("fwd stress: bad runs", bad, "of", reps)
