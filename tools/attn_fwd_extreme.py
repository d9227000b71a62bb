"""Attention forward against an fp64 reference on inputs with very large score variance (std up to 6: single-tile jumps of the row maximum
beyond 2^128, i.e. speculative exponentials that overflow) — the robustness check that rejected a row-sum-based rescale variant (NaN at std 6)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eraxvif5tts_b200 import ops
dev = torch.device("cuda", 0)
for (B, H, n, std) in ((2, 16, 1200, 4.0), (2, 16, 1200, 2.5), (1, 16, 1875, 1.2), (2, 4, 700, 6.0)):
    D = H * 64
    g = torch.Generator().manual_seed(0)
    qkv = (torch.randn(B * n, 3 * D, generator=g) * std).to(dev).bfloat16()
    q, k, v = (qkv[:, i * D:(i + 1) * D].double().reshape(B, n, H, 64) for i in range(3))
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * 0.125
    ref = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), v).reshape(B * n, D)
    out = torch.full((B * n, D), float("nan"), dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, H, n, dtype=torch.float32, device=dev)
    ops.attn_fwd_lse(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, lse, None, 0, B, H, n)
    torch.cuda.synchronize()
    err = (out.double() - ref).abs()
    ref_lse = torch.logsumexp(s, -1) * 1.4426950408889634
    print(f"std {std} n {n}: out max|err| {err.max().item():.4f} rel-fro {(err.norm() / ref.norm()).item():.3e} | lse max|err| {(lse.double() - ref_lse).abs().max().item():.3e}  (lib {os.environ.get('F5B_LIB', 'default')})")
