"""Determinism stress + timing of the attention backward: dK / dV must be bit-identical run to run (dQ only up to fp32 atomics order)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eraxvif5tts_b200 import _lib as L, ops  # noqa: E402

B, H, n = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (32, 16, 1200)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 30
dev = torch.device("cuda", 0)
D = H * 64
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(B * n, 3 * D, generator=g) * 1.2).to(dev).bfloat16()
dout = (torch.randn(B * n, D, generator=g) * 0.5).to(dev).bfloat16()
table = torch.empty(n, 64, dtype=torch.float32, device=dev)
L.check(L.load().f5b_rope_table(table.data_ptr(), n, L.stream()), "rope")
out = torch.empty(B * n, D, dtype=torch.bfloat16, device=dev)
lse = torch.empty(B, H, n, dtype=torch.float32, device=dev)
ops.attn_fwd_lse(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, lse, None, 0, B, H, n)
ref = None
bad = 0
for i in range(reps):
    dqkv = torch.full((B * n, 3 * D), float("nan"), dtype=torch.bfloat16, device=dev)
    ops.attn_bwd(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, dout, lse, dqkv, None, 0, B, H, n, rope=table, rope_heads=1)
    torch.cuda.synchronize()
    if ref is None:
        ref = dqkv.clone()
        assert torch.isfinite(ref.float()).all()
        continue
    same_kv = torch.equal(dqkv[:, D:], ref[:, D:])
    dq_err = (dqkv[:, :D].float() - ref[:, :D].float()).abs().max().item() / max(ref[:, :D].float().abs().max().item(), 1e-30)
    if not same_kv or dq_err > 1e-2:
        bad += 1
        diff = (dqkv[:, D:].float() - ref[:, D:].float()).abs()
        rows = torch.nonzero(diff.amax(dim=1) > 0).flatten()
        print(f"run {i}: dk/dv identical={same_kv} dq rel diff={dq_err:.3e} differing rows {rows[:8].tolist()} .. n={rows.numel()} max {diff.max().item():.3e}")
print("stress: bad runs", bad, "of", reps - 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dqkv = torch.empty(B * n, 3 * D, dtype=torch.bfloat16, device=dev)
e0.record()
for _ in range(10):
    ops.attn_bwd(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, dout, lse, dqkv, None, 0, B, H, n, rope=table, rope_heads=1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"attn_bwd B{B} H{H} n{n}: {ms:.3f} ms  {10.0 * B * H * n * n * 64 / ms / 1e9:.0f} TFLOP/s (incl. delta + finish kernels)")
