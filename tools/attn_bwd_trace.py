"""Per-phase clock64 breakdown of the attention-backward softmax warps (needs a build with -DAB_TRACE)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eraxvif5tts_b200 import _lib as L, ops  # noqa: E402
B, H, n = 32, 16, 1200
dev = torch.device("cuda", 0)
D = H * 64
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(B * n, 3 * D, generator=g) * 1.2).to(dev).bfloat16()
dout = (torch.randn(B * n, D, generator=g) * 0.5).to(dev).bfloat16()
out = torch.empty(B * n, D, dtype=torch.bfloat16, device=dev)
lse = torch.empty(B, H, n, dtype=torch.float32, device=dev)
ops.attn_fwd_lse(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, lse, None, 0, B, H, n)
dqkv = torch.empty(B * n, 3 * D, dtype=torch.bfloat16, device=dev)
trace = torch.zeros(64, dtype=torch.int64, device=dev)
lib = C.CDLL(L.lib_path()) if hasattr(L, "lib_path") else L.load()
lib.f5b_debug_set_attn_bwd_trace.argtypes = [C.c_void_p]
lib.f5b_debug_set_attn_bwd_trace(trace.data_ptr())
for _ in range(3):
    ops.attn_bwd(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, dout, lse, dqkv, None, 0, B, H, n)
torch.cuda.synchronize()
t = trace.cpu()
T = int(t[6]) or 1
names = ["wait sfree", "issue S/dP(j+1)", "wait pds", "issue dV,dK,dQ + commits", "until dq(j) completes"]
print("MMA warp per tile: " + "  ".join(f"{nm} {int(t[i]) / T:7.0f}" for i, nm in enumerate(names)) + f"   sum {int(t[:5].sum()) / T:7.0f}")
names2 = ["setup", "loads + S_0/dP_0", "first softmax pass", "rest of the loop", "dV/dK drain", "CTA total"]
print("CTA timeline (clocks): " + "  ".join(f"{nm} {int(t[8 + i])}" for i, nm in enumerate(names2)) + f"   tiles {T}")
import time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attn_bwd(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, dout, lse, dqkv, None, 0, B, H, n)
e1.record()
torch.cuda.synchronize()
print(f"attn_bwd (+ delta, finish) {e0.elapsed_time(e1) / 10:.3f} ms per call; CTAs per SM {((n + 127) // 128) * B * H / 148:.1f}")
