"""Per-phase clock64 trace of the attention softmax warps (needs a build with F5B_NVCC_EXTRA=-DATT_TRACE)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eraxvif5tts_b200 import ops, _lib as L
lib = L.load()
lib.f5b_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
B, H, n = 8, 16, 1875
dev = "cuda"
D = H * 64
qkv = torch.randn(B * n, 3 * D, device=dev).to(torch.bfloat16)
out = torch.empty(B * n, H * 64, dtype=torch.bfloat16, device=dev)
tr = torch.zeros(64, dtype=torch.int64, device=dev)
lib.f5b_debug_set_attn_trace(tr.data_ptr())
for _ in range(3):
    ops.attn_fwd(qkv, qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, None, 0, B, H, n)
torch.cuda.synchronize()
t = tr.cpu().view(8, 8)
names = ["wait bar_s", "LDTM+wait", "max/rescale", "wait bar_pv(P buf)", "exp+pack+STS", "fence+arrive"]
for w in range(4):
    T = int(t[w, 6])
    print(f"warp {w}: tiles {T}  per-tile clk: " + ", ".join(f"{names[i]}={int(t[w, i]) / max(T, 1):.0f}" for i in range(6)) +
          f"  total={sum(int(t[w, i]) for i in range(6)) / max(T, 1):.0f}")
