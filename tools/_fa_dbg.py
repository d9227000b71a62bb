import ctypes as C, os, sys, torch
sys.path.insert(0, os.getcwd())
from eraxvif5tts_b200 import _lib as L, ops
L.load(); raw = C.CDLL(L.LIB_PATH); raw.f5b_debug_attn_variant(1); raw.f5b_debug_attn_poly(0)
dev = "cuda"
shape = tuple(int(x) for x in sys.argv[1:4])
B, H, n = shape
D = H * 64
qkv = torch.randn(B * n, 3 * D, device=dev).bfloat16()
out = torch.empty(B * n, D, dtype=torch.bfloat16, device=dev)
try:
    ops.attn_fwd(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, None, 0, B, H, n)
    torch.cuda.synchronize()
    print(shape, "ok", float(out.float().abs().mean()))
except Exception as e:
    print(shape, "FAIL", str(e)[:80])
