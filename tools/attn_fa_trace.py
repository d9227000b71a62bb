"""clock64 trace of the persistent attention kernel (attention_fa.cu): per key tile, when each softmax warpgroup waited for S,
loaded it, finished the maximum, finished the exponentials and published P, and when the MMA thread saw P / issued the products.
Needs the instrumented build:  F5B_BUILD_TAG=trace F5B_NVCC_EXTRA=-DFA_TRACE python -m eraxvif5tts_b200.build
then  F5B_LIB=eraxvif5tts_b200/lib/libf5b200_trace.so python tools/attn_fa_trace.py [poly]"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eraxvif5tts_b200 import ops, _lib as L
lib = L.load()
raw = ctypes.CDLL(L.LIB_PATH)
raw.f5b_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
raw.f5b_debug_attn_variant(1)
raw.f5b_debug_attn_poly(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
B, H, n = 32, 16, 1875
dev = "cuda"
D = H * 64
qkv = torch.randn(B * n, 3 * D, device=dev).to(torch.bfloat16)
out = torch.empty(B * n, H * 64, dtype=torch.bfloat16, device=dev)
tr = torch.zeros(3 * 64 * 8, dtype=torch.int64, device=dev)
for _ in range(2):
    ops.attn_fwd(qkv, qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, None, 0, B, H, n)
torch.cuda.synchronize()
raw.f5b_debug_set_attn_trace(tr.data_ptr())
ops.attn_fwd(qkv, qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, None, 0, B, H, n)
torch.cuda.synchronize()
raw.f5b_debug_set_attn_trace(None)
t = tr.cpu().view(3, 64, 8)
t0 = int(t[t > 0].min())
print("softmax groups, per 64-key tile: [tile start, gap start, gap done]  (group 0: gap behind the tile; group 1: the previous tile's gap, half way through) (clk since first stamp)")
for wg in range(2):
    for i in range(40):
        r = [int(x) - t0 for x in t[wg, i, :3]]
        if r[1] < 0:
            continue
        print(f"  wg{wg} tile {i:2d}: " + " ".join(f"{x:7d}" for x in r) + "   d: to gap %5d gap %4d" % (r[1] - r[0], r[2] - r[1]))
print("MMA warp of tile 0, per key tile: [wait P start, P seen, P V issued, S(+3) issued]")
for i in range(40):
    r = [int(x) - t0 for x in t[2, i, :4]]
    if r[1] < 0:
        continue
    print(f"  step {i:2d}: " + " ".join(f"{x:7d}" for x in r) + "   d: waitP %5d pv %4d s %4d" % (r[1] - r[0], r[2] - r[1], r[3] - r[2]))
