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
print("softmax groups: per key tile  [wait S start, S ready, loaded, max done, exps done, P published]  (clk since first stamp); d = deltas")
for wg in range(2):
    for i in range(34):
        r = [int(x) - t0 for x in t[wg, i, :6]]
        if r[1] < 0:
            continue
        d = [r[k + 1] - r[k] for k in range(5)]
        print(f"  wg{wg} tile {i:2d}: " + " ".join(f"{x:7d}" for x in r) + "   d: wait %5d ld %4d max %4d exp %5d st+arrive %4d" % tuple(d))
print("MMA thread: per P V issue [wait P start, P seen, operands ready, issued]")
for i in range(64):
    r = [int(x) - t0 for x in t[2, i, :6]]
    if r[1] < 0:
        continue
    r = [r[0], r[1], r[3], r[4]]
    print(f"  pv {i:2d}: " + " ".join(f"{x:7d}" for x in r) + "   d: waitP %5d waitV %4d issue %4d" % tuple(r[k + 1] - r[k] for k in range(3)))
