import torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eraxvif5tts_b200 import ops
B, H, n = 1, 1, 128
D = 64
def run(q, k, v):
    qkv = torch.cat([q, k, v], 1).contiguous().cuda()
    out = torch.full((n, D), float('nan'), device='cuda')
    ops.attn_fwd_tf32(qkv, qkv[:, D:], qkv[:, 2*D:], 3*D, out, None, 0, B, H, n)
    torch.cuda.synchronize()
    return out.cpu()
z = torch.zeros(n, D)
o = run(z, z, torch.ones(n, D)); print("V=1:", o[0, :4], o[100, 60:], o.abs().max())
vd = torch.arange(D).float()[None, :].expand(n, D).contiguous()
o = run(z, z, vd); print("V=d:", o[0, :8], o[5, 30:36], o[127, 60:])
vk = torch.arange(n).float()[:, None].expand(n, D).contiguous()
o = run(z, z, vk); print("V=key (expect 63.5):", o[0, :4], o[64, 32:36])
# one-hot attention: q_i . k_j large only for i == j
q = torch.zeros(n, D); k = torch.zeros(n, D)
for i in range(n):
    q[i, i % 64] = 40.0 * (1 if i < 64 else -1); k[i, i % 64] = 40.0 * (1 if i < 64 else -1)
o = run(q, k, vk); print("one-hot (expect row i -> ~i):", o[:6, 0], o[60:68, 0], o[120:, 5])
