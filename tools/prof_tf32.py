"""one tf32-mode DiT forward at cfg-2's per-step shape (for ncu): python tools/prof_tf32.py [B] [n]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import f5_oracle as O
from helpers import build_cfm
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1875
cfg = O.DiTConfig(depth=2)
model, sd = build_cfm(cfg, 0)
tr = model.transformer.set_precision("tf32")
g = torch.Generator().manual_seed(0)
x = torch.randn(B, n, cfg.mel_dim, generator=g).cuda()
cond = torch.randn(B, n, cfg.mel_dim, generator=g).cuda()
text = torch.randint(0, 100, (B, 300), generator=g).cuda()
for _ in range(2):
    out = tr(x=x, cond=cond, text=text, time=torch.tensor(0.3).cuda(), drop_audio_cond=False, drop_text=False, mask=None)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
