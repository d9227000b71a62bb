"""One CFG-fused DiT forward (F5TTS_Base, 2 x 16 utterances x 1875 frames) + Euler update, for ncu captures.
    python tools/prof_forward.py [n_forwards]      (markers: cudaProfilerStart/Stop around the measured forwards)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from eraxvif5tts_b200 import ops  # noqa: E402

nfw = int(sys.argv[1]) if len(sys.argv) > 1 else 1
arch = bench.Arch(dim=1024, depth=22, heads=16)
dev = torch.device("cuda", 0)
model, voc = bench.build_product_models(arch, dev)
B, ref, n = 16, 563, 1875
cond, text, duration, lens, wav = bench.make_inputs(arch, B, ref, n, 1234)
eng = model.transformer.engine()
te_c = eng.text_embed(text.to(dev), n, False)
te_u = eng.text_embed(text.to(dev), n, True)
step_cond = torch.nn.functional.pad(cond.to(dev), (0, 0, 0, n - ref))
c0 = torch.cat((eng.input_const(step_cond, te_c), eng.input_const(None, te_u)), 0)
mod = eng.modulation(torch.tensor([0.1, 0.2], device=dev))
y = torch.randn(B, n, 100, device=dev)
yb = torch.empty(B * n, 128, dtype=torch.bfloat16, device=dev)
ops.pack_bf16(y.view(B * n, 100), yb, 100, 128)
pred = torch.empty(2 * B, n, 100, device=dev)
lens32 = duration.to(dev).to(torch.int32)
eng.forward(yb, B, c0, 2 * B, n, mod[0], 0, lens32, pred)  # warm-up
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for i in range(nfw):
    eng.forward(yb, B, c0, 2 * B, n, mod[1], 0, lens32, pred)
    ops.cfg_euler(y, pred[:B], pred[B:], 2.0, 0.03, yb)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
assert torch.isfinite(y).all()
print("prof_forward ok")
