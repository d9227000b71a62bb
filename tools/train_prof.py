"""One training step (F5TTS_Base, 32 x 1200 frames per GPU, bf16, dropout 0): forward + backward + fused AdamW.
    python tools/train_prof.py [B] [n] [steps]     prints ms/step, frames/s, model TFLOP/s and the per-kernel-class breakdown
    F5B_TRAIN_DROPOUT=0.1 runs the step with the reference's train-mode dropout, F5B_TRAIN_ATTN_DROPOUT=0.1 with SDPA's own."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from eraxvif5tts_b200 import _lib as L  # noqa: E402
from eraxvif5tts_b200.train import TrainEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
arch = bench.Arch(dim=1024, depth=22, heads=16)
dev = torch.device("cuda", 0)
model, _ = bench.build_product_models(arch, dev)
eng = TrainEngine(model, with_ema=True, dropout=float(os.environ.get("F5B_TRAIN_DROPOUT", "0")),
                  attn_dropout=float(os.environ.get("F5B_TRAIN_ATTN_DROPOUT", "0")))
g = torch.Generator().manual_seed(0)
mel = (torch.randn(B, n, 100, generator=g) * 2 - 1.5).clamp(-11.5, 5).to(dev)
text = torch.randint(0, arch.text_num_embeds, (B, int(0.16 * n)), generator=g).to(dev)


def step():
    eng.zero_grad()
    loss, _, _ = eng.loss_and_grads(mel, text)
    eng.step()
    return loss


for _ in range(2):
    loss = step()
torch.cuda.synchronize()
print("ws GB", eng._ws.numel() / 1e9, "peak alloc GB", torch.cuda.max_memory_allocated() / 1e9, "loss", float(loss))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
fwd = bench.dit_flops_per_forward(arch, B, n)
print(f"train step {ms:.1f} ms  {B * n / ms * 1e3:.0f} frames/s  model {3 * fwd / ms / 1e9:.0f} TFLOP/s (3x forward FLOPs)  loss {float(loss):.4f}")
L.prof_reset(True)
torch.cuda.cudart().cudaProfilerStart()  # ncu --profile-from-start off captures exactly this step
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
prof = L.prof_read()
L.prof_reset(False)
tot = sum(v["ms"] for v in prof.values())
for k, v in prof.items():
    if v["launches"]:
        tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0
        print(f"  {k:12s} launches {v['launches']:5d}  {v['ms']:8.2f} ms ({100 * v['ms'] / tot:4.1f} %)  {tf:7.1f} TFLOP/s  {v['bytes'] / (v['ms'] * 1e-3) / 1e9 if v['ms'] > 0 else 0:7.0f} GB/s")
print("  total profiled", tot, "ms")
