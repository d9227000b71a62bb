"""MelSpec / iSTFT-head microbenchmark at the cfg-2 shapes (HBM-bound kernels of the north star)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from eraxvif5tts_b200 import ops
from eraxvif5tts_b200.model import MelSpec
import gpu_diag as D
dev = "cuda"
ms = MelSpec().to(dev)
for B, frames in ((16, 563), (64, 1875)):
    wav = 0.1 * torch.randn(B, 256 * (frames - 1), device=dev)
    t = D._time(lambda: ops.melspec(wav, ms.fb, ms.fb_ranges, 100), 20)
    nb = B * frames * (1024 + 400)
    D.report(f"melspec_B{B}_T{frames}", ms=t, gbs=nb / t / 1e6)
for B, T in ((16, 1312), (64, 1875)):
    head = torch.randn(B * T, 1026, device=dev) * 0.5
    t = D._time(lambda: ops.istft_head(head, B, T), 20)
    nb = B * T * (1026 * 4 + 1024)
    D.report(f"istft_B{B}_T{T}", ms=t, gbs=nb / t / 1e6)
