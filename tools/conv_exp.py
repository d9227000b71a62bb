import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import gpu_diag as D
from eraxvif5tts_b200 import _lib as L, ops
lib = L.load()
for halo, bo in ((0, 0), (1, 0)):
    lib.f5b_debug_convpos_mode(halo)
    print("mode halo", halo, "base_offset", bo, flush=True)
    for (B, n, Dm) in ((2, 300, 1024), (2, 200, 768), (2, 96, 128), (1, 31, 1024), (2, 1875, 1024), (1, 513, 1024), (3, 640, 768)):
        D.convpos_case(B, n, Dm)
    # timing at the cfg-2 in-situ shape
    B, n, Dm = 32, 1875, 1024
    x = torch.randn(B * n, Dm, device="cuda").to(torch.bfloat16)
    w = torch.randn(Dm, 64, 31, device="cuda") * 0.02
    wpk = ops.pack_convpos_weight(w, 16)
    bias = torch.zeros(Dm, device="cuda")
    out = torch.empty(B * n, Dm, dtype=torch.bfloat16, device="cuda")
    ms = D._time(lambda: ops.convpos(x, wpk, bias, B, n, Dm, 16, 31, out=out), 10)
    D.report(f"convpos_time_halo{halo}_bo{bo}", ms=ms, tflops=2.0 * B * n * Dm * 64 * 31 / ms / 1e9)
