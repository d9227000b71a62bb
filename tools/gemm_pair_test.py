"""A/B of the GEMM engine's CTA-pair mode (tcgen05 cta_group::2) against the default on the DiT linear shapes: equality + timing."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eraxvif5tts_b200 import _lib as L, ops  # noqa: E402
L.load()
raw = C.CDLL(L.LIB_PATH)
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
shapes = [(60000, 2048, 1024, L.ACT_GELU_TANH, "FF1"), (60000, 1024, 2048, L.ACT_NONE, "FF2"), (60000, 3072, 1024, L.ACT_NONE, "QKV (bf16 epilogue)"),
          (60000, 1024, 1024, L.ACT_NONE, "out"), (38400, 2048, 1024, L.ACT_NONE, "train FF1"), (1880, 2048, 1024, L.ACT_NONE, "small M"),
          (333, 512, 200, L.ACT_NONE, "ragged")]
for M, N, K, act, name in shapes:
    a = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    outs, times = [], []
    for mode in (0, 1):
        raw.f5b_debug_gemm_pair_mode(mode)
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        ops.gemm(a, w, epi=L.EPI_BF16, act=act, bias=bias, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.gemm(a, w, epi=L.EPI_BF16, act=act, bias=bias, out=out)
        e1.record()
        torch.cuda.synchronize()
        outs.append(out.float())
        times.append(e0.elapsed_time(e1) / 20)
    raw.f5b_debug_gemm_pair_mode(0)
    ref = torch.nn.functional.linear(a.float(), w.float(), bias)
    if act == L.ACT_GELU_TANH:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    err = [(o - ref).abs().max().item() / ref.abs().max().item() for o in outs]
    fl = 2.0 * M * N * K
    print(f"{name:22s} M{M} N{N} K{K}: default {times[0]*1e3:7.1f} us {fl/times[0]/1e9:6.0f} TF/s | pair {times[1]*1e3:7.1f} us {fl/times[1]/1e9:6.0f} TF/s | "
          f"rel err {err[0]:.2e} / {err[1]:.2e} | identical {torch.equal(outs[0], outs[1])}")
