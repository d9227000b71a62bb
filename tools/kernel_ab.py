"""Stand-alone A/B timings (CUDA events) used in round 2's last session:
   python tools/kernel_ab.py dual   -> FeedForward forward GEMM: BF16 epilogue + f5b_act_fwd sweep vs the BF16_DUAL epilogue (cfg-5 shape)
   python tools/kernel_ab.py conv   -> conv_pos_embed at the cfg-2 (D 1024) and cfg-3 (D 768) shapes, both layers' modes
   python tools/kernel_ab.py attn   -> attention forward at the cfg-2 / cfg-5 / cfg-3 per-layer shapes (run under F5B_LIB=... for variants)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eraxvif5tts_b200 import _lib as L, ops  # noqa: E402

dev = "cuda"
bf16 = torch.bfloat16


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def dual():
    lib = L.load()
    M, N, K = 38400, 2048, 1024
    a = (torch.randn(M, K, device=dev) * 0.5).to(bf16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(bf16)
    bias = torch.randn(N, device=dev) * 0.1
    h = torch.empty(M, N, device=dev, dtype=bf16)
    u = torch.empty_like(h)

    def two():
        ops.gemm(a, w, epi=L.EPI_BF16, act=L.ACT_NONE, bias=bias, out=h)
        L.check(lib.f5b_act_fwd(h.data_ptr(), u.data_ptr(), M * N, L.ACT_GELU_TANH, L.stream()), "act_fwd")

    def one():
        ops.gemm(a, w, epi=L.EPI_BF16_DUAL, act=L.ACT_GELU_TANH, bias=bias, out=h, out2=u)

    def plain():
        ops.gemm(a, w, epi=L.EPI_BF16, act=L.ACT_NONE, bias=bias, out=h)

    fl = 2.0 * M * N * K
    for name, fn in (("gemm only", plain), ("gemm + act_fwd sweep", two), ("dual-output gemm", one), ("gemm + act_fwd sweep", two), ("dual-output gemm", one)):
        ms = timeit(fn)
        print(f"dual: {name:24s} {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)


def conv(shapes=((32, 1875, 1024), (64, 1376, 768))):
    for (B, n, D) in shapes:
        groups, ks = 16, 31
        cpg = D // groups
        x = torch.randn(B * n, D, device=dev).to(bf16)
        w = torch.randn(D, cpg, ks, device=dev) / (cpg * ks) ** 0.5
        bias = torch.randn(D, device=dev) * 0.1
        wpk = ops.pack_convpos_weight(w, groups)
        out = torch.empty(B * n, D, dtype=bf16, device=dev)
        res = torch.zeros(B * n, D, device=dev)
        fl = 2.0 * B * n * D * cpg * ks
        ms0 = timeit(lambda: ops.convpos(x, wpk, bias, B, n, D, groups, ks, out=out))
        ms1 = timeit(lambda: ops.convpos(x, wpk, bias, B, n, D, groups, ks, resid=res))
        print(f"conv: B{B} n{n} D{D}: mish->bf16 {ms0 * 1e3:7.1f} us {fl / ms0 / 1e9:6.1f} TFLOP/s | mish+residual {ms1 * 1e3:7.1f} us {fl / ms1 / 1e9:6.1f} TFLOP/s", flush=True)


def gate():
    """gated-residual GEMMs (out-proj K = D, FF2 K = 2D) at the cfg-2 / cfg-1 / cfg-3 row counts, with a bit-level checksum of x"""
    for (M, n, N, K) in ((60000, 1875, 1024, 1024), (60000, 1875, 1024, 2048), (1880, 940, 1024, 1024), (1880, 940, 1024, 2048), (88064, 1376, 768, 768),
                         (88064, 1376, 768, 1536)):
        g = torch.Generator(device=dev).manual_seed(1)
        a = (torch.randn(M, K, device=dev, generator=g) * 0.5).to(bf16)
        w = (torch.randn(N, K, device=dev, generator=g) * 0.03).to(bf16)
        bias = torch.randn(N, device=dev, generator=g) * 0.1
        B = M // n
        gate_t = torch.randn(B, N, device=dev, generator=g)
        x = torch.zeros(M, N, device=dev)
        fn = lambda: ops.gemm(a, w, epi=L.EPI_GATE_RESID, bias=bias, out=x, rows_per_batch=n, gate=gate_t, gate_bstride=N)
        fn()
        torch.cuda.synchronize()
        ref = (a[:256].float() @ w.float().t() + bias) * gate_t[0]
        err = (x[:256] - ref).abs().max().item()
        chk = x.double().sum().item()
        ms = timeit(fn)
        print(f"gate: M{M} N{N} K{K}: {ms * 1e3:8.1f} us {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s  err(first call, 256 rows) {err:.2e} sum {chk:.6e}  (lib {os.environ.get('F5B_LIB', 'default')})", flush=True)


def sweeps():
    """the memory-bound sweeps of the training step at cfg-5's shape (32 x 1200 rows of 1024), algorithmic bytes / time"""
    lib = L.load()
    B, n, D = 32, 1200, 1024
    x = torch.randn(B, n, D, device=dev)
    z = torch.randn(B, n, D, device=dev).to(bf16)
    mod = torch.randn(B, 6 * D, device=dev) * 0.1
    lens = torch.full((B,), n, dtype=torch.int32, device=dev)
    xo = torch.empty_like(x)
    hb = torch.empty(B, n, D, dtype=bf16, device=dev)
    S = L.stream
    ms = timeit(lambda: L.check(lib.f5b_gate_add_ln_modulate(x.data_ptr(), z.data_ptr(), mod.data_ptr(), 6 * D, lens.data_ptr(), xo.data_ptr(),
                                                             mod.data_ptr() + 4 * D, mod.data_ptr() + 8 * D, 6 * D, hb.data_ptr(), B, n, D, 1e-6, S()), "gal"))
    print(f"sweep: gate_add_ln_modulate {ms * 1e3:7.1f} us  {12.0 * B * n * D / ms / 1e6:7.1f} GB/s", flush=True)
    dy = torch.randn(B, n, D, device=dev).to(bf16)
    dx = torch.zeros(B, n, D, device=dev)
    dmod = torch.zeros(B, 6 * D, device=dev)
    ms = timeit(lambda: L.check(lib.f5b_ln_modulate_bwd(dy.data_ptr(), x.data_ptr(), mod.data_ptr() + 4 * D, 6 * D, dx.data_ptr(), 1, dmod.data_ptr() + 4 * D,
                                                        dmod.data_ptr(), B, n, D, 1e-6, S()), "lnb"))
    print(f"sweep: ln_modulate_bwd      {ms * 1e3:7.1f} us  {14.0 * B * n * D / ms / 1e6:7.1f} GB/s", flush=True)
    dz = torch.empty(B, n, D, dtype=bf16, device=dev)
    dgate, dbias = torch.zeros(B, 6 * D, device=dev), torch.zeros(D, device=dev)
    ms = timeit(lambda: L.check(lib.f5b_gate_bwd(dx.data_ptr(), z.data_ptr(), mod.data_ptr(), 6 * D, lens.data_ptr(), dz.data_ptr(), dgate.data_ptr(),
                                                 dbias.data_ptr(), B, n, D, S()), "gb"))
    print(f"sweep: gate_bwd             {ms * 1e3:7.1f} us  {8.0 * B * n * D / ms / 1e6:7.1f} GB/s", flush=True)


def attn():
    for (B, H, n) in ((32, 16, 1875), (32, 16, 1200), (64, 12, 1376), (2, 16, 940)):
        D = H * 64
        qkv = torch.randn(B * n, 3 * D, device=dev).to(bf16)
        out = torch.empty(B * n, D, dtype=bf16, device=dev)
        ms = timeit(lambda: ops.attn_fwd(qkv, qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, None, 0, B, H, n), iters=10)
        print(f"attn: B{B} H{H} n{n}: {ms * 1e3:8.1f} us {4.0 * B * H * n * n * 64 / ms / 1e9:7.1f} TFLOP/s  (lib {os.environ.get('F5B_LIB', 'default')})", flush=True)


if __name__ == "__main__":
    for what in sys.argv[1:] or ["dual", "conv", "attn"]:
        {"dual": dual, "conv": conv, "conv3": lambda: conv(((64, 1376, 768),)), "attn": attn, "gate": gate, "sweeps": sweeps}[what]()
