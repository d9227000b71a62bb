"""Kernel-by-kernel diagnostics on a B200 (development aid; the asserted versions live in tests/).  Each check runs in its own
try block so one failure does not hide the rest; results go to stdout and gpurun_out/diag.json."""
from __future__ import annotations

import json
import math
import os
import sys
import time
import traceback

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from eraxvif5tts_b200 import _lib as L  # noqa: E402
from eraxvif5tts_b200 import ops  # noqa: E402

dev = "cuda"
bf16, f32 = torch.bfloat16, torch.float32
RES = {}


def report(name, **kw):
    RES[name] = kw
    print(f"[diag] {name}: " + ", ".join(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}" for k, v in kw.items()), flush=True)


def run(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        traceback.print_exc()
        RES[name] = {"error": repr(e)}
        print(f"[diag] {name}: ERROR {e!r}", flush=True)
        try:
            torch.cuda.synchronize()
        except Exception as e2:  # noqa: BLE001
            print("[diag] CUDA context is dead:", e2, flush=True)
            finish()
            sys.exit(3)


def finish():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
        json.dump(RES, f, indent=1)


def relerr(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


def gemm_case(M, N, K, epi="bf16", act=L.ACT_NONE):
    g = torch.Generator(device=dev).manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, device=dev, generator=g) * 0.5).to(bf16)
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).to(bf16)
    bias = torch.randn(N, device=dev, generator=g) * 0.1
    ref = a.float() @ w.float().t() + bias
    if act == L.ACT_GELU_TANH:
        ref = F.gelu(ref, approximate="tanh")
    elif act == L.ACT_GELU_ERF:
        ref = F.gelu(ref)
    elif act == L.ACT_SILU:
        ref = F.silu(ref)
    if epi == "bf16":
        out = torch.full((M, N), float("nan"), dtype=bf16, device=dev)
        ops.gemm(a, w, epi=L.EPI_BF16, act=act, bias=bias, out=out)
    else:
        out = torch.full((M, N), float("nan"), dtype=f32, device=dev)
        ops.gemm(a, w, epi=L.EPI_F32, act=act, bias=bias, out=out)
    torch.cuda.synchronize()
    report(f"gemm_{epi}_{M}x{N}x{K}_act{act}", rel=relerr(out, ref), nan=int(torch.isnan(out.float()).sum()))


def gemm_gate_case(B, n, N, K):
    M = B * n
    g = torch.Generator(device=dev).manual_seed(5)
    a = (torch.randn(M, K, device=dev, generator=g) * 0.5).to(bf16)
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).to(bf16)
    bias = torch.randn(N, device=dev, generator=g) * 0.1
    x = torch.randn(M, N, device=dev, generator=g)
    gate = torch.randn(B, N, device=dev, generator=g)
    lens = torch.tensor([n - 3 * i for i in range(B)], dtype=torch.int32, device=dev)
    y = (a.float() @ w.float().t() + bias).view(B, n, N) * gate[:, None, :]
    mask = torch.arange(n, device=dev)[None, :] < lens[:, None]
    ref = x.view(B, n, N) + y * mask[..., None]
    out = x.clone()
    ops.gemm(a, w, epi=L.EPI_GATE_RESID, bias=bias, out=out, rows_per_batch=n, gate=gate, gate_bstride=N, lens=lens)
    report(f"gemm_gate_{B}x{n}x{N}x{K}", rel=relerr(out, ref.view(M, N)))


def rope_ref(t, pos_freqs):
    tf = t.float()
    x = tf.reshape(*tf.shape[:-1], -1, 2)
    rot = torch.stack((-x[..., 1], x[..., 0]), dim=-1).flatten(-2)
    return tf * pos_freqs.cos() + rot * pos_freqs.sin()


def qkv_attn_case(B, H, n, lens_list=None, rope_heads=1):
    D = H * 64
    g = torch.Generator(device=dev).manual_seed(11)
    h = (torch.randn(B * n, D, device=dev, generator=g)).to(bf16)
    w = (torch.randn(3 * D, D, device=dev, generator=g) / math.sqrt(D)).to(bf16)
    bias = torch.randn(3 * D, device=dev, generator=g) * 0.1
    inv = 1.0 / (10000 ** (torch.arange(0, 64, 2, device=dev).float() / 64))
    fr = torch.outer(torch.arange(n, device=dev).float(), inv)
    rope = torch.stack((fr.cos(), fr.sin()), dim=-1).contiguous()  # [n,32,2]
    qkv_out = torch.full((B * n, 3 * D), float("nan"), dtype=bf16, device=dev)  # token-major q | k | v
    ops.gemm(h, w, epi=L.EPI_QKV_ROPE, bias=bias, out=qkv_out, rows_per_batch=n, rope=rope, rope_heads=rope_heads, heads=H)
    torch.cuda.synchronize()
    qkv = (h.float() @ w.float().t() + bias).view(B, n, 3, H, 64).permute(2, 0, 3, 1, 4)  # [3,B,H,n,64]
    fr2 = torch.stack((fr, fr), dim=-1).flatten(-2)  # [n,64]
    qr, kr, vr = qkv[0].clone(), qkv[1].clone(), qkv[2]
    if rope_heads > 0:
        qr[:, :rope_heads] = rope_ref(qr[:, :rope_heads], fr2)
        kr[:, :rope_heads] = rope_ref(kr[:, :rope_heads], fr2)
    got = qkv_out.view(B, n, 3, H, 64).permute(2, 0, 3, 1, 4)
    report(f"qkv_rope_B{B}H{H}n{n}", q_rel=relerr(got[0], qr), k_rel=relerr(got[1], kr), vt_rel=relerr(got[2], vr))
    # attention on the kernel's own (bf16) q, k, v
    lens = None
    if lens_list is not None:
        lens = torch.tensor(lens_list, dtype=torch.int32, device=dev)
    out = torch.full((B * n, D), float("nan"), dtype=bf16, device=dev)
    ops.attn_fwd(qkv_out, qkv_out[:, D:], qkv_out[:, 2 * D:], 3 * D, out, lens, 0, B, H, n)
    torch.cuda.synchronize()
    qf, kf, vf = got[0].float(), got[1].float(), got[2].float()
    s = qf @ kf.transpose(-1, -2) / 8.0
    if lens is not None:
        km = torch.arange(n, device=dev)[None, :] < lens[:, None]
        s = s.masked_fill(~km[:, None, None, :], float("-inf"))
    o = torch.softmax(s, dim=-1) @ vf
    o = o.transpose(1, 2).reshape(B, n, D)
    if lens is not None:
        o = o * km[..., None]
    report(f"attn_B{B}H{H}n{n}_{'ragged' if lens_list else 'full'}", rel=relerr(out.view(B, n, D), o),
           nan=int(torch.isnan(out.float()).sum()))


def convpos_case(B, n, D, groups=16, ks=31):
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(B, n, D, device=dev, generator=g).to(bf16)
    cpg = D // groups
    w = (torch.randn(D, cpg, ks, device=dev, generator=g) / math.sqrt(cpg * ks)).to(bf16).float()
    bias = torch.randn(D, device=dev, generator=g) * 0.1
    wpk = ops.pack_convpos_weight(w, groups)
    out = torch.full((B * n, D), float("nan"), dtype=bf16, device=dev)
    ops.convpos(x.view(B * n, D), wpk, bias, B, n, D, groups, ks, out=out)
    ref = F.mish(F.conv1d(x.float().transpose(1, 2), w, bias, padding=ks // 2, groups=groups)).transpose(1, 2)
    torch.cuda.synchronize()
    report(f"convpos_B{B}n{n}D{D}", rel=relerr(out.view(B, n, D), ref), nan=int(torch.isnan(out.float()).sum()))
    res = torch.randn(B * n, D, device=dev, generator=g)
    r0 = res.clone()
    ops.convpos(x.view(B * n, D), wpk, bias, B, n, D, groups, ks, resid=res)
    report(f"convpos_resid_B{B}n{n}D{D}", rel=relerr(res.view(B, n, D), r0.view(B, n, D) + ref))


def small_kernels():
    g = torch.Generator(device=dev).manual_seed(1)
    B, n, D = 3, 77, 1024
    x = torch.randn(B * n, D, device=dev, generator=g) * 2 + 0.3
    mod = torch.randn(B, 6 * D, device=dev, generator=g) * 0.2
    out = torch.empty(B * n, D, dtype=bf16, device=dev)
    lib = L.load()
    L.check(lib.f5b_ln_modulate(x.data_ptr(), mod.data_ptr() + 4 * D, mod.data_ptr(), 6 * D, 0, out.data_ptr(), B * n, n, D, 1e-6,
                                L.stream()), "f5b_ln_modulate")
    ref = F.layer_norm(x, (D,), eps=1e-6).view(B, n, D) * (1 + mod[:, None, D:2 * D]) + mod[:, None, :D]
    report("ln_modulate", rel=relerr(out.view(B, n, D), ref))
    # dwconv + LN
    C_ = 512
    xt = torch.randn(B, n, C_, device=dev, generator=g)
    w = torch.randn(C_, 1, 7, device=dev, generator=g) / math.sqrt(7)
    b = torch.randn(C_, device=dev, generator=g) * 0.1
    lw = 1 + 0.1 * torch.randn(C_, device=dev, generator=g)
    lb = 0.1 * torch.randn(C_, device=dev, generator=g)
    o = ops.dwconv7_ln(xt.view(B * n, C_), w.view(C_, 7).contiguous(), b, lw, lb, B, n)
    r = F.layer_norm(F.conv1d(xt.transpose(1, 2), w, b, padding=3, groups=C_).transpose(1, 2), (C_,), lw, lb, 1e-6)
    report("dwconv7_ln", rel=relerr(o.view(B, n, C_), r))
    # GRN
    hh = torch.randn(B, n, 1024, device=dev, generator=g).to(bf16)
    gam = torch.randn(1024, device=dev, generator=g) * 0.2
    bet = torch.randn(1024, device=dev, generator=g) * 0.02
    o = ops.grn(hh.view(B * n, 1024), gam, bet, B, n)
    hf = hh.float()
    gx = torch.norm(hf, p=2, dim=1, keepdim=True)
    nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
    report("grn", rel=relerr(o.view(B, n, 1024), gam * (hf * nx) + bet + hf))
    # cfg euler
    y = torch.randn(B * n, 100, device=dev, generator=g)
    pc = torch.randn(B * n, 100, device=dev, generator=g)
    pu = torch.randn(B * n, 100, device=dev, generator=g)
    yb = torch.full((B * n, 128), float("nan"), dtype=bf16, device=dev)
    y0 = y.clone()
    ops.cfg_euler(y, pc, pu, 2.0, 0.03, yb)
    ref = y0 + 0.03 * (pc + (pc - pu) * 2.0)
    report("cfg_euler", rel=relerr(y, ref), bf_rel=relerr(yb[:, :100], ref), pad=float(yb[:, 100:].float().abs().max()))
    # time sinus
    t = torch.tensor([0.0, 0.37, 1.0], device=dev)
    o = ops.time_sinus(t)
    k = math.log(10000) / 127
    fr = torch.exp(torch.arange(128, device=dev).float() * -k)
    e = 1000.0 * t[:, None] * fr[None]
    report("time_sinus", abs=float((o.float() - torch.cat((e.sin(), e.cos()), -1)).abs().max()))


def spectral():
    import torchaudio
    g = torch.Generator(device=dev).manual_seed(7)
    wav = 0.1 * torch.randn(2, 256 * 37 + 19, device=dev, generator=g)
    fb = torchaudio.functional.melscale_fbanks(513, 0.0, 12000.0, 100, 24000, norm=None, mel_scale="htk").to(dev)
    nz = fb > 0
    f0 = torch.where(nz.any(0), nz.float().argmax(0), torch.zeros(100, device=dev, dtype=torch.long))
    f1 = torch.where(nz.any(0), 513 - nz.flip(0).float().argmax(0), torch.zeros(100, device=dev, dtype=torch.long))
    ranges = torch.stack((f0, f1), -1).to(torch.int32).contiguous()
    mel = ops.melspec(wav, fb.contiguous(), ranges, 100)
    x = F.pad(wav.unsqueeze(1), (512, 512), mode="reflect").squeeze(1)
    fr = x.unfold(-1, 1024, 256)
    spec = torch.fft.rfft(fr * torch.hann_window(1024, device=dev), dim=-1).abs()
    ref = (spec @ fb).clamp(min=1e-5).log()
    report("melspec", abs=float((mel - ref).abs().max()), mean_abs=float((mel - ref).abs().mean()))
    # istft
    T = 21
    head = torch.randn(2 * T, 1026, device=dev, generator=g)
    head[:, :513] = head[:, :513] * 0.5 - 1.0
    wavo = ops.istft_head(head, 2, T)
    hm = head.view(2, T, 1026).transpose(1, 2)
    mag, ph = hm[:, :513], hm[:, 513:]
    mag = torch.clip(torch.exp(mag), max=1e2)
    S = mag * (torch.cos(ph) + 1j * torch.sin(ph))
    ref = torch.istft(S, 1024, 256, 1024, torch.hann_window(1024, device=dev), center=True)
    report("istft", abs=float((wavo - ref).abs().max()), ref_max=float(ref.abs().max()))


def bench_gemm(M, N, K, iters=20):
    a = torch.randn(M, K, device=dev).to(bf16)
    w = torch.randn(N, K, device=dev).to(bf16)
    out = torch.empty(M, N, dtype=bf16, device=dev)
    for _ in range(3):
        ops.gemm(a, w, epi=L.EPI_BF16, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, w, epi=L.EPI_BF16, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    for _ in range(3):
        torch.matmul(a, w.t(), out=out)
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.t(), out=out)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    report(f"bench_gemm_{M}x{N}x{K}", ms=ms, tflops=2 * M * N * K / ms / 1e9, cublas_ms=ms2, cublas_tflops=2 * M * N * K / ms2 / 1e9)


def _time(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_epilogues(B=32, n=1875, D=1024):
    """every GEMM flavour of one DiT block at the cfg-2 in-situ shape (2B x n rows)"""
    M, H, F_ = B * n, D // 64, 2 * D
    n_pad = (n + 7) // 8 * 8
    h = torch.randn(M, D, device=dev).to(bf16)
    fb = torch.randn(M, F_, device=dev).to(bf16)
    wqkv = (torch.randn(3 * D, D, device=dev) / 32).to(bf16)
    wo = (torch.randn(D, D, device=dev) / 32).to(bf16)
    w1 = (torch.randn(F_, D, device=dev) / 32).to(bf16)
    w2 = (torch.randn(D, F_, device=dev) / 45).to(bf16)
    bq, bo, b1 = torch.randn(3 * D, device=dev), torch.randn(D, device=dev), torch.randn(F_, device=dev)
    rope = torch.randn(n, 32, 2, device=dev)
    x = torch.randn(M, D, device=dev)
    gate = torch.randn(D, device=dev) * 0.1
    lens = torch.full((B // 2,), n, dtype=torch.int32, device=dev)
    ob = torch.empty(M, F_, dtype=bf16, device=dev)
    o3 = torch.empty(M, 3 * D, dtype=bf16, device=dev)
    o1 = torch.empty(M, D, dtype=bf16, device=dev)
    cases = {
        "qkv_rope": (lambda: ops.gemm(h, wqkv, epi=L.EPI_QKV_ROPE, bias=bq, out=o3, rows_per_batch=n, rope=rope, rope_heads=1, heads=H),
                     2.0 * M * 3 * D * D),
        "qkv_plain_bf16": (lambda: ops.gemm(h, wqkv, epi=L.EPI_BF16, bias=bq, out=o3), 2.0 * M * 3 * D * D),
        "out_gate_resid": (lambda: ops.gemm(h, wo, epi=L.EPI_GATE_RESID, bias=bo, out=x, rows_per_batch=n, gate=gate, gate_bstride=0, lens=lens,
                                             batch_mod=B // 2), 2.0 * M * D * D),
        "out_plain_bf16": (lambda: ops.gemm(h, wo, epi=L.EPI_BF16, bias=bo, out=o1), 2.0 * M * D * D),
        "ff1_gelu": (lambda: ops.gemm(h, w1, epi=L.EPI_BF16, act=L.ACT_GELU_TANH, bias=b1, out=ob), 2.0 * M * F_ * D),
        "ff1_plain": (lambda: ops.gemm(h, w1, epi=L.EPI_BF16, bias=b1, out=ob), 2.0 * M * F_ * D),
        "ff2_gate_resid": (lambda: ops.gemm(fb, w2, epi=L.EPI_GATE_RESID, bias=bo, out=x, rows_per_batch=n, gate=gate, gate_bstride=0),
                           2.0 * M * D * F_),
        "ff2_plain_bf16": (lambda: ops.gemm(fb, w2, epi=L.EPI_BF16, bias=bo, out=o1), 2.0 * M * D * F_),
        "cublas_qkv": (lambda: torch.matmul(h, wqkv.t(), out=o3), 2.0 * M * 3 * D * D),
        "cublas_ff2": (lambda: torch.matmul(fb, w2.t(), out=o1), 2.0 * M * D * F_),
    }
    for name, (fn, fl) in cases.items():
        ms = _time(fn)
        report(f"epi_{name}", ms=ms, tflops=fl / ms / 1e9)
    xx = torch.randn(M, D, device=dev)
    mod = torch.randn(6 * D, device=dev)
    hb = torch.empty(M, D, dtype=bf16, device=dev)
    lib = L.load()
    ms = _time(lambda: L.check(lib.f5b_ln_modulate(xx.data_ptr(), mod.data_ptr() + 4 * D, mod.data_ptr(), 0, 0, hb.data_ptr(), M, n, D, 1e-6,
                                                   L.stream()), "ln"))
    report("epi_ln_modulate", ms=ms, gbs=M * D * 6 / ms / 1e6)


def bench_attn(B, H, n, iters=10):
    D = H * 64
    qkv = torch.randn(B * n, 3 * D, device=dev).to(bf16)
    out = torch.empty(B * n, D, dtype=bf16, device=dev)
    fn = lambda: ops.attn_fwd(qkv, qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, None, 0, B, H, n)
    ms = _time(fn, iters)
    fl = 4.0 * B * H * n * n * 64
    q, k, v = (t.contiguous() for t in qkv.view(B, n, 3, H, 64).permute(2, 0, 3, 1, 4))
    ms2 = _time(lambda: F.scaled_dot_product_attention(q, k, v), iters)
    report(f"bench_attn_B{B}H{H}n{n}", ms=ms, tflops=fl / ms / 1e9, sdpa_ms=ms2, sdpa_tflops=fl / ms2 / 1e9)


def main():
    t0 = time.time()
    L.load()
    print("[diag] device:", torch.cuda.get_device_name(0), flush=True)
    run("gemm1", lambda: gemm_case(128, 128, 64))
    run("gemm2", lambda: gemm_case(128, 128, 256))
    run("gemm3", lambda: gemm_case(300, 200, 136, "f32"))
    run("gemm4", lambda: gemm_case(1000, 1024, 1024))
    run("gemm5", lambda: gemm_case(30000, 2048, 1024, "bf16", L.ACT_GELU_TANH))
    run("gemm6", lambda: gemm_case(4100, 100, 1024, "f32"))
    run("gemm7", lambda: gemm_case(33, 1024, 256, "bf16", L.ACT_SILU))
    run("gemm8", lambda: gemm_case(700, 512, 1024, "bf16", L.ACT_GELU_ERF))
    run("gemm_gate", lambda: gemm_gate_case(3, 200, 1024, 2048))
    run("qkv_attn_small", lambda: qkv_attn_case(1, 2, 128))
    run("qkv_attn_mid", lambda: qkv_attn_case(2, 16, 300, [300, 211]))
    run("qkv_attn_big", lambda: qkv_attn_case(2, 16, 1875, None, rope_heads=16))
    run("convpos_1024", lambda: convpos_case(2, 300, 1024))
    run("convpos_768", lambda: convpos_case(2, 200, 768))
    run("convpos_128", lambda: convpos_case(2, 96, 128))
    run("small", small_kernels)
    run("spectral", spectral)
    if "--epi" in sys.argv:
        run("bench_epilogues", bench_epilogues)
        run("bench_attn32", lambda: bench_attn(32, 16, 1875))
        finish()
        return
    run("bench_gemm_qkv", lambda: bench_gemm(30000, 3072, 1024))
    run("bench_gemm_ff2", lambda: bench_gemm(30000, 1024, 2048))
    run("bench_gemm_big", lambda: bench_gemm(8192, 8192, 8192, 5))
    run("bench_attn", lambda: bench_attn(16, 16, 1875))
    print(f"[diag] done in {time.time() - t0:.1f}s", flush=True)
    finish()


if __name__ == "__main__":
    main()
