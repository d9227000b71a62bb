"""Attention forward A/B on a B200: the product kernel (variant 0, attention.cu) vs the experimental persistent kernel (variant 1,
attention_fa.cu; needs the experiment build: F5B_BUILD_TAG=fa F5B_NVCC_EXTRA=-DF5B_WITH_ATTN_FA python -m eraxvif5tts_b200.build, then
F5B_LIB=eraxvif5tts_b200/lib/libf5b200_fa.so) at two polynomial-exp2 fractions.  Correctness against an fp32 torch reference on ragged / tail
shapes, run-to-run determinism, then CUDA-event timing at the cfg-1/2/3 shapes with torch SDPA as the library comparator.
Writes gpurun_out/attn_fa_eval.json.   python tools/attn_fa_eval.py [--quick]"""
import ctypes as C
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eraxvif5tts_b200 import _lib as L, ops  # noqa: E402

L.load()
raw = C.CDLL(L.LIB_PATH)
dev = torch.device("cuda", 0)
RES = {"correctness": [], "bench": []}


def set_variant(v, poly=None):
    raw.f5b_debug_attn_variant(v)
    if poly is not None:
        raw.f5b_debug_attn_poly(poly)


def reference(qkv, B, H, n, lens):
    D = H * 64
    q, k, v = (t.float() for t in qkv.view(B, n, 3, H, 64).permute(2, 0, 3, 1, 4))
    s = q @ k.transpose(-1, -2) / 8.0
    km = None
    if lens is not None:
        km = torch.arange(n, device=dev)[None, :] < lens[:, None]
        s = s.masked_fill(~km[:, None, None, :], float("-inf"))
    o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, n, D)
    if km is not None:
        o = o * km[..., None]
    lse = torch.logsumexp(s * 1.0, dim=-1) * 1.4426950408889634  # log2 domain
    return o, lse


def run(qkv, B, H, n, lens, want_lse=False):
    D = H * 64
    out = torch.full((B * n, D), float("nan"), dtype=torch.bfloat16, device=dev)
    lse = torch.full((B, H, n), float("nan"), dtype=torch.float32, device=dev) if want_lse else None
    if want_lse:
        ops.attn_fwd_lse(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, lse, lens, 0, B, H, n)
    else:
        ops.attn_fwd(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, lens, 0, B, H, n)
    torch.cuda.synchronize()
    return out, lse


def correctness():
    g = torch.Generator().manual_seed(3)
    shapes = [(1, 2, 1, None), (2, 3, 127, [127, 5]), (1, 1, 128, None), (2, 2, 129, [129, 128]), (2, 4, 256, [256, 1]), (1, 3, 257, None),
              (3, 2, 500, [500, 260, 37]), (1, 16, 1000, None), (2, 16, 1875, [1875, 1300]), (1, 2, 4096, None), (2, 2, 4093, [4093, 1]),
              (4, 12, 1376, [1376, 1290, 700, 129])]
    ok = True
    for B, H, n, lens in shapes:
        D = H * 64
        scale = 1.0 if n < 2000 else 0.6
        qkv = (torch.randn(B * n, 3 * D, generator=g) * scale).to(dev).bfloat16()
        # a few rows with a strongly growing maximum along the keys (exercises the lazy rescale)
        qkv[:: 7, :D] *= 3.0
        lens_t = torch.tensor(lens, dtype=torch.int32, device=dev) if lens else None
        ref, ref_lse = reference(qkv, B, H, n, lens_t)
        for variant, poly in ((0, None), (1, 0), (1, 4)):
            set_variant(variant, poly)
            out, lse = run(qkv, B, H, n, lens_t, want_lse=True)
            out2, _ = run(qkv, B, H, n, lens_t)
            err = (out.float().view(B, n, D) - ref).abs().max().item()
            rel = err / ref.abs().max().item()
            valid = torch.isfinite(ref_lse)
            if lens_t is not None:
                valid = valid & (torch.arange(n, device=dev)[None, None, :] < lens_t[:, None, None])
            lse_err = (lse - ref_lse)[valid].abs().max().item()
            det = torch.equal(out, out2)
            nan = int(torch.isnan(out.float()).sum())
            good = rel <= 1e-2 and nan == 0 and det and lse_err <= 2e-2
            ok &= good
            RES["correctness"].append(dict(B=B, H=H, n=n, ragged=lens is not None, variant=variant, poly=poly, max_abs=err, rel=rel,
                                           lse_err=lse_err, deterministic=det, nan=nan, ok=good))
            print(f"[fa] B{B} H{H} n{n} {'ragged' if lens else 'full  '} variant {variant} poly {poly}: rel {rel:.2e} lse {lse_err:.2e} "
                  f"det {det} nan {nan} {'ok' if good else 'FAIL'}", flush=True)
    return ok


def _time(fn, iters):
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


def bench(quick):
    shapes = [(32, 16, 1875, False), (32, 16, 1875, True), (64, 12, 1376, False), (2, 16, 940, False), (8, 16, 7500, False)]
    if quick:
        shapes = shapes[:1]
    for B, H, n, ragged in shapes:
        D = H * 64
        qkv = torch.randn(B * n, 3 * D, device=dev).bfloat16()
        out = torch.empty(B * n, D, dtype=torch.bfloat16, device=dev)
        lens = None
        if ragged:
            lens = torch.randint(int(0.8 * n), n + 1, (B // 2,), generator=torch.Generator().manual_seed(1)).to(torch.int32).to(dev)
        fl = 4.0 * B * H * n * n * 64
        if lens is not None:
            ll = lens.repeat(2).double()
            fl = float((4.0 * H * 64 * ll * ll).sum())  # un-padded work
        row = dict(B=B, H=H, n=n, ragged=ragged)
        for variant, poly in ((0, None), (1, 0), (1, 4)):
            set_variant(variant, poly)
            ms = _time(lambda: ops.attn_fwd(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, lens, B // 2 if ragged else 0, B, H, n), 10)
            row[f"v{variant}" + (f"_poly{poly}" if poly is not None else "")] = dict(ms=ms, tflops=fl / ms / 1e9)
        if not ragged:
            q, k, v = (t.contiguous() for t in qkv.view(B, n, 3, H, 64).permute(2, 0, 3, 1, 4))
            ms2 = _time(lambda: F.scaled_dot_product_attention(q, k, v), 10)
            row["torch_sdpa"] = dict(ms=ms2, tflops=fl / ms2 / 1e9)
        RES["bench"].append(row)
        print("[fa] bench", json.dumps(row), flush=True)


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    ok = correctness()
    RES["all_correct"] = ok
    if ok or "--force-bench" in sys.argv:
        bench(quick)
    set_variant(0, 4)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "attn_fa_eval.json"), "w") as f:
        json.dump(RES, f, indent=1)
    sys.exit(0 if ok else 1)
