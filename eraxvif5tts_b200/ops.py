"""Thin torch-tensor wrappers over the C ABI (one function per exported kernel entry point).  Tensors provide device memory and
the current CUDA stream only; all math runs in libf5b200.so.  Every wrapper validates dtype / device / contiguity and raises."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

bf16, f32 = torch.bfloat16, torch.float32


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise L.F5bError(f"{name}: tensor must live on a CUDA device (no CPU fallback)")
    if t.dtype != dtype:
        raise L.F5bError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise L.F5bError(f"{name}: tensor must be contiguous")


def gemm(a, w, *, epi, act=L.ACT_NONE, bias=None, out=None, out2=None, out3=None, addsrc=None, rows_per_batch=0, gate=None,
         gate_bstride=0, lens=None, batch_mod=0, rope=None, rope_heads=0, heads=0, tf32=False, M=None, N=None, K=None, lda=None,
         ldw=None, ldc=None, ldc2=0):
    """out = epilogue(a[M,K] @ w[N,K]^T).  a, w bf16 2-D (row pitch = shape[1] unless lda/ldw given); with tf32=True a, w are
    f32 tensors whose values were rounded to tf32 (ops.round_tf32) and the BF16 / QKV_ROPE epilogues write f32."""
    lib = L.load()
    _chk(a, f32 if tf32 else bf16, "a"); _chk(w, f32 if tf32 else bf16, "w"); _chk(bias, f32, "bias"); _chk(addsrc, f32, "addsrc"); _chk(gate, f32, "gate")
    _chk(lens, torch.int32, "lens"); _chk(rope, f32, "rope")
    g = L.GemmArgs()
    g.M = M if M is not None else a.shape[0]
    g.K = K if K is not None else a.shape[1]
    g.N = N if N is not None else w.shape[0]
    g.epi, g.act = epi, act
    g.bias = L.ptr(bias)
    g.out = L.ptr(out)
    g.ldc = ldc if ldc is not None else (out.shape[-1] if out is not None and out.dim() == 2 else 0)
    g.out2 = L.ptr(out2)
    g.ldc2 = ldc2 if ldc2 else (out2.shape[-1] if out2 is not None and out2.dim() == 2 else 0)
    g.out3 = L.ptr(out3)
    g.addsrc = L.ptr(addsrc)
    g.ld_add = addsrc.shape[-1] if addsrc is not None else 0
    g.rows_per_batch = rows_per_batch
    g.gate = L.ptr(gate)
    g.gate_bstride = gate_bstride
    g.lens = L.ptr(lens)
    g.batch_mod = batch_mod
    g.rope = L.ptr(rope)
    g.rope_heads, g.heads, g.tf32 = rope_heads, heads, int(bool(tf32))
    L.check(lib.f5b_gemm(a.data_ptr(), lda or a.stride(0), w.data_ptr(), ldw or w.stride(0), C.byref(g), L.stream()), "f5b_gemm")
    return out


def ln_modulate(x, scale, shift, mod_bstride, batch_mod, rows_per_batch, eps=1e-6, out=None):
    lib = L.load()
    _chk(x, f32, "x"); _chk(scale, f32, "scale"); _chk(shift, f32, "shift")
    rows, D = x.shape
    if out is None:
        out = torch.empty(rows, D, dtype=bf16, device=x.device)
    L.check(lib.f5b_ln_modulate(x.data_ptr(), L.ptr(scale), L.ptr(shift), mod_bstride, batch_mod, out.data_ptr(), rows,
                                rows_per_batch, D, eps, L.stream()), "f5b_ln_modulate")
    return out


def ln_affine(x, w, b, eps=1e-6, out_f32=None, out_bf16=None):
    lib = L.load()
    _chk(x, f32, "x"); _chk(w, f32, "w"); _chk(b, f32, "b")
    rows, D = x.shape
    L.check(lib.f5b_ln_affine(x.data_ptr(), w.data_ptr(), b.data_ptr(), L.ptr(out_f32), L.ptr(out_bf16), rows, D, eps, L.stream()),
            "f5b_ln_affine")


def attn_fwd(q, k, v, ld, out, lens, lens_mod, B, H, n, scale=0.125):
    """q, k, v: bf16 views into token-major [B*n, ld] matrices (column 0 of each = head 0)"""
    lib = L.load()
    for t, nm in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        if not t.is_cuda or t.dtype != bf16:
            raise L.F5bError(f"{nm}: expected a CUDA bf16 tensor")
    _chk(lens, torch.int32, "lens")
    L.check(lib.f5b_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), ld, out.data_ptr(), L.ptr(lens), lens_mod, B, H, n, scale,
                             L.stream()), "f5b_attn_fwd")
    return out


def pack_convpos_weight(w, groups):
    lib = L.load()
    _chk(w, f32, "w")
    D, cpg, ks = w.shape
    n = lib.f5b_convpos_packed_elems(D, groups, ks)
    out = torch.empty(n, dtype=bf16, device=w.device)
    L.check(lib.f5b_pack_convpos_weight(w.data_ptr(), out.data_ptr(), D, groups, ks, L.stream()), "f5b_pack_convpos_weight")
    return out


def convpos(x, wpk, bias, B, n, D, groups, ksize, out=None, resid=None):
    lib = L.load()
    _chk(x, bf16, "x"); _chk(wpk, bf16, "wpk"); _chk(bias, f32, "bias"); _chk(out, bf16, "out"); _chk(resid, f32, "resid")
    mode = 0 if resid is None else 1
    L.check(lib.f5b_convpos(x.data_ptr(), wpk.data_ptr(), bias.data_ptr(), L.ptr(out), L.ptr(resid), B, n, D, groups, ksize, mode,
                            L.stream()), "f5b_convpos")


def dwconv7_ln(x, w, b, ln_w, ln_b, B, n, eps=1e-6):
    lib = L.load()
    for t, nm in ((x, "x"), (w, "w"), (b, "b"), (ln_w, "ln_w"), (ln_b, "ln_b")):
        _chk(t, f32, nm)
    Cc = x.shape[-1]
    out = torch.empty(B * n, Cc, dtype=bf16, device=x.device)
    L.check(lib.f5b_dwconv7_ln(x.data_ptr(), w.data_ptr(), b.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), out.data_ptr(), B, n, Cc,
                               eps, L.stream()), "f5b_dwconv7_ln")
    return out


def grn(h, gamma, beta, B, n):
    lib = L.load()
    _chk(h, bf16, "h"); _chk(gamma, f32, "gamma"); _chk(beta, f32, "beta")
    Cc = h.shape[-1]
    out = torch.empty_like(h)
    ws = torch.empty(B * Cc, dtype=f32, device=h.device)
    L.check(lib.f5b_grn(h.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), ws.data_ptr(), B, n, Cc, L.stream()), "f5b_grn")
    return out


def cfg_euler(y, pc, pu, cfg, dt, y_bf16=None, vel_out=None):
    lib = L.load()
    _chk(y, f32, "y"); _chk(pc, f32, "pc"); _chk(pu, f32, "pu"); _chk(y_bf16, bf16, "y_bf16"); _chk(vel_out, f32, "vel_out")
    Cc = y.shape[-1]
    rows = y.numel() // Cc
    ld = y_bf16.shape[-1] if y_bf16 is not None else Cc
    L.check(lib.f5b_cfg_euler(y.data_ptr(), pc.data_ptr(), L.ptr(pu), float(cfg), float(dt), L.ptr(y_bf16), ld, L.ptr(vel_out), rows, Cc,
                              L.stream()), "f5b_cfg_euler")


def pack_bf16(x, out, cols, width):
    lib = L.load()
    _chk(x, f32, "x"); _chk(out, bf16, "out")
    rows = out.shape[0]
    L.check(lib.f5b_pack_bf16(L.ptr(x), x.shape[-1] if x is not None else 0, out.data_ptr(), out.shape[-1], rows, cols, width,
                              L.stream()), "f5b_pack_bf16")


# ---- tf32 operand mode (include/f5b200.h "tf32 operand mode") ---------------------------------------------------------
def round_tf32(t):
    """fp32 -> nearest tf32 value kept as fp32 (what every tf32-mode operand must hold; = PTX cvt.rna.tf32.f32)"""
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def pack_tf32(x, out, cols, width):
    lib = L.load()
    _chk(x, f32, "x"); _chk(out, f32, "out")
    rows = out.shape[0]
    L.check(lib.f5b_pack_tf32(L.ptr(x), x.shape[-1] if x is not None else 0, out.data_ptr(), out.shape[-1], rows, cols, width,
                              L.stream()), "f5b_pack_tf32")


def ln_modulate_tf32(x, scale, shift, mod_bstride, batch_mod, rows_per_batch, eps=1e-6, out=None):
    lib = L.load()
    _chk(x, f32, "x"); _chk(scale, f32, "scale"); _chk(shift, f32, "shift")
    rows, D = x.shape
    if out is None:
        out = torch.empty(rows, D, dtype=f32, device=x.device)
    L.check(lib.f5b_ln_modulate_tf32(x.data_ptr(), L.ptr(scale), L.ptr(shift), mod_bstride, batch_mod, out.data_ptr(), rows,
                                     rows_per_batch, D, eps, L.stream()), "f5b_ln_modulate_tf32")
    return out


def attn_fwd_tf32(q, k, v, ld, out, lens, lens_mod, B, H, n, scale=0.125):
    lib = L.load()
    for t, nm in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        if t.dtype != f32 or not t.is_cuda:
            raise L.F5bError(f"attn_fwd_tf32: {nm} must be a CUDA f32 tensor")
    _chk(lens, torch.int32, "lens")
    vt = torch.empty(lib.f5b_attn_tf32_ws_floats(B, H, n), dtype=f32, device=q.device)
    L.check(lib.f5b_attn_fwd_tf32(q.data_ptr(), k.data_ptr(), v.data_ptr(), ld, out.data_ptr(), vt.data_ptr(), L.ptr(lens), lens_mod, B, H,
                                  n, scale, L.stream()), "f5b_attn_fwd_tf32")
    return out


def pack_convpos_weight_tf32(w, groups):
    lib = L.load()
    _chk(w, f32, "w")
    D, cpg, ks = w.shape
    n = lib.f5b_convpos_packed_elems_tf32(D, groups, ks)
    out = torch.empty(n, dtype=f32, device=w.device)
    L.check(lib.f5b_pack_convpos_weight_tf32(w.data_ptr(), out.data_ptr(), D, groups, ks, L.stream()), "f5b_pack_convpos_weight_tf32")
    return out


def convpos_tf32(x, wpk, bias, B, n, D, groups, ksize, out=None, resid=None):
    lib = L.load()
    _chk(x, f32, "x"); _chk(wpk, f32, "wpk"); _chk(bias, f32, "bias"); _chk(out, f32, "out"); _chk(resid, f32, "resid")
    mode = 0 if resid is None else 1
    L.check(lib.f5b_convpos_tf32(x.data_ptr(), wpk.data_ptr(), bias.data_ptr(), L.ptr(out), L.ptr(resid), B, n, D, groups, ksize, mode,
                                 L.stream()), "f5b_convpos_tf32")
    return out if resid is None else resid


def melspec(wav, fb, ranges, n_mels):
    lib = L.load()
    _chk(wav, f32, "wav"); _chk(fb, f32, "fb"); _chk(ranges, torch.int32, "ranges")
    B, Ls = wav.shape
    T = 1 + Ls // 256
    out = torch.empty(B, T, n_mels, dtype=f32, device=wav.device)
    L.check(lib.f5b_melspec(wav.data_ptr(), fb.data_ptr(), ranges.data_ptr(), out.data_ptr(), B, Ls, n_mels, L.stream()), "f5b_melspec")
    return out


def istft_head(head, B, T):
    lib = L.load()
    _chk(head, f32, "head")
    frames = torch.empty(B * T, 1024, dtype=f32, device=head.device)
    wav = torch.empty(B, 256 * (T - 1), dtype=f32, device=head.device)
    L.check(lib.f5b_istft_head(head.data_ptr(), head.shape[-1], frames.data_ptr(), wav.data_ptr(), B, T, L.stream()), "f5b_istft_head")
    return wav


def time_sinus(t):
    lib = L.load()
    _chk(t, f32, "t")
    out = torch.empty(t.numel(), 256, dtype=bf16, device=t.device)
    L.check(lib.f5b_time_sinus(t.data_ptr(), out.data_ptr(), t.numel(), L.stream()), "f5b_time_sinus")
    return out


def attn_fwd_lse(q, k, v, ld, out, lse, lens, lens_mod, B, H, n, scale=0.125):
    """training forward: attn_fwd + per-row log2-sum-exp lse f32 [B, H, n]"""
    lib = L.load()
    for t, nm in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        if not t.is_cuda or t.dtype != bf16:
            raise L.F5bError(f"{nm}: expected a CUDA bf16 tensor")
    _chk(lse, f32, "lse"); _chk(lens, torch.int32, "lens")
    L.check(lib.f5b_attn_fwd_lse(q.data_ptr(), k.data_ptr(), v.data_ptr(), ld, out.data_ptr(), lse.data_ptr(), L.ptr(lens), lens_mod,
                                 B, H, n, scale, L.stream()), "f5b_attn_fwd_lse")
    return out


def attn_bwd(q, k, v, ld, out, dout, lse, dqkv, lens, lens_mod, B, H, n, scale=0.125, rope=None, rope_heads=0):
    """dqkv bf16 [B*n, 3*H*64] = gradient of the fused (pre-RoPE) QKV projection output"""
    lib = L.load()
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        if not t.is_cuda or t.dtype != bf16:
            raise L.F5bError(f"{nm}: expected a CUDA bf16 tensor")
    _chk(out, bf16, "out"); _chk(dout, bf16, "dout"); _chk(lse, f32, "lse"); _chk(dqkv, bf16, "dqkv"); _chk(rope, f32, "rope")
    _chk(lens, torch.int32, "lens")
    delta = torch.empty(B * H * n, dtype=f32, device=out.device)
    dq_ws = torch.empty(B * n, H * 64, dtype=f32, device=out.device)
    L.check(lib.f5b_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), ld, out.data_ptr(), dout.data_ptr(), out.shape[-1], lse.data_ptr(),
                             delta.data_ptr(), dq_ws.data_ptr(), dqkv.data_ptr(), dqkv.shape[-1], L.ptr(lens), lens_mod, B, H, n, scale,
                             L.ptr(rope), rope_heads, L.stream()), "f5b_attn_bwd")
    return dqkv


def crossfade_concat(waves, cross_fade_samples: int):
    """Cross-fade concatenation of 1-D fp32 CUDA waveforms with the fold of infer/f5tts_wrapper.py:549-575 (each step blends
    min(cross_fade_samples, len(accumulated), len(next)) samples), on the device: K - 1 small launches, no host round trip."""
    if not waves:
        raise ValueError("no waves")
    waves = [w.reshape(-1).contiguous() for w in waves]
    for w in waves:
        _chk(w, torch.float32, "wave")
    out = torch.empty(sum(w.numel() for w in waves), dtype=torch.float32, device=waves[0].device)
    n = waves[0].numel()
    out[:n].copy_(waves[0])
    lib = L.load()
    for w in waves[1:]:
        cfs = max(0, min(int(cross_fade_samples), n, w.numel()))
        L.check(lib.f5b_crossfade_append(out.data_ptr(), n, w.data_ptr(), w.numel(), cfs, L.stream()), "f5b_crossfade_append")
        n = n - cfs + w.numel()
    return out[:n]


def pcm16(wave):
    """np.int16(wave * 32767) (socket_server.py:54) on the device, saturating"""
    w = wave.reshape(-1).contiguous()
    _chk(w, torch.float32, "wave")
    out = torch.empty(w.numel(), dtype=torch.int16, device=w.device)
    L.check(L.load().f5b_pcm16(w.data_ptr(), out.data_ptr(), w.numel(), L.stream()), "f5b_pcm16")
    return out.view(wave.shape)
