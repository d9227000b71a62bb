"""Attention backward (tcgen05, transposed orientation) against fp32 torch autograd of the same masked softmax attention + RoPE
(model/modules.py:470-493 under loss.backward()).  Tolerances: gradients are bf16 on output and P / dS are rounded to bf16 before
the tensor-core products, so errors are judged relative to each gradient's own scale (max-abs <= 2e-2 * max|ref|, the hot
path's bf16 bar) and by mean relative error (<= 1e-2)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rope_apply(t, table):
    # t [B, n, h, 64] fp32; table [n, 32, 2] (cos, sin), interleaved pairs
    x = t.reshape(*t.shape[:-1], 32, 2)
    c, s = table[None, :, None, :, 0], table[None, :, None, :, 1]
    y0 = x[..., 0] * c - x[..., 1] * s
    y1 = x[..., 1] * c + x[..., 0] * s
    return torch.stack((y0, y1), dim=-1).flatten(-2)


def _run(B, H, n, lens, rope_heads, seed=0):
    from eraxvif5tts_b200 import _lib as L
    from eraxvif5tts_b200 import ops
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(seed)
    D = H * 64
    qkv_pre = (torch.randn(B * n, 3 * D, generator=g) * 1.2).to(dev).bfloat16()
    dout = (torch.randn(B * n, D, generator=g) * 0.5).to(dev).bfloat16()
    lens_t = torch.tensor(lens, dtype=torch.int32, device=dev) if lens is not None else None
    table = torch.empty(n, 32, 2, dtype=torch.float32, device=dev)
    L.check(L.load().f5b_rope_table(table.data_ptr(), n, L.stream()), "rope")

    # fp32 reference with autograd w.r.t. the pre-RoPE projections
    x = qkv_pre.float().requires_grad_(True)
    q, k, v = (x[:, i * D:(i + 1) * D].reshape(B, n, H, 64) for i in range(3))
    if rope_heads:
        q = torch.cat((_rope_apply(q[:, :, :rope_heads], table), q[:, :, rope_heads:]), dim=2)
        k = torch.cat((_rope_apply(k[:, :, :rope_heads], table), k[:, :, rope_heads:]), dim=2)
    qkv_post = torch.cat((q.reshape(B * n, D), k.reshape(B * n, D), v.reshape(B * n, D)), dim=1).detach().bfloat16().contiguous()
    # the kernels see bf16 post-RoPE q / k: use the same rounded values in the reference forward (straight-through)
    qr = q + (qkv_post[:, :D].float().reshape(B, n, H, 64) - q).detach()
    kr = k + (qkv_post[:, D:2 * D].float().reshape(B, n, H, 64) - k).detach()
    s = torch.einsum("bqhd,bkhd->bhqk", qr, kr) * 0.125
    keymask = None
    if lens is not None:
        keymask = torch.arange(n, device=dev)[None, :] < lens_t[:, None]
        s = s.masked_fill(~keymask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("bhqk,bkhd->bqhd", p, v).reshape(B, n, D)
    do = dout.float().reshape(B, n, D)
    if keymask is not None:
        o = o * keymask[:, :, None]  # padded query rows: output forced to zero (reference zeroes them after to_out)
    (o * do).sum().backward()
    ref = x.grad
    lse_ref = torch.logsumexp(s, dim=-1) * math.log2(math.e)  # [B, H, n]

    out = torch.empty(B * n, D, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, H, n, dtype=torch.float32, device=dev)
    ops.attn_fwd_lse(qkv_post[:, :D], qkv_post[:, D:], qkv_post[:, 2 * D:], 3 * D, out, lse, lens_t, 0, B, H, n)
    dqkv = torch.empty(B * n, 3 * D, dtype=torch.bfloat16, device=dev)
    ops.attn_bwd(qkv_post[:, :D], qkv_post[:, D:], qkv_post[:, 2 * D:], 3 * D, out, dout, lse, dqkv, lens_t, 0, B, H, n,
                 rope=table.reshape(n, 64), rope_heads=rope_heads)
    torch.cuda.synchronize()

    valid = torch.ones(B, n, dtype=torch.bool, device=dev) if keymask is None else keymask
    tol = 2e-2 + 2e-2 * o.detach().abs().max().item()
    fwd_err = (out.float().reshape(B, n, D) - o.detach()).abs() * valid[:, :, None]
    assert fwd_err.max().item() <= tol
    vm = valid[:, None, :].expand(B, H, n)
    assert (lse[vm] - lse_ref[vm]).abs().max() < 2e-2
    if keymask is not None:
        assert torch.isinf(lse[~vm]).all()
    got = dqkv.float()
    assert torch.isfinite(got).all()
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        a, r = got[:, sl], ref[:, sl]
        scale = r.abs().max().item()
        err = (a - r).abs().max().item()
        mre = ((a - r).abs().mean() / r.abs().mean()).item()
        assert err <= 2e-2 * scale + 1e-6, (name, err, scale)
        assert mre <= 2e-2, (name, mre)  # bf16 P / dS / outputs: typically 3-8e-3
        if keymask is not None:
            pad = (~valid).reshape(-1)
            assert got[pad][:, sl].abs().max().item() == 0.0, name  # masked rows take no gradient


@pytest.mark.parametrize("B,H,n", [(1, 1, 128), (2, 2, 256), (1, 2, 200), (2, 4, 333), (1, 16, 1200)])
def test_attn_bwd_full(B, H, n):
    _run(B, H, n, None, 0)


def test_attn_bwd_rope():
    _run(2, 4, 300, None, 1)
    _run(1, 3, 257, None, 3, seed=1)


def test_attn_bwd_ragged_lens():
    _run(3, 2, 400, [400, 130, 257], 1)
    _run(2, 2, 384, [1, 384], 0, seed=3)
