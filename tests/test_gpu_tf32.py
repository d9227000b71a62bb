"""The tf32 operand mode (F5bGemmArgs.tf32 / F5bDitDesc.precision 1 / DiT(precision="tf32")): the path that holds north_star's
"1e-3 relative error on the DiT velocity field in fp32".  Kernel-level checks against fp64 torch on the SAME tf32-rounded operands
(so the only differences are accumulation order and the rounding of outputs), then the model-level acceptance against the fp32 oracle
at F5TTS_Base depth 22 (relative Frobenius error of the velocity over the valid frames <= 1e-3, both CFG branches, two trajectory
points) and the NFE-32 sample."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import acceptance as A
from oracle import f5_oracle as O

from helpers import build_cfm

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TF32_EPS = 2.0 ** -11  # half an ulp of a 10-bit mantissa: the rounding of a tf32 output


def _ops():
    from eraxvif5tts_b200 import ops
    return ops


def _rand(*shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


@pytest.mark.parametrize("M,N,K,act", [
    (300, 384, 200, 0),       # ragged M, K not a multiple of the 32-element k-block, 128-wide tiles
    (1000, 100, 1024, 0),     # proj_out's shape class: N < one tile
    (4096, 1024, 1024, 1),    # 256-wide CTA-pair tiles (cta_group::2), GELU-tanh
    (2100, 2048, 512, 3),     # pair tiles with a ragged last m-block, SiLU
])
def test_gemm_tf32(M, N, K, act):
    ops = _ops()
    from eraxvif5tts_b200 import _lib as L
    a, w = ops.round_tf32(_rand(M, K, seed=1)), ops.round_tf32(_rand(N, K, seed=2, scale=K ** -0.5))
    bias = _rand(N, seed=3)
    ref = a.double() @ w.double().t() + bias.double()
    if act == 1:
        ref = F.gelu(ref, approximate="tanh")
    elif act == 3:
        ref = F.silu(ref)
    # F32 epilogue: plain fp32 result
    if act == 0:
        out = torch.empty(M, N, device="cuda")
        ops.gemm(a, w, epi=L.EPI_F32, bias=bias, out=out, tf32=True)
        torch.cuda.synchronize()
        err = (out.double() - ref).abs().max() / ref.abs().max()
        assert float(err) < 1e-5, float(err)  # fp32 accumulation inside the tensor core
    # BF16 epilogue in tf32 mode: fp32 output rounded to tf32
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(a, w, epi=L.EPI_BF16, act=act, bias=bias, out=out, tf32=True)
    torch.cuda.synchronize()
    assert torch.equal(out, ops.round_tf32(out)), "outputs must be tf32 values"
    excess = ((out.double() - ref).abs() - 1.05 * TF32_EPS * ref.abs()).max()  # tf32 rounding of the output + fp32 accumulation
    assert float(excess) < 1e-5, float(excess)


def test_gemm_tf32_gate_residual_and_addsrc():
    ops = _ops()
    from eraxvif5tts_b200 import _lib as L
    B, n, D, K = 3, 150, 256, 320
    M = B * n
    a, w = ops.round_tf32(_rand(M, K, seed=4)), ops.round_tf32(_rand(D, K, seed=5, scale=K ** -0.5))
    bias, gate, x0 = _rand(D, seed=6), _rand(B, D, seed=7), _rand(M, D, seed=8)
    lens = torch.tensor([150, 77, 1], dtype=torch.int32, device="cuda")
    x = x0.clone()
    ops.gemm(a, w, epi=L.EPI_GATE_RESID, bias=bias, out=x, rows_per_batch=n, gate=gate, gate_bstride=D, lens=lens, tf32=True)
    y = (a.double() @ w.double().t() + bias.double()).view(B, n, D) * gate.double()[:, None, :]
    keep = (torch.arange(n, device="cuda")[None, :] < lens[:, None]).unsqueeze(-1)
    ref = x0.double().view(B, n, D) + torch.where(keep, y, torch.zeros_like(y))
    assert float((x.double().view(B, n, D) - ref).abs().max()) < 1e-5
    # F32 epilogue with addsrc and the tf32 copy
    out, out2 = torch.empty(M, D, device="cuda"), torch.empty(M, D, device="cuda")
    ops.gemm(a, w, epi=L.EPI_F32, out=out, out2=out2, addsrc=x0, tf32=True)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t() + x0.double()
    assert float((out.double() - ref).abs().max()) < 1e-5
    assert torch.equal(out2, ops.round_tf32(out))


@pytest.mark.parametrize("B,H,n,lens", [(2, 3, 333, [333, 200]), (1, 16, 1000, None), (3, 2, 64, [64, 1, 33]), (2, 2, 130, [129, 130])])
def test_attention_tf32(B, H, n, lens):
    ops = _ops()
    D = H * 64
    qkv = ops.round_tf32(_rand(B * n, 3 * D, seed=11))
    lt = torch.tensor(lens, dtype=torch.int32, device="cuda") if lens else None
    out = torch.full((B * n, D), float("nan"), device="cuda")
    ops.attn_fwd_tf32(qkv, qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, lt, 0, B, H, n)
    torch.cuda.synchronize()
    q, k, v = (t.view(B, n, H, 64).permute(0, 2, 1, 3).double() for t in qkv.split(D, dim=1))
    s = q @ k.transpose(-1, -2) / 8.0
    L_ = torch.tensor(lens if lens else [n] * B, device="cuda")
    keym = torch.arange(n, device="cuda")[None, :] < L_[:, None]
    s = s.masked_fill(~keym[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B, n, D)
    ref = torch.where(keym.unsqueeze(-1), ref, torch.zeros_like(ref))  # query rows past len are written as zeros
    got = out.view(B, n, D).double()
    assert torch.isfinite(got).all()
    rel = float((got - ref).norm() / ref.norm())
    mx = float((got - ref).abs().max() / ref.abs().max())
    assert rel < 4e-4 and mx < 2e-3, (rel, mx)  # P and O are rounded to tf32 (2^-11 each)


@pytest.mark.parametrize("B,n,D", [(2, 300, 1024), (1, 129, 512), (3, 77, 768)])
def test_convpos_tf32(B, n, D):
    ops = _ops()
    groups, ks = 16, 31
    cpg = D // groups
    x = ops.round_tf32(_rand(B * n, D, seed=21))
    w = _rand(D, cpg, ks, seed=22, scale=(cpg * ks) ** -0.5)
    bias = _rand(D, seed=23, scale=0.1)
    wpk = ops.pack_convpos_weight_tf32(w, groups)
    wr = ops.round_tf32(w)
    y = F.conv1d(x.view(B, n, D).permute(0, 2, 1).double(), wr.double(), bias.double(), padding=ks // 2, groups=groups)
    ref = F.mish(y).permute(0, 2, 1).reshape(B * n, D)
    out = torch.full((B * n, D), float("nan"), device="cuda")
    ops.convpos_tf32(x, wpk, bias, B, n, D, groups, ks, out=out)
    torch.cuda.synchronize()
    excess = ((out.double() - ref).abs() - 1.05 * TF32_EPS * ref.abs()).max()
    assert float(excess) < 1e-5, float(excess)
    resid0 = _rand(B * n, D, seed=24)
    resid = resid0.clone()
    ops.convpos_tf32(x, wpk, bias, B, n, D, groups, ks, resid=resid)
    torch.cuda.synchronize()
    assert float((resid.double() - (resid0.double() + ref)).abs().max()) < 5e-5


def _record(tag, res):
    path = os.path.join(ROOT, "gpurun_out", "parity_r02.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = {}
    for src in (os.path.join(ROOT, "profiles", "r02_parity_acceptance.json"), path):  # start from the committed record
        if os.path.exists(src):
            try:
                data.update(json.load(open(src)))
            except Exception:  # noqa: BLE001
                pass
    data[tag] = res
    with open(path, "w") as f:
        json.dump(data, f, indent=1)
    print(f"[parity] {tag}: velocity rel-fro {res['velocity_rel_fro']:.3e}, max-abs {res['velocity_max_abs']:.3e}, "
          f"mel mean-abs (generated) {res['mel_mean_abs_generated']:.3e}", flush=True)


def test_dit_forward_tf32_small_vs_oracle():
    """every piece of the tf32 driver (text ConvNeXt blocks, input projection, conv_pos_embed, blocks, masks) at a small width"""
    cfg = O.DiTConfig(dim=256, depth=3, heads=4, text_dim=128, conv_layers=2)
    model, sd = build_cfm(cfg, 0)
    model.transformer.set_precision("tf32")
    res = A.sample_parity(model, sd, cfg, 60, [200, 131, 170], steps=8, cfg_strength=2.0, sway=-1.0, seed=0)
    assert res["velocity_rel_fro"] <= A.VEL_RTOL_FP32, res["velocity"]
    assert res["mel_mean_abs_generated"] <= 2e-3, res


@pytest.mark.parametrize("tag,cfg,ref_frames,totals", [
    ("tf32_base_d22_n1875", O.DiTConfig(), 563, [1875, 1610, 1333]),
    ("tf32_base_d22_b1_n940", O.DiTConfig(), 376, [940]),
])
def test_acceptance_tf32_velocity_1e3(tag, cfg, ref_frames, totals):
    """north_star's fp32 clause at BASELINE's own configuration: F5TTS_Base depth 22, NFE 32, sway -1, CFG 2.  Metric: relative
    Frobenius error of the velocity over the valid frames, per DiT.forward (cond and uncond branch, first and middle trajectory point),
    against the strict-fp32 oracle: <= 1e-3."""
    model, sd = build_cfm(cfg, 0)
    model.transformer.set_precision("tf32")
    res = A.sample_parity(model, sd, cfg, ref_frames, totals, steps=32, cfg_strength=2.0, sway=-1.0, seed=0)
    res["precision"] = "tf32"
    _record(tag, res)
    assert res["velocity_rel_fro"] <= A.VEL_RTOL_FP32, res["velocity"]
    assert res["mel_mean_abs_generated"] <= A.MEL_MEAN_TOL / 4, res


def test_acceptance_tf32_with_fp32_weights():
    """The oracle's synthetic weights are bf16-exact (so the bf16 path's weight copies are lossless); a real checkpoint is not.  With
    weights that need all 24 mantissa bits the tf32 mode also rounds the WEIGHTS (to nearest): same bar, F5TTS_Base depth 22."""
    cfg = O.DiTConfig()
    model, sd = build_cfm(cfg, 0)
    g = torch.Generator().manual_seed(77)
    sd2 = {k: (v * (1.0 + 1e-3 * torch.randn(v.shape, generator=g)) if v.is_floating_point() and v.ndim >= 2 else v) for k, v in sd.items()}
    model.load_state_dict(sd2, strict=False)
    model.transformer.set_precision("tf32")
    res = A.sample_parity(model, sd2, cfg, 376, [940], steps=32, cfg_strength=2.0, sway=-1.0, seed=0)
    res["precision"] = "tf32, weights not bf16-exact"
    _record("tf32_base_d22_b1_n940_fp32_weights", res)
    assert res["velocity_rel_fro"] <= A.VEL_RTOL_FP32, res["velocity"]
