"""Per-kernel GPU parity tests (through the C ABI) against plain PyTorch fp32 references of the same op.
bf16-output kernels: relative max error <= 1e-2 (bf16 rounding is 2^-9 = 2e-3 per value); f32-output kernels: <= 1e-4."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2
F32_TOL = 1e-4


@pytest.fixture(scope="module")
def D():
    import gpu_diag
    gpu_diag.RES.clear()
    return gpu_diag


def _check(D, tol_map=None):
    import torch
    torch.cuda.synchronize()
    for name, r in D.RES.items():
        assert "error" not in r, (name, r)
        for k, v in r.items():
            if k.endswith("rel"):
                tol = F32_TOL if ("f32" in name or "gate" in name or "resid" in name or name in ("cfg_euler",)) and k != "bf_rel" else BF16_TOL
                assert v <= tol, (name, k, v, tol)
            if k == "nan":
                assert v == 0, (name, r)
    D.RES.clear()


@pytest.mark.parametrize("M,N,K,epi,act", [
    (128, 128, 64, "bf16", 0), (128, 128, 256, "bf16", 0), (300, 200, 136, "f32", 0), (1000, 1024, 1024, "bf16", 0),
    (4100, 100, 1024, "f32", 0), (33, 1024, 256, "bf16", 3), (700, 512, 1024, "bf16", 2), (5000, 2048, 1024, "bf16", 1),
    (1, 8, 8, "f32", 0), (129, 3072, 1024, "bf16", 0), (20000, 1024, 2048, "bf16", 0),
])
def test_gemm(D, M, N, K, epi, act):
    D.gemm_case(M, N, K, epi, act)
    _check(D)


def test_gemm_gate_residual(D):
    D.gemm_gate_case(3, 200, 1024, 2048)
    D.gemm_gate_case(2, 77, 768, 768)
    _check(D)


@pytest.mark.parametrize("B,H,n,lens,rh", [(1, 2, 128, None, 1), (2, 16, 300, [300, 211], 1), (1, 16, 1875, None, 16),
                                             (3, 12, 257, [257, 1, 130], 1), (2, 2, 90, [90, 83], 2)])
def test_qkv_rope_and_attention(D, B, H, n, lens, rh):
    D.qkv_attn_case(B, H, n, lens, rope_heads=rh)
    _check(D)


@pytest.mark.parametrize("B,n,Dm", [(2, 300, 1024), (2, 200, 768), (2, 96, 128), (1, 31, 1024), (2, 1875, 1024), (1, 513, 1024),
                                     (3, 640, 768), (1, 4096, 1024)])
def test_convpos(D, B, n, Dm):
    D.convpos_case(B, n, Dm)
    _check(D)


def test_small_kernels(D):
    D.small_kernels()
    r = dict(D.RES)
    _check(D)
    assert r["cfg_euler"]["pad"] == 0.0
    assert r["time_sinus"]["abs"] < 1e-2


def test_spectral(D):
    D.spectral()
    r = dict(D.RES)
    D.RES.clear()
    assert r["melspec"]["abs"] < 1e-3 and r["melspec"]["mean_abs"] < 1e-5
    assert r["istft"]["abs"] < 1e-4 * max(1.0, r["istft"]["ref_max"])


def test_gemm_random_shape_sweep(D):
    """seeded sweep over ragged M / N / K (tails in every dimension, N and K multiples of 8 as the TMA pitch rule requires),
    all store paths: bf16 TMA store, f32 direct, gated residual TMA reduce-add"""
    import random
    import torch
    rng = random.Random(1234)
    for it in range(24):
        M = rng.choice([1, 7, 127, 128, 129, 255, 257, 1000, rng.randint(1, 3000)])
        N = 8 * rng.randint(1, 160)
        K = 8 * rng.randint(1, 140)
        kind = it % 3
        if kind == 0:
            D.gemm_case(M, N, K, "bf16", rng.choice([0, 1, 2, 3]))
        elif kind == 1:
            D.gemm_case(M, N, K, "f32", 0)
        else:
            n = rng.choice([1, 3, 50, 333])
            B = max(1, M // n)
            D.gemm_gate_case(B, n, 4 * ((N + 3) // 4), K)
        torch.cuda.synchronize()
    _check(D)


def test_attention_random_sweep(D):
    import random
    rng = random.Random(7)
    for it in range(10):
        B = rng.randint(1, 3)
        H = rng.choice([1, 2, 5, 12, 16])
        n = rng.choice([1, 2, 63, 64, 65, 127, 128, 129, 191, 200, 500, 1000])
        lens = None if it % 2 == 0 else [rng.randint(1, n) for _ in range(B)]
        D.qkv_attn_case(B, H, n, lens, rope_heads=rng.choice([0, 1, H]))
    _check(D)
