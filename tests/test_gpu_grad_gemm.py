"""Backward GEMMs of nn.Linear on the tcgen05 engine with MN-major operands (no transposed copies): dgrad, wgrad (split-K with
TMA reduce-add), and the generic four operand-layout combinations, against fp32 torch matmuls of the same bf16 inputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _tn(A, a_mn, B, b_mn, M, N, K, f32_acc, splits=1, out=None):
    from eraxvif5tts_b200 import _lib as L
    lib = L.load()
    if out is None:
        out = torch.zeros(M, N, dtype=torch.float32 if f32_acc else bf16, device="cuda")
    L.check(lib.f5b_gemm_tn(A.data_ptr(), A.stride(0), int(a_mn), B.data_ptr(), B.stride(0), int(b_mn), out.data_ptr(), out.stride(0),
                            int(f32_acc), M, N, K, splits, L.stream()), "f5b_gemm_tn")
    torch.cuda.synchronize()
    return out


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


@pytest.mark.parametrize("M,N,K", [(1000, 1024, 1024), (300, 200, 136), (4100, 1024, 3072), (129, 2048, 1024)])
def test_dgrad(M, N, K):
    """dX[M,K] = dY[M,N] @ W[N,K]  (reduction over N)"""
    g = torch.Generator(device="cuda").manual_seed(M + N)
    dy = (torch.randn(M, N, device="cuda", generator=g) * 0.5).to(bf16)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(N)).to(bf16)
    ref = dy.float() @ w.float()
    dx = _tn(dy, False, w, True, M, K, N, f32_acc=False)
    assert _rel(dx, ref) < 1e-2
    acc0 = torch.randn(M, K, device="cuda", generator=g)
    dx32 = _tn(dy, False, w, True, M, K, N, f32_acc=True, out=acc0.clone())
    assert _rel(dx32, acc0 + ref) < 1e-4


@pytest.mark.parametrize("M,N,K,splits", [(4000, 1024, 1024, 8), (38400, 1024, 2048, 37), (777, 200, 136, 3), (64, 3072, 1024, 1)])
def test_wgrad_split_k(M, N, K, splits):
    """dW[N,K] = dY[M,N]^T @ X[M,K]  (reduction over the M tokens, split across CTAs, accumulated into an existing gradient)"""
    g = torch.Generator(device="cuda").manual_seed(M + K)
    dy = (torch.randn(M, N, device="cuda", generator=g) * 0.5).to(bf16)
    x = (torch.randn(M, K, device="cuda", generator=g) / math.sqrt(M)).to(bf16)
    ref = dy.float().t() @ x.float()
    prev = torch.randn(N, K, device="cuda", generator=g) * 0.1
    dw = _tn(dy, True, x, True, N, K, M, f32_acc=True, splits=splits, out=prev.clone())
    assert _rel(dw, prev + ref) < 1e-4, _rel(dw, prev + ref)


def test_all_operand_layouts():
    M, N, K = 264, 328, 200
    g = torch.Generator(device="cuda").manual_seed(3)
    a = (torch.randn(M, K, device="cuda", generator=g)).to(bf16)
    b = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(bf16)
    ref = a.float() @ b.float().t()
    at, bt = a.t().contiguous(), b.t().contiguous()  # [K, M], [K, N]: the MN-major sources
    for a_mn in (False, True):
        for b_mn in (False, True):
            out = _tn(at if a_mn else a, a_mn, bt if b_mn else b, b_mn, M, N, K, f32_acc=True)
            assert _rel(out, ref) < 1e-4, (a_mn, b_mn, _rel(out, ref))
