"""Unit parity of the backward-pass building blocks (through the C ABI) against torch autograd of the same op in fp32.
bf16 outputs: relative max error <= 1e-2; fp32 reductions (bias / modulation gradients): <= 2e-3 relative."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from eraxvif5tts_b200 import _lib as L
    return L, L.load()


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


@pytest.mark.parametrize("D,affine", [(1024, False), (768, False), (512, True), (128, False)])
def test_ln_modulate_bwd(D, affine):
    L, lib = _lib()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(D)
    B, n = 3, 75
    x = (torch.randn(B, n, D, generator=g) * 1.5 + 0.3).to(dev).requires_grad_(True)
    if affine:
        w = (torch.randn(D, generator=g) * 0.3 + 1).to(dev).requires_grad_(True)
        b = torch.randn(D, generator=g).to(dev).requires_grad_(True)
        y = F.layer_norm(x, (D,), w, b, 1e-6)
    else:
        sc = (torch.randn(B, D, generator=g) * 0.3).to(dev).requires_grad_(True)
        sh = torch.randn(B, D, generator=g).to(dev).requires_grad_(True)
        y = F.layer_norm(x, (D,), eps=1e-6) * (1 + sc[:, None]) + sh[:, None]
    dy = torch.randn(B, n, D, generator=g).to(dev).bfloat16()
    y.backward(dy.float())
    prev = torch.randn(B, n, D, generator=g).to(dev)
    for accumulate in (0, 1):
        dx = prev.clone()
        if affine:
            dw, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
            L.check(lib.f5b_ln_affine_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), dx.data_ptr(), accumulate, dw.data_ptr(), db.data_ptr(),
                                          B, n, D, 1e-6, L.stream()), "ln_affine_bwd")
            ref_s, ref_h = w.grad, b.grad
        else:
            dw, db = torch.zeros(B, D, device=dev), torch.zeros(B, D, device=dev)
            L.check(lib.f5b_ln_modulate_bwd(dy.data_ptr(), x.data_ptr(), sc.data_ptr(), D, dx.data_ptr(), accumulate, dw.data_ptr(),
                                            db.data_ptr(), B, n, D, 1e-6, L.stream()), "ln_modulate_bwd")
            ref_s, ref_h = sc.grad, sh.grad
        torch.cuda.synchronize()
        ref_dx = x.grad + (prev if accumulate else 0)
        assert _rel(dx, ref_dx) <= 2e-3
        assert _rel(dw, ref_s) <= 2e-3 and _rel(db, ref_h) <= 2e-3


def test_gate_add_and_gate_bwd_and_fused_ln():
    L, lib = _lib()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(2)
    B, n, D = 3, 70, 256
    lens = torch.tensor([70, 33, 1], dtype=torch.int32, device=dev)
    live = (torch.arange(n, device=dev)[None] < lens[:, None])[:, :, None]
    x = torch.randn(B, n, D, generator=g).to(dev)
    z = torch.randn(B, n, D, generator=g).to(dev).bfloat16()
    gate = torch.randn(B, D, generator=g).to(dev)
    sc, sh = (torch.randn(B, D, generator=g) * 0.2).to(dev), torch.randn(B, D, generator=g).to(dev)
    ref_x = x + gate[:, None] * z.float() * live
    out = torch.empty_like(x)
    L.check(lib.f5b_gate_add(x.data_ptr(), z.data_ptr(), gate.data_ptr(), D, lens.data_ptr(), out.data_ptr(), B, n, D, L.stream()), "gate_add")
    out2, hb = torch.empty_like(x), torch.empty(B, n, D, dtype=torch.bfloat16, device=dev)
    L.check(lib.f5b_gate_add_ln_modulate(x.data_ptr(), z.data_ptr(), gate.data_ptr(), D, lens.data_ptr(), out2.data_ptr(), sc.data_ptr(),
                                         sh.data_ptr(), D, hb.data_ptr(), B, n, D, 1e-6, L.stream()), "gate_add_ln_modulate")
    torch.cuda.synchronize()
    assert _rel(out, ref_x) <= 1e-6 and torch.equal(out, out2)
    ref_h = F.layer_norm(ref_x, (D,), eps=1e-6) * (1 + sc[:, None]) + sh[:, None]
    assert _rel(hb, ref_h) <= 1e-2
    # backward of the gated residual
    dx = torch.randn(B, n, D, generator=g).to(dev)
    dz = torch.empty(B, n, D, dtype=torch.bfloat16, device=dev)
    dgate, dbias = torch.zeros(B, D, device=dev), torch.zeros(D, device=dev)
    L.check(lib.f5b_gate_bwd(dx.data_ptr(), z.data_ptr(), gate.data_ptr(), D, lens.data_ptr(), dz.data_ptr(), dgate.data_ptr(), dbias.data_ptr(),
                             B, n, D, L.stream()), "gate_bwd")
    torch.cuda.synchronize()
    ref_dz = gate[:, None] * dx * live
    assert _rel(dz, ref_dz) <= 1e-2
    assert _rel(dgate, (dx * z.float() * live).sum(1)) <= 2e-3
    assert _rel(dbias, ref_dz.sum((0, 1))) <= 2e-3


@pytest.mark.parametrize("act,fn", [(1, lambda t: F.gelu(t, approximate="tanh")), (2, F.gelu), (3, F.silu), (4, F.mish)])
def test_act_fwd_bwd(act, fn):
    L, lib = _lib()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(act)
    rows, C = 200, 512
    h = (torch.randn(rows, C, generator=g) * 2).to(dev).bfloat16()
    du = torch.randn(rows, C, generator=g).to(dev).bfloat16()
    hf = h.float().requires_grad_(True)
    y = fn(hf)
    y.backward(du.float())
    u = torch.empty_like(h)
    L.check(lib.f5b_act_fwd(h.data_ptr(), u.data_ptr(), rows * C, act, L.stream()), "act_fwd")
    dh = torch.empty_like(h)
    db = torch.zeros(C, device=dev)
    L.check(lib.f5b_act_bwd(du.data_ptr(), h.data_ptr(), dh.data_ptr(), db.data_ptr(), rows, C, C, act, L.stream()), "act_bwd")
    torch.cuda.synchronize()
    assert _rel(u, y.detach()) <= 1e-2
    assert _rel(dh, hf.grad) <= 1e-2
    assert _rel(db, dh.float().sum(0)) <= 2e-3


@pytest.mark.parametrize("M,N,K", [(1000, 2048, 1024), (38400, 2048, 1024), (300, 384, 128), (4100, 1536, 768)])
def test_dual_output_gemm_matches_gemm_plus_activation_sweep(M, N, K):
    """F5B_EPI_BF16_DUAL (the training forward's FeedForward GEMM, model/modules.py:348-353): pre-activation and GELU output from one
    accumulator tile must be bit-identical to the plain GEMM followed by the f5b_act_fwd sweep it replaces (128- and 256-wide tiles)."""
    L, lib = _lib()
    from eraxvif5tts_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    h_ref = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    u_ref = torch.empty_like(h_ref)
    ops.gemm(a, w, epi=L.EPI_BF16, act=L.ACT_NONE, bias=bias, out=h_ref)
    L.check(lib.f5b_act_fwd(h_ref.data_ptr(), u_ref.data_ptr(), M * N, L.ACT_GELU_TANH, L.stream()), "act_fwd")
    h = torch.full_like(h_ref, float("nan"))
    u = torch.full_like(h_ref, float("nan"))
    ops.gemm(a, w, epi=L.EPI_BF16_DUAL, act=L.ACT_GELU_TANH, bias=bias, out=h, out2=u)
    torch.cuda.synchronize()
    assert torch.equal(h, h_ref)
    assert torch.equal(u, u_ref)
    ref = torch.nn.functional.gelu(a.float() @ w.float().t() + bias, approximate="tanh")
    assert (u.float() - ref).abs().max().item() < 3e-2


def test_grn_gelu_bwd_and_dwconv_bwd_and_lookup_bwd():
    L, lib = _lib()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(9)
    B, n, C = 2, 90, 256
    p1 = torch.randn(B, n, C, generator=g).to(dev).bfloat16()
    gamma, beta = torch.randn(C, generator=g).to(dev).requires_grad_(True), torch.randn(C, generator=g).to(dev).requires_grad_(True)
    pf = p1.float().requires_grad_(True)
    t2 = F.gelu(pf)
    t2r = t2 + (t2.detach().bfloat16().float() - t2.detach())   # the kernels see bf16 t2
    gx = torch.norm(t2r, p=2, dim=1, keepdim=True)
    nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
    t3 = gamma * (t2r * nx) + beta + t2r
    dt3 = torch.randn(B, n, C, generator=g).to(dev).bfloat16()
    t3.backward(dt3.float())
    t2b = t2.detach().bfloat16().contiguous()
    dp1 = torch.empty_like(p1)
    dgam, dbet, db1 = torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    stats = torch.empty(B * 3 * C, device=dev)
    L.check(lib.f5b_grn_gelu_bwd(dt3.data_ptr(), t2b.data_ptr(), p1.data_ptr(), gamma.data_ptr(), dp1.data_ptr(), dgam.data_ptr(), dbet.data_ptr(),
                                 db1.data_ptr(), stats.data_ptr(), B, n, C, L.stream()), "grn_gelu_bwd")
    torch.cuda.synchronize()
    assert _rel(dp1, pf.grad) <= 1.5e-2
    assert _rel(dgam, gamma.grad) <= 5e-3 and _rel(dbet, beta.grad) <= 2e-3 and _rel(db1, dp1.float().sum((0, 1))) <= 2e-3
    # depth-wise conv k=7
    x = torch.randn(B, n, C, generator=g).to(dev).requires_grad_(True)
    w = (torch.randn(C, 1, 7, generator=g) * 0.3).to(dev).requires_grad_(True)
    b = torch.randn(C, generator=g).to(dev).requires_grad_(True)
    y = F.conv1d(x.transpose(1, 2), w, b, padding=3, groups=C).transpose(1, 2)
    dy = torch.randn(B, n, C, generator=g).to(dev)
    y.backward(dy)
    dx = torch.ones(B, n, C, device=dev)
    dw, dbb = torch.zeros(C, 7, device=dev), torch.zeros(C, device=dev)
    L.check(lib.f5b_dwconv7_bwd(dy.contiguous().data_ptr(), x.data_ptr(), w.data_ptr(), dx.data_ptr(), dw.data_ptr(), dbb.data_ptr(), B, n, C,
                                L.stream()), "dwconv7_bwd")
    torch.cuda.synchronize()
    assert _rel(dx - 1, x.grad) <= 1e-4 and _rel(dw, w.grad.reshape(C, 7)) <= 1e-4 and _rel(dbb, b.grad) <= 1e-4
    # embedding scatter
    V, T, nt = 50, 64, 30
    ids = torch.randint(-1, V - 1, (B, nt), generator=g).to(dev)
    dh = torch.randn(B, n, T, generator=g).to(dev)
    table = torch.zeros(V, T, device=dev, requires_grad=True)
    tok = torch.zeros(B, n, dtype=torch.long, device=dev)
    tok[:, :nt] = ids + 1
    F.embedding(tok, table).backward(dh)
    dt = torch.zeros(V, T, device=dev)
    L.check(lib.f5b_text_lookup_bwd(dh.data_ptr(), ids.data_ptr(), nt, dt.data_ptr(), B, n, T, V, 0, L.stream()), "text_lookup_bwd")
    torch.cuda.synchronize()
    assert _rel(dt, table.grad) <= 1e-5


def test_convpos_input_gradient_and_mse_grad():
    L, lib = _lib()
    from eraxvif5tts_b200 import ops
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(4)
    B, n, D, G, K = 2, 150, 256, 16, 31
    x = torch.randn(B, n, D, generator=g).to(dev).bfloat16()
    w = (torch.randn(D, D // G, K, generator=g) * 0.05).to(dev)
    xf = x.float().requires_grad_(True)
    y = F.conv1d(xf.transpose(1, 2), w.bfloat16().float(), None, padding=K // 2, groups=G).transpose(1, 2)
    dy = torch.randn(B, n, D, generator=g).to(dev).bfloat16()
    y.backward(dy.float())
    wt = torch.empty(lib.f5b_convpos_packed_elems(D, G, K), dtype=torch.bfloat16, device=dev)
    L.check(lib.f5b_pack_convpos_weight_t(w.data_ptr(), wt.data_ptr(), D, G, K, L.stream()), "pack_t")
    dx = torch.empty(B * n, D, dtype=torch.bfloat16, device=dev)
    L.check(lib.f5b_convpos(dy.data_ptr(), wt.data_ptr(), None, dx.data_ptr(), None, B, n, D, G, K, 2, L.stream()), "convpos mode 2")
    torch.cuda.synchronize()
    assert _rel(dx.reshape(B, n, D), xf.grad) <= 1e-2
    # loss gradient
    rows, C = 333, 100
    pred, flow = torch.randn(rows, C, generator=g).to(dev).requires_grad_(True), torch.randn(rows, C, generator=g).to(dev)
    mask = (torch.rand(rows, generator=g) < 0.6).to(dev)
    loss = F.mse_loss(pred, flow, reduction="none")[mask].mean()
    loss.backward()
    ws, out2 = torch.empty(2048, device=dev), torch.empty(2, device=dev)
    m8 = mask.to(torch.uint8).contiguous()
    L.check(lib.f5b_masked_mse(pred.data_ptr(), flow.data_ptr(), m8.data_ptr(), ws.data_ptr(), out2.data_ptr(), rows, C, L.stream()), "mse")
    dp = torch.empty(rows, 128, dtype=torch.bfloat16, device=dev)
    L.check(lib.f5b_mse_grad(pred.data_ptr(), flow.data_ptr(), m8.data_ptr(), out2.data_ptr(), dp.data_ptr(), rows, C, 128, L.stream()), "mse_grad")
    torch.cuda.synchronize()
    assert abs(float(out2[0]) - float(loss)) <= 1e-5 * float(loss)
    assert _rel(dp[:, :C], pred.grad) <= 1e-2 and float(dp[:, C:].abs().max()) == 0.0
