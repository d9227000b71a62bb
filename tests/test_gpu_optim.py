"""Optimizer step (trainer.py:1280-1287, 1321) on the CUDA path vs torch.optim.AdamW + clip_grad_norm_ + an EMA lerp."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.GELU(), torch.nn.Linear(64, 13), torch.nn.LayerNorm(13)).cuda()


def test_fused_adamw_clip_ema_matches_torch():
    from eraxvif5tts_b200.optim import EmaSchedule, FlatAdamW, WarmupLinearDecay
    ref = _net()
    ours = copy.deepcopy(ref)
    opt_ref = torch.optim.AdamW(ref.parameters(), lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01)
    ema_ref = [p.detach().clone() for p in ref.parameters()]
    sched = EmaSchedule(update_after_step=2, update_every=2)
    opt = FlatAdamW(ours, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0, with_ema=True, ema_schedule=sched)
    lrs = WarmupLinearDecay(1e-3, 3, 10)
    g = torch.Generator(device="cuda").manual_seed(1)
    for it in range(8):
        grads = [torch.randn(p.shape, device="cuda", generator=g) * (3.0 if it % 2 else 0.05) for p in ref.parameters()]
        for p, gr in zip(ref.parameters(), grads):
            p.grad = gr.clone()
        for p, gr in zip(ours.parameters(), grads):
            p.grad.copy_(gr)
        lr = lrs.lr(it)
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        for grp in opt_ref.param_groups:
            grp["lr"] = lr
        opt_ref.step()
        d = sched.decay_for_call(it + 1)
        if d == "copy":
            ema_ref = [p.detach().clone() for p in ref.parameters()]
        elif d is not None:
            ema_ref = [e.lerp(p.detach(), 1 - d) for e, p in zip(ema_ref, ref.parameters())]
        opt.step(lr=lr)
        torch.cuda.synchronize()
        for a, b in zip(ref.parameters(), ours.parameters()):
            assert torch.allclose(a, b, rtol=2e-5, atol=1e-6), (it, float((a - b).abs().max()))
    ema = opt.ema_state_dict()
    for (k, p), e in zip(ours.named_parameters(), ema_ref):
        assert torch.allclose(ema["ema_model." + k], e, rtol=2e-5, atol=1e-6), k
    assert abs(lrs.lr(0) - 1e-3 * 1e-8) < 1e-12 and abs(lrs.lr(3) - 1e-3) < 1e-12 and lrs.lr(10) < 1e-10


def test_grad_norm_and_no_clip_path():
    from eraxvif5tts_b200.optim import FlatAdamW
    net = _net()
    opt = FlatAdamW(net, lr=1e-3, max_grad_norm=0.0)
    for p in net.parameters():
        p.grad.normal_()
    ref = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in net.parameters())).float()
    assert torch.allclose(opt.grad_norm(), ref, rtol=1e-5)
    before = [p.detach().clone() for p in net.parameters()]
    opt.step()
    torch.cuda.synchronize()
    assert all(not torch.equal(a, b) for a, b in zip(before, net.parameters()))
