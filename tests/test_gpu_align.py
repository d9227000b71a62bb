"""GPU alignment search + duration predictor (csrc/align.cu) through the reference-named Python API, against the oracle and the
reference-generated golden outputs.  Alignments are index work: bit-exact.  The predictor is fp32: 1e-4 absolute."""
import os

import numpy as np
import pytest
import torch

from oracle import align_oracle as A

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "align_golden.pt")


def test_golden_cases_bit_exact():
    from eraxvif5tts_b200.model import alignment_utils as U
    gold = torch.load(GOLD)
    for c in gold["cases"]:
        sim = c["sim"].cuda()
        al, dur = U.viterbi_vectorized_alignment(sim, return_durations=True)
        assert torch.equal(al.cpu(), c["viterbi"]), tuple(sim.shape)
        assert torch.equal(dur.cpu().long(), c["viterbi"].sum(-1).long())
        assert torch.equal(U.get_durations_from_alignment(al).cpu(), c["viterbi"].sum(-1))
        if c["window"] is not None:
            aw = U.monotonic_alignment_search(sim, algorithm="window")
            assert torch.equal(aw.cpu(), c["window"]), tuple(sim.shape)
        assert torch.equal(U.monotonic_alignment_search(sim).cpu(), c["viterbi"])


@pytest.mark.parametrize("b,nt,T,kind", [(3, 37, 410, "randn"), (2, 150, 1200, "band"), (1, 1100, 48, "randn"), (2, 64, 64, "pos"),
                                        (1, 300, 30, "randn"), (2, 1, 17, "randn"), (1, 5, 1, "randn")])
def test_against_oracle(b, nt, T, kind):
    from eraxvif5tts_b200.model import alignment_utils as U
    g = torch.Generator().manual_seed(b * 1000 + nt + T)
    sim = torch.randn(b, nt, T, generator=g)
    if kind == "pos":
        sim = sim.abs()
    if kind == "band":
        n_idx = torch.arange(nt)[:, None].float() / nt
        t_idx = torch.arange(T)[None, :].float() / T
        sim = 3.0 * torch.exp(-((n_idx - t_idx) ** 2) * 200.0)[None] + 0.3 * sim
    ref_al, ref_dur = A.viterbi_alignment(sim.numpy())
    al, dur = U.viterbi_vectorized_alignment(sim.cuda(), return_durations=True)
    assert np.array_equal(al.cpu().numpy(), ref_al)
    assert np.array_equal(dur.cpu().numpy(), ref_dur)
    try:
        ref_w, ref_wd = A.windowed_alignment(sim.numpy())
    except IndexError:
        with pytest.raises(IndexError):
            U.windowed_monotonic_alignment(sim.cuda())
    else:
        aw, dw = U.windowed_monotonic_alignment(sim.cuda(), return_durations=True)
        assert np.array_equal(aw.cpu().numpy(), ref_w)
        assert np.array_equal(dw.cpu().numpy(), ref_wd)


def test_path_prob_bit_exact_and_dtype_preserved():
    from eraxvif5tts_b200 import _lib as L
    sim = torch.randn(2, 40, 333, generator=torch.Generator().manual_seed(5))
    d = sim.cuda()
    path, al = torch.empty_like(d), torch.empty_like(d)
    L.check(L.load().f5b_align_viterbi(d.data_ptr(), path.data_ptr(), al.data_ptr(), None, 2, 40, 333, L.stream()), "viterbi")
    assert np.array_equal(path.cpu().numpy(), A.viterbi_path_prob(sim.numpy()))
    from eraxvif5tts_b200.model import alignment_utils as U
    assert U.viterbi_vectorized_alignment(d.half()).dtype == torch.float16
    with pytest.raises(ValueError):
        U.monotonic_alignment_search(d, algorithm="nope")
    with pytest.raises(L.F5bError):
        U.viterbi_vectorized_alignment(sim)  # CPU tensor: no fallback


def test_duration_predictor_golden_and_oracle():
    from eraxvif5tts_b200.model import DurationPredictor
    d = torch.load(GOLD)["dp"]
    dp = DurationPredictor(*d["args"])
    dp.load_state_dict(d["state_dict"])
    dp = dp.cuda().eval()
    out = dp(d["ids"].cuda(), d["mask"].cuda())
    assert out.shape == d["out"].shape
    assert float((out.cpu() - d["out"]).abs().max()) <= 1e-4
    pout = dp.phoneme_forward((d["ids"] + 1).cuda(), d["mask"].cuda())
    assert float((pout.cpu() - d["phoneme_out"]).abs().max()) <= 1e-4
    # the wrapper's configuration (f5tts_wrapper-dur_pred.py:196): DurationPredictor(vocab, 512, 32, 3, 0.5), longer text
    torch.manual_seed(3)
    big = DurationPredictor(2545, 512, 32, 3, 0.5).eval()
    ids = torch.randint(0, 2545, (4, 300))
    lens = torch.tensor([300, 257, 31, 2])
    mask = (torch.arange(300)[None] < lens[:, None]).int()
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, -1))
    ref = A.duration_predictor(big.state_dict(), ids, mask, 1)
    got = big.cuda()(ids.cuda(), mask.cuda())
    assert float((got.cpu() - ref).abs().max()) <= 1e-4
    with pytest.raises(NotImplementedError):
        big.train()(ids.cuda(), mask.cuda())


@pytest.mark.parametrize("train,per_item", [(False, False), (True, False), (True, True)])
def test_duration_predictor_training_step_vs_oracle_autograd(train, per_item):
    """loss_and_grads (train-mode forward with the shared dropout masks, duration loss of distil_reload.py:1096-1124, hand-written
    backward) against torch autograd through the oracle: loss to 1e-4 relative, every parameter gradient to 2e-3 (fp32, different
    summation order).  Then a few SGD steps on one batch must drive the loss down."""
    from eraxvif5tts_b200.model import DurationPredictor, alignment_utils as U
    torch.manual_seed(11)
    dp = DurationPredictor(60, 48, 32, 3, 0.5).cuda()
    with torch.no_grad():
        for p in dp.parameters():
            p.add_(0.05 * torch.randn_like(p))
    dp.train(train)
    b, nt, T = 3, 28, 150
    g = torch.Generator().manual_seed(4)
    ids = torch.randint(0, 60, (b, nt), generator=g)
    lens = torch.tensor([28, 19, 7])
    mask = (torch.arange(nt)[None] < lens[:, None]).int()
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, -1))
    attn = U.viterbi_vectorized_alignment(torch.randn(b, nt, T, generator=g).cuda()).cpu()
    seed = 77
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in dp.state_dict().items()}
    logw_ref = A.duration_predictor(sd, ids, mask, 1, dropout=(0.5, seed) if train else None)
    loss_ref = A.duration_loss(logw_ref, attn, mask, per_item=per_item)
    loss_ref.backward()
    loss, logw = dp.loss_and_grads(ids.cuda(), mask.cuda(), attn=attn.cuda(), seed=seed, per_item=per_item)
    assert float((logw.cpu() - logw_ref.detach()).abs().max()) <= 1e-4
    assert abs(float(loss) - float(loss_ref.detach())) <= 1e-4 * abs(float(loss_ref.detach()))
    for k, p in dp.named_parameters():
        r = sd[k].grad
        assert r is not None, k
        err = float((p.grad.cpu() - r).norm()) / (float(r.norm()) + 1e-12)
        assert err <= 2e-3, (k, err)
    # gradients accumulate
    g0 = {k: p.grad.clone() for k, p in dp.named_parameters()}
    dp.loss_and_grads(ids.cuda(), mask.cuda(), attn=attn.cuda(), seed=seed, per_item=per_item)
    for k, p in dp.named_parameters():
        assert torch.allclose(p.grad, 2 * g0[k], rtol=1e-3, atol=1e-6 + 1e-4 * float(g0[k].abs().max())), k
    dp.eval()
    opt = torch.optim.SGD(dp.parameters(), lr=0.02)
    losses = []
    for _ in range(12):
        opt.zero_grad(set_to_none=False)
        l_, _ = dp.loss_and_grads(ids.cuda(), mask.cuda(), attn=attn.cuda(), per_item=True)
        opt.step()
        losses.append(float(l_))
    assert losses[-1] < 0.8 * losses[0], losses
