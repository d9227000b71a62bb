"""Training step parity: the hand-written backward (f5b_dit_train_backward) against torch autograd of the oracle's fp32 CPU
restatement of CFM.forward (oracle.cfm_loss, pinned to the reference modules by tests/test_oracle_golden.py), same weights, same
random draws.  Tolerance: activations and the gradients flowing between kernels are bf16, so every parameter gradient is judged
by its relative Frobenius error (<= 4e-2) and its cosine similarity (>= 0.999) with the autograd gradient; the loss itself to 2 %."""
import os

import pytest
import torch

from helpers import build_cfm, maxabs
from oracle import f5_oracle as O

pytestmark = pytest.mark.gpu

TEXT_PARAMS = "text_embed."


def _draws(cfg, B, n, seed):
    g = torch.Generator().manual_seed(seed)
    x1 = (torch.randn(B, n, cfg.mel_dim, generator=g) * 2 - 1.5).clamp(-11.5, 5)
    x0 = torch.randn(B, n, cfg.mel_dim, generator=g)
    time = torch.rand(B, generator=g)
    text = torch.randint(0, cfg.text_num_embeds, (B, max(4, n // 6)), generator=g)
    span = torch.zeros(B, n, dtype=torch.bool)
    for b in range(B):
        a = int(torch.randint(0, n // 3, (1,), generator=g))
        span[b, a:a + n // 2] = True
    return x1, x0, time, text, span


def _oracle_grads(sd, cfg, x1, text, span, x0, time, da, dt, dropout=None):
    leaf = {k: v.clone().float().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    full = dict(sd)
    full.update(leaf)
    loss, _, pred = O.cfm_loss(full, cfg, x1, text, span, x0, time, da, dt, dropout=dropout)
    loss.backward()
    return float(loss), pred.detach(), {k: v.grad for k, v in leaf.items() if v.grad is not None}


def _compare(model, ref, skip_prefix=None, fro_tol=4e-2, cos_tol=0.999):
    worst = []
    for k, p in model.named_parameters():
        if skip_prefix and k.startswith("transformer." + skip_prefix):
            continue
        r = ref.get(k)
        assert r is not None, k
        g = p.grad.detach().float().cpu().reshape(r.shape)
        rn = float(r.norm())
        if rn < 1e-12:
            assert float(g.norm()) < 1e-6, k
            continue
        fro = float((g - r).norm()) / rn
        cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
        worst.append((fro, cos, k))
    worst.sort(reverse=True)
    # softmax attention is invariant to adding one vector to every key (q.(k + c) shifts a row's scores by a constant), so the exact
    # gradient of to_k.bias is ZERO for every head without RoPE: what autograd and the kernels return there is rounding noise around
    # a tiny signal from the RoPE head, and gets a correspondingly looser bar
    def ok(w):
        loose = w[2].endswith("attn.to_k.bias")
        return w[0] <= (0.15 if loose else fro_tol) and w[1] >= (0.99 if loose else cos_tol)
    bad = [w for w in worst if not ok(w)]
    assert not bad, bad[:8]
    return worst


@pytest.mark.parametrize("da,dt", [(False, False), (True, True)])
def test_backward_vs_oracle_autograd(da, dt):
    from eraxvif5tts_b200.train import TrainEngine
    cfg = O.DiTConfig.tiny()
    model, sd = build_cfm(cfg, 0)
    eng = TrainEngine(model)
    B, n = 3, 150
    x1, x0, time, text, span = _draws(cfg, B, n, 7)
    ref_loss, ref_pred, ref = _oracle_grads(sd, cfg, x1, text, span, x0, time, da, dt)
    eng.zero_grad()
    loss, cond, pred = eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dict(rand_span_mask=span, x0=x0, time=time, drop_audio_cond=da,
                                                                               drop_text=dt))
    eng._fold_split_grads()
    torch.cuda.synchronize()
    assert maxabs(pred, ref_pred) <= 2e-2
    assert abs(float(loss) - ref_loss) <= 2e-2 * ref_loss
    worst = _compare(model, ref)
    print("worst gradients (fro err, cos, key):", worst[:5])


def test_activation_checkpointing_same_gradients_less_memory():
    """TrainEngine(checkpoint_activations=True) — the reference's DiT(checkpoint_activations=True), dit.py:221-223: every block keeps
    only its input and its forward is re-run in the backward.  Same loss / prediction bit for bit, gradients equal up to the order of
    the split-K / reduce-add sums (also with dropout: the masks are regenerated from the seed), and a several times smaller workspace."""
    from eraxvif5tts_b200.train import TrainEngine
    cfg = O.DiTConfig.tiny(depth=4)
    B, n = 3, 152
    x1, x0, time, text, span = _draws(cfg, B, n, 21)
    for p in (0.0, 0.1):
        draws = dict(rand_span_mask=span, x0=x0, time=time, drop_audio_cond=False, drop_text=False, dropout_seed=7)
        res = {}
        for ck in (False, True):
            model, sd = build_cfm(cfg, 0)
            eng = TrainEngine(model, dropout=p, checkpoint_activations=ck)
            eng.zero_grad()
            loss, _, pred = eng.loss_and_grads(x1.cuda(), text.cuda(), draws=draws)
            torch.cuda.synchronize()
            res[ck] = (float(loss), pred.clone(), eng.g.clone(), eng._ws.numel())
        assert res[True][0] == res[False][0] and torch.equal(res[True][1], res[False][1])
        assert float((res[True][2] - res[False][2]).norm()) <= 1e-3 * float(res[False][2].norm())
        assert res[True][3] < 0.6 * res[False][3], (res[True][3], res[False][3])  # depth 4: fixed buffers weigh in; depth 22: 4.6 vs 29.6 GB
    # the DiT's own flag is the default
    from eraxvif5tts_b200.model import CFM, DiT
    tr = DiT(dim=cfg.dim, depth=2, heads=cfg.heads, ff_mult=cfg.ff_mult, mel_dim=cfg.mel_dim, text_num_embeds=cfg.text_num_embeds,
             text_dim=cfg.text_dim, conv_layers=cfg.conv_layers, checkpoint_activations=True)
    m = CFM(transformer=tr, mel_spec_kwargs=dict(n_mel_channels=cfg.mel_dim)).cuda()
    assert TrainEngine(m).checkpoint_activations is True


def test_backward_with_dropout_vs_oracle_autograd():
    """train-mode dropout (p = 0.1 as the reference trains, and a heavy p = 0.5): the oracle applies the SAME masks
    (oracle.dropout_multipliers restates the product's counter-based generator), so forward, loss and every gradient are judged
    at the bars of the p = 0 test; the same seed reproduces the step, another seed changes it, and p = 0 is the
    eval-mode forward."""
    from eraxvif5tts_b200.train import TrainEngine
    cfg = O.DiTConfig.tiny()
    B, n = 3, 152
    x1, x0, time, text, span = _draws(cfg, B, n, 21)
    base = dict(rand_span_mask=span, x0=x0, time=time, drop_audio_cond=False, drop_text=False)
    preds = {}
    for p, seed in ((0.1, 1234), (0.5, 99)):
        model, sd = build_cfm(cfg, 0)
        eng = TrainEngine(model, dropout=p)
        ref_loss, ref_pred, ref = _oracle_grads(sd, cfg, x1, text, span, x0, time, False, False, dropout=(p, seed))
        eng.zero_grad()
        loss, _, pred = eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dict(base, dropout_seed=seed))
        eng._fold_split_grads()
        torch.cuda.synchronize()
        assert maxabs(pred, ref_pred) <= 2e-2
        assert abs(float(loss) - ref_loss) <= 2e-2 * ref_loss
        _compare(model, ref)
        g1 = eng.g.clone()
        eng.zero_grad()
        _, _, pred2 = eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dict(base, dropout_seed=seed))
        eng._fold_split_grads()
        # same seed, same masks: the forward repeats bit for bit; the backward's split-K / reduce-add sums are order-dependent
        assert torch.equal(pred, pred2)
        assert float((g1 - eng.g).norm()) <= 1e-3 * float(g1.norm())
        _, _, pred3 = eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dict(base, dropout_seed=seed + 1))
        assert maxabs(pred3, pred) > 1e-3
        preds[p] = pred
    # the masks matter (the p = 0.5 forward is far from the eval forward) and p = 0 is exactly the eval forward
    model, sd = build_cfm(cfg, 0)
    e0 = TrainEngine(model, dropout=0.0)
    _, _, pa = e0.loss_and_grads(x1.cuda(), text.cuda(), draws=dict(base, dropout_seed=1))
    _, _, pb = e0.loss_and_grads(x1.cuda(), text.cuda(), draws=dict(base, dropout_seed=2))
    assert torch.equal(pa, pb)
    assert maxabs(preds[0.5], pa) > 1e-2


@pytest.mark.parametrize("p,attn_p,n", [(0.0, 0.1, 152), (0.1, 0.3, 203)])
def test_backward_with_attention_dropout_vs_oracle_autograd(p, attn_p, n):
    """the third dropout site, inside scaled_dot_product_attention (model/modules.py:490): forward and backward regenerate the same
    counter-based mask on the normalised probabilities; the oracle applies the identical mask, so prediction, loss and every
    parameter gradient are held to the bars of the p = 0 test.  n = 203 is not a multiple of 4 (mask rows are padded to 4*ceil(n/4))."""
    from eraxvif5tts_b200.train import TrainEngine
    cfg = O.DiTConfig.tiny()
    B = 2
    x1, x0, time, text, span = _draws(cfg, B, n, 23)
    base = dict(rand_span_mask=span, x0=x0, time=time, drop_audio_cond=False, drop_text=False)
    seed = 4321
    model, sd = build_cfm(cfg, 0)
    eng = TrainEngine(model, dropout=p, attn_dropout=attn_p)
    ref_loss, ref_pred, ref = _oracle_grads(sd, cfg, x1, text, span, x0, time, False, False, dropout=(p, seed, attn_p))
    eng.zero_grad()
    loss, _, pred = eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dict(base, dropout_seed=seed))
    eng._fold_split_grads()
    torch.cuda.synchronize()
    assert maxabs(pred, ref_pred) <= 2e-2
    assert abs(float(loss) - ref_loss) <= 2e-2 * ref_loss
    _compare(model, ref)
    # the mask matters: the same step without SDPA dropout predicts something else
    e0 = TrainEngine(build_cfm(cfg, 0)[0], dropout=p, attn_dropout=0.0)
    _, _, pred0 = e0.loss_and_grads(x1.cuda(), text.cuda(), draws=dict(base, dropout_seed=seed))
    assert maxabs(pred0, pred) > 1e-3


@pytest.mark.parametrize("kind,w,da,dt", [("mse", 0.0, False, False), ("l1", 0.0, True, True), ("mse", 0.3, True, False)])
def test_distillation_step_vs_oracle_autograd(kind, w, da, dt):
    """train/distil_reload.py:1044-1093: frozen deeper teacher + pruned student in one launch sequence; the four losses and every
    student gradient against autograd through the oracle's restatement (same bars as the plain training step)."""
    from eraxvif5tts_b200.train import TrainEngine
    tcfg, scfg = O.DiTConfig.tiny(depth=3), O.DiTConfig.tiny()
    teacher, tsd = build_cfm(tcfg, 5)
    student, ssd = build_cfm(scfg, 0)
    eng = TrainEngine(student)
    B, n = 3, 136
    x1, x0, time, text, span = _draws(scfg, B, n, 31)
    leaf = {k: v.clone().float().requires_grad_(True) for k, v in ssd.items() if v.is_floating_point()}
    full = dict(ssd)
    full.update(leaf)
    total, st, di, sp, ref_pred = O.distill_losses(full, scfg, tsd, tcfg, x1, text, span, x0, time, da, dt, alpha=0.4, loss_type=kind,
                                                   spec_l1_weight=w)
    total.backward()
    ref = {k: v.grad for k, v in leaf.items() if v.grad is not None}
    eng.zero_grad()
    loss, _, pred = eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dict(rand_span_mask=span, x0=x0, time=time, drop_audio_cond=da,
                                                                            drop_text=dt),
                                       distill=dict(teacher=teacher, alpha=0.4, loss_type=kind, spec_l1_weight=w))
    eng._fold_split_grads()
    torch.cuda.synchronize()
    got = eng.last_losses.cpu()
    assert maxabs(pred, ref_pred.detach()) <= 2e-2
    for g, r in zip(got[:4], (total, st, di, sp)):
        assert abs(float(g) - float(r)) <= 2e-2 * abs(float(r)) + 1e-6, (got, float(total), float(st), float(di), float(sp))
    assert float(got[4]) == float(span.sum())
    assert float(loss) == float(got[0])
    # the l1 terms' gradient is sign(p - T): where |p - T| is within the bf16 forward noise the sign is a coin flip, so those
    # runs get a wider Frobenius bar on top of the cosine check
    _compare(student, ref, fro_tol=4e-2 if kind == "mse" and w == 0 else 0.12, cos_tol=0.999 if kind == "mse" and w == 0 else 0.992)


def test_gradient_accumulation_and_step_changes_loss():
    """two backward passes accumulate; a few fused AdamW steps on one batch drive the loss down (end-to-end sanity of the step)"""
    from eraxvif5tts_b200.train import TrainEngine
    cfg = O.DiTConfig.tiny()
    model, sd = build_cfm(cfg, 0)
    eng = TrainEngine(model, lr=2e-3, weight_decay=0.0)
    B, n = 2, 96
    x1, x0, time, text, span = _draws(cfg, B, n, 11)
    dr = dict(rand_span_mask=span, x0=x0, time=time, drop_audio_cond=False, drop_text=False)
    eng.zero_grad()
    eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dr)
    eng._fold_split_grads()
    g1 = eng.g.clone()
    eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dr)
    eng._fold_split_grads()
    torch.cuda.synchronize()
    assert torch.allclose(eng.g, 2 * g1, rtol=2e-2, atol=1e-6 + 2e-3 * float(g1.abs().max()))
    losses = []
    for _ in range(8):
        eng.zero_grad()
        loss, _, _ = eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dr)
        eng.step()
        losses.append(float(loss))
    assert losses[-1] < 0.9 * losses[0], losses
    # the inference engine sees the updated weights (re-packed lazily)
    out = model.transformer(x=x1.cuda(), cond=x1.cuda(), text=text.cuda(), time=time.cuda(), drop_audio_cond=False, drop_text=False)
    assert torch.isfinite(out).all()


def test_backward_other_widths():
    """non-power-of-two width (dim 384 = 6 heads, 24 channels per conv group, text_dim 96, pe on ALL heads): the F5TTS_Small-like
    code paths of the LN / conv / attention backward"""
    from eraxvif5tts_b200.train import TrainEngine
    cfg = O.DiTConfig.tiny(dim=384, heads=6, depth=1, text_dim=96, conv_layers=1, pe_attn_head=None)
    model, sd = build_cfm(cfg, 3)
    eng = TrainEngine(model)
    B, n = 2, 200
    x1, x0, time, text, span = _draws(cfg, B, n, 21)
    ref_loss, ref_pred, ref = _oracle_grads(sd, cfg, x1, text, span, x0, time, False, False)
    eng.zero_grad()
    loss, cond, pred = eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dict(rand_span_mask=span, x0=x0, time=time, drop_audio_cond=False,
                                                                               drop_text=False))
    eng._fold_split_grads()
    torch.cuda.synchronize()
    assert maxabs(pred, ref_pred) <= 2e-2
    assert abs(float(loss) - ref_loss) <= 2e-2 * ref_loss
    _compare(model, ref)


def test_ragged_lens_and_raw_audio_batches():
    """what the data path hands over: zero-padded ragged batches with `lens`, and raw-audio batches (2-D input -> GPU MelSpec)"""
    from eraxvif5tts_b200.train import TrainEngine
    cfg = O.DiTConfig.tiny()
    model, sd = build_cfm(cfg, 0)
    eng = TrainEngine(model)
    B, n = 3, 120
    x1, x0, time, text, span = _draws(cfg, B, n, 31)
    lens = torch.tensor([n, 61, 7])
    x1 = x1 * (torch.arange(n)[None, :, None] < lens[:, None, None])  # collate pads with zeros
    mask = torch.arange(n)[None] < lens[:, None]
    ref_loss, ref_pred, ref = _oracle_grads(sd, cfg, x1, text, span & mask, x0, time, False, False)
    eng.zero_grad()
    loss, cond, pred = eng.loss_and_grads(x1.cuda(), text.cuda(), lens=lens.cuda(),
                                          draws=dict(rand_span_mask=span, x0=x0, time=time, drop_audio_cond=False, drop_text=False))
    eng._fold_split_grads()
    torch.cuda.synchronize()
    assert abs(float(loss) - ref_loss) <= 2e-2 * ref_loss
    _compare(model, ref)
    # raw audio: [b, nw] -> MelSpec on the GPU -> the same step; the loss must match the oracle fed with the oracle's mel
    g = torch.Generator().manual_seed(5)
    wav = 0.1 * torch.randn(2, 256 * 63, generator=g)
    mel = O.melspec(wav).permute(0, 2, 1).contiguous()  # [b, T, 100]
    T = mel.shape[1]
    x0 = torch.randn(2, T, cfg.mel_dim, generator=g)
    time = torch.rand(2, generator=g)
    text = torch.randint(0, cfg.text_num_embeds, (2, 10), generator=g)
    span = torch.zeros(2, T, dtype=torch.bool)
    span[:, 10:50] = True
    ref_loss, _, _ = O.cfm_loss(sd, cfg, mel, text, span, x0, time, False, False)
    eng.zero_grad()
    loss, _, _ = eng.loss_and_grads(wav.cuda(), text.cuda(), draws=dict(rand_span_mask=span, x0=x0, time=time, drop_audio_cond=False,
                                                                         drop_text=False))
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref_loss)) <= 2e-2 * float(ref_loss)
    assert torch.isfinite(eng.g).all() and float(eng.g.abs().max()) > 0


def test_training_checkpoint_round_trip_in_reference_format(tmp_path):
    """save -> keep training -> load restores weights, Adam moments and EMA bit for bit; the file has the reference Trainer's
    keys (trainer.py:521-530), its optimizer state loads into torch.optim.AdamW, and the inference loader reads both weight sets"""
    from eraxvif5tts_b200.train import TrainEngine
    from eraxvif5tts_b200.infer.utils_infer import load_checkpoint
    cfg = O.DiTConfig.tiny()
    model, sd = build_cfm(cfg, 0)
    eng = TrainEngine(model, lr=1e-3, with_ema=True)
    B, n = 2, 64
    x1, x0, time, text, span = _draws(cfg, B, n, 3)
    dr = dict(rand_span_mask=span, x0=x0, time=time, drop_audio_cond=False, drop_text=False)
    for _ in range(3):
        eng.zero_grad()
        eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dr)
        eng.step()
    path = str(tmp_path / "model_3.pt")
    eng.save_checkpoint(path, update=3, scheduler_state={"last_epoch": 3})
    snap = (eng.p.clone(), eng.m.clone(), eng.v.clone(), eng.ema.clone(), eng.step_count, eng.ema_calls)
    for _ in range(2):
        eng.zero_grad()
        eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dr)
        eng.step()
    assert not torch.equal(eng.p, snap[0])
    ck = torch.load(path, map_location="cpu", weights_only=True)
    assert set(ck) == {"model_state_dict", "optimizer_state_dict", "ema_model_state_dict", "scheduler_state_dict", "update"}
    assert {"initted", "step"} <= set(ck["ema_model_state_dict"]) and all(
        k.startswith("ema_model.") for k in ck["ema_model_state_dict"] if k not in ("initted", "step"))
    assert eng.load_checkpoint(path) == 3 + 1  # the reference resumes at update + 1 (trainer.py:812)
    torch.cuda.synchronize()
    assert torch.equal(eng.p, snap[0]) and torch.equal(eng.m, snap[1]) and torch.equal(eng.v, snap[2]) and torch.equal(eng.ema, snap[3])
    assert (eng.step_count, eng.ema_calls) == snap[4:]
    assert torch.equal(eng.mirror.float(), eng.p.bfloat16().float())
    # torch's own optimizer accepts the state
    ref_opt = torch.optim.AdamW([torch.nn.Parameter(p.detach().clone().cpu()) for p in model.parameters()], lr=1e-3)
    ref_opt.load_state_dict(ck["optimizer_state_dict"])
    # the reference-style inference loader reads the online and the EMA weights (f5tts_wrapper.py:224-249)
    m2, _ = build_cfm(cfg, 1)
    load_checkpoint(m2, path, "cuda", use_ema=False)
    k0 = "transformer.transformer_blocks.0.attn.to_q.weight"
    assert torch.equal(dict(m2.named_parameters())[k0].cpu(), ck["model_state_dict"][k0])
    m3, _ = build_cfm(cfg, 1)
    load_checkpoint(m3, path, "cuda", use_ema=True)
    assert torch.equal(dict(m3.named_parameters())[k0].cpu(), ck["ema_model_state_dict"]["ema_model." + k0])


def test_backward_base_width_vs_oracle_autograd_on_gpu():
    """F5TTS_Base widths (D 1024, 16 heads, 64 channels per conv group, text_dim 512, 4 ConvNeXt blocks) at a realistic length: the
    CTA-pair gradient GEMMs, the attention backward with many key tiles and the grouped-conv weight gradient at the sizes cfg-5 uses.
    Oracle autograd runs in fp32 on the same GPU."""
    from eraxvif5tts_b200.train import TrainEngine
    cfg = O.DiTConfig(depth=2)
    model, sd = build_cfm(cfg, 0)
    dev = torch.device("cuda", 0)
    eng = TrainEngine(model)
    B, n = 2, 700
    x1, x0, time, text, span = _draws(cfg, B, n, 17)
    leaf = {k: v.to(dev).clone().float().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    full = {k: v.to(dev) for k, v in sd.items()}
    full.update(leaf)
    rope_cpu = O.rotary_freqs
    O.rotary_freqs = lambda n_, d=64, theta=10000.0: rope_cpu(n_, d, theta).to(dev)
    try:
        ref_loss, _, ref_pred = O.cfm_loss(full, cfg, x1.to(dev), text.to(dev), span.to(dev), x0.to(dev), time.to(dev), False, False)
        ref_loss.backward()
    finally:
        O.rotary_freqs = rope_cpu
    ref = {k: v.grad.cpu() for k, v in leaf.items() if v.grad is not None}
    eng.zero_grad()
    loss, cond, pred = eng.loss_and_grads(x1.cuda(), text.cuda(), draws=dict(rand_span_mask=span, x0=x0, time=time, drop_audio_cond=False,
                                                                               drop_text=False))
    eng._fold_split_grads()
    torch.cuda.synchronize()
    assert maxabs(pred, ref_pred.detach()) <= 2e-2 * max(1.0, float(ref_pred.abs().max()))
    assert abs(float(loss) - float(ref_loss)) <= 2e-2 * float(ref_loss)
    worst = _compare(model, ref)
    print("worst gradients at Base width (fro err, cos, key):", worst[:4])


class _ToyDataset(torch.utils.data.Dataset):
    """items shaped like the reference's CustomDataset (dataset.py:92-135): mel_spec [n_mels, T], text"""

    def __init__(self, n_items=22, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.lens = [int(x) for x in torch.randint(40, 120, (n_items,), generator=g)]
        self.mels = [(torch.randn(100, L_, generator=g) * 2 - 1.5).clamp(-11.5, 5) for L_ in self.lens]
        self.texts = ["".join(chr(97 + int(c)) for c in torch.randint(0, 26, (max(3, L_ // 8),), generator=g)) for L_ in self.lens]

    def get_frame_len(self, index):
        return self.lens[index]

    def __len__(self):
        return len(self.lens)

    def __getitem__(self, index):
        return dict(mel_spec=self.mels[index], text=self.texts[index])


def test_trainer_loop_checkpoints_and_resume(tmp_path):
    """model.Trainer: the reference's loop shape (frame batching, accumulation incl. the trailing partial window, warm-up / decay,
    checkpoint rotation, resume from model_last.pt) over TrainEngine"""
    from eraxvif5tts_b200.model import Trainer
    cfg = O.DiTConfig.tiny()
    ds = _ToyDataset()

    def make():
        model, _ = build_cfm(cfg, 0)
        model.vocab_char_map = {chr(97 + i): i for i in range(26)}
        return Trainer(model, epochs=2, learning_rate=2e-3, weight_decay=0.0, num_warmup_updates=2, save_per_updates=3,
                       keep_last_n_checkpoints=1, checkpoint_path=str(tmp_path / "ck"), batch_size_per_gpu=400, batch_size_type="frame",
                       max_samples=8, grad_accumulation_steps=2, last_per_updates=4)
    tr = make()
    import math
    from eraxvif5tts_b200.data import DynamicBatchSampler
    per_epoch = len(DynamicBatchSampler(torch.utils.data.SequentialSampler(ds), 400, max_samples=8))
    updates = tr.train(ds, num_workers=0, resumable_with_seed=666)
    assert updates == math.ceil(per_epoch / 2) * 2 and len(tr.losses) == updates
    assert all(math.isfinite(x) for x in tr.losses)
    assert sum(tr.losses[-3:]) < sum(tr.losses[:3])  # it learns the toy set
    files = sorted(os.listdir(tmp_path / "ck"))
    assert "model_last.pt" in files and len([f for f in files if f != "model_last.pt"]) == 1, files
    ck = torch.load(tmp_path / "ck" / "model_last.pt", map_location="cpu", weights_only=True)
    assert ck["update"] == updates and {"model_state_dict", "optimizer_state_dict", "ema_model_state_dict"} <= set(ck)
    # resume: everything is already trained -> no further update, weights equal the checkpoint's
    tr2 = make()
    assert tr2.load_checkpoint() == updates + 1  # trainer.py:812
    assert tr2.train(ds, num_workers=0, resumable_with_seed=666) == updates + 1 and not tr2.losses
    w = dict(tr2.model.named_parameters())["transformer.proj_out.weight"]
    assert torch.equal(w.detach().cpu(), ck["model_state_dict"]["transformer.proj_out.weight"])
    # "sample" batching, one epoch, no accumulation
    model, _ = build_cfm(cfg, 0)
    model.vocab_char_map = {chr(97 + i): i for i in range(26)}
    tr3 = Trainer(model, epochs=1, learning_rate=1e-3, num_warmup_updates=1, save_per_updates=1000, checkpoint_path=str(tmp_path / "ck3"),
                  batch_size_per_gpu=6, batch_size_type="sample", keep_last_n_checkpoints=0)
    assert tr3.train(ds, num_workers=2) == math.ceil(len(ds) / 6)  # forked loader workers: collate must not pin there
    with pytest.raises(NotImplementedError):
        Trainer(model, epochs=1, learning_rate=1e-3, duration_predictor=object())
