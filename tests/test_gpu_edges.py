"""GPU edge cases the reference's code paths allow: maximum duration (4096 frames), one-frame key lengths, ragged batches with a
row that is almost all padding, text longer than the mel (truncation), no_ref_audio / edit_mask, cfg_strength 0 (no uncond
branch), duplicate state reuse across calls, non-multiple-of-8 sequence lengths."""
import pytest
import torch

from oracle import f5_oracle as O
from oracle.weights import synthetic_inputs

from helpers import build_cfm, maxabs

pytestmark = pytest.mark.gpu


def test_attention_and_gemm_at_max_sequence_length():
    import gpu_diag as D
    D.RES.clear()
    D.qkv_attn_case(1, 16, 4096, None, rope_heads=1)       # cfm.py:135 clamps durations to 4096
    D.qkv_attn_case(2, 2, 4093, [4093, 1], rope_heads=2)   # odd length, a row with a single valid key
    torch.cuda.synchronize()
    for name, r in D.RES.items():
        for k, v in r.items():
            if k.endswith("rel"):
                assert v <= 1e-2, (name, k, v)
            if k == "nan":
                assert v == 0, (name, r)
    D.RES.clear()


def test_dit_forward_ragged_extreme_and_long_text():
    cfg = O.DiTConfig.tiny()
    model, sd = build_cfm(cfg, 0)
    B, n = 3, 131
    cond, text, _, _ = synthetic_inputs(cfg, B, n, n, seed=5)
    text = torch.randint(0, cfg.text_num_embeds, (B, n + 40))  # more tokens than frames -> truncated (dit.py:51)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, n, cfg.mel_dim, generator=g)
    lens = torch.tensor([n, 2, 77])
    mask = torch.arange(n)[None, :] < lens[:, None]
    time = torch.tensor([0.1, 0.5, 0.9])  # per-row times (CFM.forward style), exercises the per-row modulation stride
    ref = O.dit_forward(sd, cfg, x, cond, text, time, False, False, mask)
    out = model.transformer(x=x.cuda(), cond=cond.cuda(), text=text.cuda(), time=time.cuda(), drop_audio_cond=False, drop_text=False,
                            mask=mask.cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert maxabs(out, ref) <= 2e-2, maxabs(out, ref)


@pytest.mark.parametrize("kw", [dict(cfg_strength=0.0), dict(no_ref_audio=True), dict(edit=True), dict(sway=None)])
def test_cfm_sample_option_paths_vs_oracle(kw):
    cfg = O.DiTConfig.tiny()
    model, sd = build_cfm(cfg, 0)
    cond, text, duration, lens = synthetic_inputs(cfg, 2, 40, [90, 71], seed=21)
    noise = []
    for d in duration.tolist():
        torch.manual_seed(0)
        noise.append(torch.randn(d, cfg.mel_dim))
    noise = torch.nn.utils.rnn.pad_sequence(noise, batch_first=True)
    edit_mask = None
    if kw.get("edit"):
        edit_mask = torch.ones(2, 40, dtype=torch.bool)
        edit_mask[:, 10:20] = False  # speech_edit.py:175-184 semantics: False = regenerate
    args = dict(steps=3, cfg_strength=kw.get("cfg_strength", 2.0), sway_sampling_coef=kw.get("sway", -1.0) if "sway" not in kw else None,
                seed=0, no_ref_audio=kw.get("no_ref_audio", False))
    ref_out, ref_traj = O.cfm_sample(sd, cfg, cond, text, duration, lens=lens, edit_mask=edit_mask, **args)
    out, traj = model.sample(cond=cond.cuda(), text=text.cuda(), duration=duration.cuda(), lens=lens.cuda(), noise=noise,
                             edit_mask=None if edit_mask is None else edit_mask.cuda(), **args)
    torch.cuda.synchronize()
    assert out.shape == ref_out.shape
    err = (out.cpu() - ref_out).abs()
    assert float(err.mean()) <= 1e-2 and float(err.max()) <= 0.15, (kw, float(err.mean()), float(err.max()))


def test_cfm_sample_duration_rule_and_int_duration():
    """duration = max(max(#text, lens) + 1, duration) clamped to max_duration (cfm.py:132-136); int duration; list[str] text"""
    cfg = O.DiTConfig.tiny()
    model, sd = build_cfm(cfg, 0)
    model.vocab_char_map = {c: i for i, c in enumerate("abcdefghijklmnopqrstuvwxyz ")}
    cond = (torch.randn(1, 30, cfg.mel_dim) * 2 - 1.5).clamp(-11.5, 5)
    out, traj = model.sample(cond=cond.cuda(), text=["hello world"], duration=10, steps=2, cfg_strength=2.0, seed=1)
    assert out.shape == (1, 31, cfg.mel_dim)  # requested 10 < ref 30 -> 31
    out2, _ = model.sample(cond=cond.cuda(), text=["hello world"], duration=100, steps=2, cfg_strength=2.0, seed=1, max_duration=64)
    assert out2.shape == (1, 64, cfg.mel_dim)
    assert torch.equal(out[:, :30].cpu(), cond) and torch.isfinite(out2).all()


def test_repeated_calls_are_deterministic_and_independent():
    cfg = O.DiTConfig.tiny()
    model, sd = build_cfm(cfg, 0)
    a = synthetic_inputs(cfg, 2, 40, [96, 83], seed=1)
    b = synthetic_inputs(cfg, 1, 33, 70, seed=2)
    r = []
    for cond, text, duration, lens in (a, b, a):
        out, _ = model.sample(cond=cond.cuda(), text=text.cuda(), duration=duration.cuda(), lens=lens.cuda(), steps=2, cfg_strength=2.0,
                              sway_sampling_coef=-1.0, seed=0)
        r.append(out.clone())
    assert torch.equal(r[0], r[2])


@pytest.mark.parametrize("B,totals", [(1, 70), (2, [96, 83])])
def test_cfg_branches_as_concurrent_chains_are_bit_identical(B, totals, monkeypatch):
    """launch-bound shapes run the conditioned and the unconditioned forward of a CFG step as two concurrent chains inside the
    captured graph (DiTEngine.step_session) instead of one fused 2B-row batch: same kernels on the same rows, bit-identical"""
    cfg = O.DiTConfig.tiny()
    cond, text, duration, lens = synthetic_inputs(cfg, B, 33, totals, seed=5)
    outs = []
    for split in ("0", "1"):
        monkeypatch.setenv("F5B_SPLIT_CFG", split)
        monkeypatch.setenv("F5B_CUDA_GRAPH", "1")
        model, _ = build_cfm(cfg, 0)
        for _ in range(2):  # second call replays the cached session
            out, traj = model.sample(cond=cond.cuda(), text=text.cuda(), duration=duration.cuda(), lens=lens.cuda(), steps=4,
                                     cfg_strength=2.0, sway_sampling_coef=-1.0, seed=0)
        torch.cuda.synchronize()
        assert any(k[-1] == (split == "1") for k in model.transformer.engine()._sessions)
        outs.append((out.clone(), traj.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_inference_attention_dropout_vs_oracle_and_sampling():
    """the reference's SDPA dropout is active at inference too (model/modules.py:490); DiT.set_attn_dropout(p) switches
    it on: one DiT.forward against the oracle with the identical mask, then CFM.sample through the captured-graph path (mask stream
    driven by a device word per ODE step): deterministic per seed, different across seeds, different from p = 0"""
    cfg = O.DiTConfig.tiny()
    model, sd = build_cfm(cfg, 0)
    B, n = 2, 203
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, n, cfg.mel_dim, generator=g)
    cond = torch.randn(B, n, cfg.mel_dim, generator=g)
    text = torch.randint(0, cfg.text_num_embeds - 1, (B, 30), generator=g)
    time = torch.tensor(0.4)
    dit = model.transformer
    base = dit(x=x.cuda(), cond=cond.cuda(), text=text.cuda(), time=time.cuda(), drop_audio_cond=False, drop_text=False)
    dit.set_attn_dropout(0.5)
    out = dit(x=x.cuda(), cond=cond.cuda(), text=text.cuda(), time=time.cuda(), drop_audio_cond=False, drop_text=False,
              attn_dropout_seed=77)
    ref = O.dit_forward(sd, cfg, x, cond, text, time, False, False, None, dropout=(0.0, 77, 0.5))
    ref0 = O.dit_forward(sd, cfg, x, cond, text, time, False, False, None)
    torch.cuda.synchronize()
    assert maxabs(out, ref) <= 2e-2, maxabs(out, ref)
    # the mask matters, and it is the oracle's mask: the change it causes on the GPU is the change it causes in the oracle
    dg, do = (out - base).cpu().flatten().double(), (ref - ref0).flatten().double()
    cos = float((dg * do).sum() / (dg.norm() * do.norm()))
    assert float(do.abs().max()) > 2e-3 and cos > 0.8, (float(do.abs().max()), cos)
    dit.set_attn_dropout(0.2)
    # sampling: graph path (default) with the per-step device word
    cond_s, text_s, duration, lens = synthetic_inputs(cfg, 2, 40, [96, 83], seed=1)
    kw = dict(cond=cond_s.cuda(), text=text_s.cuda(), duration=duration.cuda(), lens=lens.cuda(), steps=4, cfg_strength=2.0,
              sway_sampling_coef=-1.0)
    a, _ = model.sample(seed=5, **kw)
    a = a.clone()
    b, _ = model.sample(seed=5, **kw)
    b = b.clone()
    dit.set_attn_dropout(0.0)
    z, _ = model.sample(seed=5, **kw)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.isfinite(a).all()
    gen = slice(40, None)
    d = (a - z)[:, gen].abs()
    assert float(d.max()) > 1e-3 and float(d.mean()) < 0.5, (float(d.max()), float(d.mean()))
    dit.set_attn_dropout(0.2)
    model.transformer.set_precision("tf32")
    with pytest.raises(Exception):
        dit(x=x.cuda(), cond=cond.cuda(), text=text.cuda(), time=time.cuda(), drop_audio_cond=False, drop_text=False)
    model.transformer.set_precision("bf16")
    dit.set_attn_dropout(0.0)


@pytest.mark.parametrize("B,H,n,std", [(2, 4, 700, 6.0), (2, 16, 1200, 4.0), (1, 2, 333, 2.5)])
def test_attention_forward_with_overflowing_speculative_exponentials(B, H, n, std):
    """Scores whose row maximum jumps by more than 2^128 from one key tile to the next: the exponentials the kernel issues speculatively
    against the old reference overflow to +inf, and the lazy-rescale path must recompute the tile exactly (SDPA handles any finite input;
    a variant that rescaled P from the row sum returned NaN here).  Checked against fp64 softmax attention, with the log-sum-exp."""
    from eraxvif5tts_b200 import ops
    dev = torch.device("cuda", 0)
    D = H * 64
    g = torch.Generator().manual_seed(0)
    qkv = (torch.randn(B * n, 3 * D, generator=g) * std).to(dev).bfloat16()
    q, k, v = (qkv[:, i * D:(i + 1) * D].double().reshape(B, n, H, 64) for i in range(3))
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * 0.125
    ref = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), v).reshape(B * n, D)
    out = torch.full((B * n, D), float("nan"), dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, H, n, dtype=torch.float32, device=dev)
    ops.attn_fwd_lse(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, lse, None, 0, B, H, n)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out.float()).all())
    assert float((out.double() - ref).norm() / ref.norm()) <= 4e-3
    ref_lse = torch.logsumexp(s, -1) * 1.4426950408889634
    assert float((lse.double() - ref_lse).abs().max()) <= 1e-3 * max(1.0, float(ref_lse.abs().max()))


def test_attention_forward_is_deterministic_with_many_ctas_per_sm():
    """regression: the P.V completion barrier used to be a single mbarrier whose phase advanced once per key tile, while the softmax
    threads waited on it only in the epilogue; a warp running a full tile ahead of the slowest one then saw the parity of a phase
    two tiles back and read O before the last products had landed (intermittent row errors of a few percent, only with >= 3 CTAs
    per SM over the kernel's lifetime).  Outputs must be bit-identical run to run and within tolerance."""
    from eraxvif5tts_b200 import ops
    dev = torch.device("cuda", 0)
    for B, H, n in ((1, 16, 4096), (1, 32, 2048), (3, 16, 1875)):
        D = H * 64
        g = torch.Generator().manual_seed(n)
        qkv = torch.randn(B * n, 3 * D, generator=g).to(dev).bfloat16()
        q, k, v = (qkv[:, i * D:(i + 1) * D].float().reshape(B, n, H, 64) for i in range(3))
        ref = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        ref = ref.transpose(1, 2).reshape(B * n, D)
        first = None
        for _ in range(6):
            out = torch.full((B * n, D), float("nan"), dtype=torch.bfloat16, device=dev)
            ops.attn_fwd(qkv[:, :D], qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, None, 0, B, H, n)
            torch.cuda.synchronize()
            assert float((out.float() - ref).abs().max()) <= 1e-2
            if first is None:
                first = out.clone()
            assert torch.equal(out, first)


@pytest.mark.parametrize("B,H,n,lens", [(1, 16, 940, None), (2, 4, 1000, [1000, 517]), (1, 2, 300, [65]), (2, 2, 256, [256, 1]),
                                        (1, 3, 1875, [1874])])
def test_attention_split_kv_matches_single_cta(B, H, n, lens):
    """split-KV (cluster of two CTAs per query tile, partials merged through distributed shared memory) against the single-CTA
    kernel on the same inputs: same online-softmax arithmetic on each half, one extra exp2 rescale at the merge"""
    from eraxvif5tts_b200 import ops, _lib as L
    raw = L.load()
    D = H * 64
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(B * n, 3 * D, generator=g).cuda().to(torch.bfloat16)
    lt = torch.tensor(lens, dtype=torch.int32, device="cuda") if lens else None
    outs = []
    try:
        for mode in (0, 1):
            raw.f5b_debug_attn_split(mode)
            out = torch.full((B * n, D), float("nan"), dtype=torch.bfloat16, device="cuda")
            ops.attn_fwd(qkv, qkv[:, D:], qkv[:, 2 * D:], 3 * D, out, lt, 0, B, H, n)
            torch.cuda.synchronize()
            outs.append(out.float())
    finally:
        raw.f5b_debug_attn_split(0)
    assert torch.isfinite(outs[1]).all()
    ref = outs[0]
    assert float((outs[1] - ref).abs().max()) <= 2e-2 * float(ref.abs().max()) + 1e-3   # bf16 output rounding of a re-associated sum
    assert float((outs[1] - ref).norm() / ref.norm()) < 3e-3
