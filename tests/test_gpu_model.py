"""GPU parity tests of the product path (through the C ABI) against the CPU oracle and the golden vectors that the REAL
reference modules produced (tests/golden/, generator oracle/gen_golden.py).

Tolerances (BASELINE.json north_star): bf16 tensor-core path -> max |error| <= 2e-2 on the DiT velocity field,
mean |error| <= 1e-2 on the final log-mel.  The oracle is fp32; weights are bf16-exact so only arithmetic differs."""
import numpy as np
import pytest
import torch

from oracle import f5_oracle as O
from oracle.weights import synthetic_inputs

from helpers import build_cfm, build_vocos, maxabs, relerr

pytestmark = pytest.mark.gpu

VEL_TOL = 2e-2
MEL_MEAN_TOL = 1e-2


def _cfg(d):
    return O.DiTConfig(**d)


@pytest.mark.parametrize("tag", ["tiny", "tiny_v1"])
def test_dit_forward_matches_reference_golden(golden, tag):
    g = golden(f"dit_{tag}.pt")
    cfg = _cfg(g["cfg"])
    model, sd = build_cfm(cfg, g["seed"])
    tr = model.transformer
    n = g["x"].shape[1]
    eng = tr.engine()
    te = eng.text_embed(g["text"], n, False)
    tu = eng.text_embed(g["text"], n, True)
    torch.cuda.synchronize()
    assert maxabs(te, g["text_cond"]) < 3e-2 * max(1.0, float(g["text_cond"].abs().max()))
    assert maxabs(tu, g["text_unc"]) < 3e-2 * max(1.0, float(g["text_unc"].abs().max()))
    dev = "cuda"
    for name, (da, dt, m) in dict(cond=(False, False, g["mask"]), uncond=(True, True, g["mask"]), nomask=(False, False, None)).items():
        out = tr(x=g["x"].to(dev), cond=g["cond"].to(dev), text=g["text"].to(dev), time=g["time"].to(dev), drop_audio_cond=da,
                 drop_text=dt, mask=None if m is None else m.to(dev))
        torch.cuda.synchronize()
        ref = g["out"][name]
        assert out.shape == ref.shape
        assert torch.isfinite(out).all()
        assert maxabs(out, ref) <= VEL_TOL, (name, maxabs(out, ref))


def _ref_noise(duration, mel_dim, seed):
    y0 = []
    for dur in duration.tolist():
        torch.manual_seed(seed)
        y0.append(torch.randn(int(dur), mel_dim))
    return torch.nn.utils.rnn.pad_sequence(y0, padding_value=0, batch_first=True)


@pytest.mark.parametrize("tag", ["tiny_b2", "tiny_b1", "tiny_mid", "tiny_dup"])
def test_cfm_sample_matches_reference_golden(golden, tag):
    g = golden(f"sample_{tag}.pt")
    cfg = _cfg(g["cfg"])
    model, sd = build_cfm(cfg, g["seed"], method=g["method"])
    noise = _ref_noise(g["duration"], cfg.mel_dim, g["sample_seed"])
    out, traj = model.sample(cond=g["cond"].cuda(), text=g["text"].cuda(), duration=g["duration"].cuda(), lens=g["lens"].cuda(),
                             steps=g["steps"], cfg_strength=g["cfg_strength"], sway_sampling_coef=g["sway"], seed=g["sample_seed"],
                             noise=noise, **g.get("extra", {}))
    torch.cuda.synchronize()
    assert out.shape == g["out"].shape and traj.shape[0] == g.get("traj_len", g["steps"] + 1)
    if not g.get("extra"):
        assert maxabs(traj[0], noise) == 0.0
    # first ODE state: one velocity evaluation scaled by dt
    assert maxabs(traj[1], g["traj_1"]) <= VEL_TOL
    err = (out.cpu() - g["out"]).abs()
    assert float(err.mean()) <= MEL_MEAN_TOL, float(err.mean())
    assert float(err.max()) <= 0.15, float(err.max())


@pytest.mark.parametrize("name,cfg,B,n,ragged", [
    ("base_width", O.DiTConfig(depth=2), 2, 200, True),
    ("small_width", O.DiTConfig(dim=768, heads=12, depth=2), 2, 150, True),
    ("v1_allheads", O.DiTConfig(depth=1, pe_attn_head=None, text_mask_padding=True), 1, 130, False),
])
def test_dit_forward_real_widths_vs_oracle(name, cfg, B, n, ragged):
    model, sd = build_cfm(cfg, 0)
    cond, text, _, _ = synthetic_inputs(cfg, B, n, n, seed=5)
    cond[:, n // 2:] = 0
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, n, cfg.mel_dim, generator=g)
    time = torch.tensor(0.41)
    mask = None
    if ragged:
        lens = torch.tensor([n, n - 37][:B])
        mask = torch.arange(n)[None, :] < lens[:, None]
    for da, dt in ((False, False), (True, True)):
        ref = O.dit_forward(sd, cfg, x, cond, text, time, da, dt, mask)
        out = model.transformer(x=x.cuda(), cond=cond.cuda(), text=text.cuda(), time=time.cuda(), drop_audio_cond=da, drop_text=dt,
                                mask=None if mask is None else mask.cuda())
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        assert maxabs(out, ref) <= VEL_TOL * max(1.0, float(ref.abs().max())), (name, da, maxabs(out, ref), float(ref.abs().max()))


def test_cfm_sample_base_width_fused_cfg_vs_oracle():
    cfg = O.DiTConfig(depth=2)
    model, sd = build_cfm(cfg, 0)
    cond, text, duration, lens = synthetic_inputs(cfg, 2, 60, [150, 131], seed=1234)
    noise = _ref_noise(duration, cfg.mel_dim, 0)
    ref_out, ref_traj = O.cfm_sample(sd, cfg, cond, text, duration, lens=lens, steps=4, cfg_strength=2.0, sway_sampling_coef=-1.0, seed=0)
    out, traj = model.sample(cond=cond.cuda(), text=text.cuda(), duration=duration.cuda(), lens=lens.cuda(), steps=4, cfg_strength=2.0,
                             sway_sampling_coef=-1.0, seed=0, noise=noise)
    torch.cuda.synchronize()
    assert maxabs(traj[1], ref_traj[1]) <= VEL_TOL
    err = (out.cpu() - ref_out).abs()
    assert float(err.mean()) <= MEL_MEAN_TOL and float(err.max()) <= 0.15, (float(err.mean()), float(err.max()))


def test_melspec_matches_reference_golden(golden):
    from eraxvif5tts_b200.model import MelSpec
    g = golden("melspec.pt")
    mel = MelSpec().cuda()(g["wav"].cuda())
    torch.cuda.synchronize()
    assert mel.shape == g["mel"].shape
    assert maxabs(mel, g["mel"]) < 2e-3
    assert float((mel.cpu() - g["mel"]).abs().mean()) < 1e-4


@pytest.mark.parametrize("vc,T", [(O.VocosConfig.tiny(), 12), (O.VocosConfig(), 40)])
def test_vocos_decode_vs_oracle(vc, T):
    voc, vsd = build_vocos(vc)
    g = torch.Generator().manual_seed(2)
    mel = (torch.randn(2, vc.n_mels, T, generator=g) * 2.0 - 1.5).clamp(-11.5, 5.0)
    ref = O.vocos_decode(vsd, vc, mel)
    wav = voc.decode(mel.cuda())
    torch.cuda.synchronize()
    assert wav.shape == ref.shape == (2, 256 * (T - 1))
    assert torch.isfinite(wav).all()
    scale = float(ref.abs().max())
    assert maxabs(wav, ref) <= 5e-2 * scale, (maxabs(wav, ref), scale)
    assert float((wav.cpu() - ref).abs().mean()) <= 1e-2 * scale


def test_wrapper_generate_end_to_end(tmp_path):
    """F5TTSWrapper.generate: chunking, duration rule, sample, vocoder, rms rescale, cross-fade; serial vs batched chunks."""
    from eraxvif5tts_b200.infer import F5TTSWrapper
    yaml_path = tmp_path / "custom_tiny.yaml"
    yaml_path.write_text("model:\n  backbone: DiT\n  arch:\n    dim: 128\n    depth: 2\n    heads: 2\n    ff_mult: 2\n    text_dim: 64\n"
                         "    text_mask_padding: False\n    conv_layers: 2\n    pe_attn_head: 1\n")
    vocab = {c: i for i, c in enumerate(" abcdefghijklmnopqrstuvwxyz.,!?")}
    torch.manual_seed(0)
    w = F5TTSWrapper(model_name=str(yaml_path), vocab_char_map=vocab, device="cuda")
    # give the zero-initialised AdaLN / proj_out tensors values so the model output is not identically zero
    with torch.no_grad():
        for p in w.model.parameters():
            if float(p.abs().max()) == 0.0:
                p.normal_(0, 0.02)
    w.model.transformer.invalidate()
    with pytest.raises(ValueError):
        w.generate("hello")
    ref = 0.05 * torch.randn(24000 * 2)
    w.preprocess_reference(ref, "this is a reference.", sample_rate=24000)
    assert w.ref_audio_len == w.ref_audio_processed.shape[-1] // 256
    text = "the quick brown fox jumps over the lazy dog. " * 6
    wave, sr = w.generate(text, nfe_step=2, return_numpy=True, seed=0)
    assert sr == 24000 and wave.ndim == 1 and wave.size > 24000 and bool((wave == wave).all())
    wave_b, _ = w.generate(text, nfe_step=2, return_numpy=True, seed=0, batch_chunks=True)
    assert abs(wave_b.size - wave.size) <= 256 * 4
    # on-device cross-fade / PCM packing against the host fold of the same run (same seed -> same chunk waves)
    wave_d, _ = w.generate(text, nfe_step=2, return_numpy=True, seed=0, device_crossfade=True)
    assert wave_d.dtype == np.float32 and wave_d.shape == wave.shape
    assert np.array_equal(wave_d, wave.astype(np.float32))
    pcm, _ = w.generate(text, nfe_step=2, return_numpy=True, seed=0, return_pcm16=True)
    assert pcm.dtype == np.int16 and np.array_equal(pcm, np.int16(np.clip(wave_d * np.float32(32767), -32768, 32767)))
    out = w.generate("short text here.", output_path=str(tmp_path / "o.wav"), nfe_step=2)
    assert out.endswith("o.wav") and (tmp_path / "o.wav").stat().st_size > 1000
    # utils_infer.infer_process / infer_batch_process (the call sites of the reference's api.py / CLI / socket server)
    from eraxvif5tts_b200.infer import utils_infer as UI
    ref_in = (ref.unsqueeze(0), 24000)
    wave_i, sr_i, spec_i = UI.infer_process(ref_in, "this is a reference. ", text, w.model, w.vocoder, nfe_step=2, batch_chunks=False,
                                            show_info=lambda *_: None)
    assert sr_i == 24000 and wave_i.ndim == 1 and np.isfinite(wave_i).all() and spec_i.shape[0] == 100
    wave_j, _, spec_j = UI.infer_process(ref_in, "this is a reference. ", text, w.model, w.vocoder, nfe_step=2, batch_chunks=True,
                                         show_info=lambda *_: None)
    assert abs(wave_j.size - wave_i.size) <= 256 * 4 and abs(spec_j.shape[1] - spec_i.shape[1]) <= 4
    chunks = chunk_chars = UI.chunk_text(text, max_chars=int(len("this is a reference. ".encode()) / 2.0 * (22 - 2.0)))
    pieces = list(UI.infer_batch_process(ref_in, "this is a reference. ", chunks, w.model, w.vocoder, nfe_step=2, streaming=True,
                                         chunk_size=2048))
    assert all(sr_ == 24000 and 0 < len(c_) <= 2048 for c_, sr_ in pieces)
    cfs = int(0.15 * 24000)
    assert sum(len(c_) for c_, _ in pieces) == wave_i.size + (len(chunks) - 1) * cfs
    assert next(UI.infer_batch_process(ref_in, "this is a reference. ", [], w.model, w.vocoder))[0] is None
    # duration-predictor variant of the wrapper (model/f5tts_wrapper-dur_pred.py): chunk duration = ref frames + predicted frames / speed
    from eraxvif5tts_b200.model import DurationPredictor
    torch.manual_seed(1)
    dp = DurationPredictor(len(vocab), 32, 16, 3, 0.5)
    with torch.no_grad():
        dp.proj.bias.fill_(1.0)  # exp(1) = 2.7 frames per character
    w.attach_duration_predictor(dp)
    chunk = "hello there, general."
    d = w.calculate_duration_with_predictor(chunk, 1.0)
    from oracle import align_oracle as A
    from eraxvif5tts_b200.model.utils import list_str_to_idx
    ids = list_str_to_idx([chunk], vocab)
    want = w.ref_audio_len + int(torch.exp(torch.clamp(A.duration_predictor(dp.cpu().state_dict(), ids, torch.ones_like(ids), 1), -20, 20)).sum())
    dp.cuda()
    assert abs(d - want) <= 1
    wave_p, _ = w.generate(chunk, nfe_step=2, return_numpy=True, seed=0)
    assert abs(wave_p.size - 256 * (d - w.ref_audio_len - 1)) <= 256 * 2
    wave_r, _ = w.generate(chunk, nfe_step=2, return_numpy=True, seed=0, use_duration_predictor=False)
    assert wave_r.size != wave_p.size


def test_device_crossfade_fold_matches_numpy_fold():
    """ops.crossfade_concat against the reference's numpy fold (f5tts_wrapper.py:549-575), including chunks shorter than the fade
    (chained blends) and cfs = 1 / 0: the fp64 blend rounded to fp32 once must equal the numpy result cast to fp32"""
    from eraxvif5tts_b200 import ops
    g = torch.Generator().manual_seed(3)
    for lens, cf in (((5000, 9000, 4100), 3600), ((5000, 7000, 4100), 3600), ((900, 400, 120, 3000), 500), ((10, 2, 7), 1), ((64, 64), 0),
                     ((3000,), 3600), ((2000, 50, 60, 2000), 3600), ((4000, 1000, 1000, 1000, 2500), 500)):
        waves = [torch.randn(n, generator=g) * 0.3 for n in lens]
        final = waves[0].numpy()
        for nxt in waves[1:]:
            nxt = nxt.numpy()
            cfs = min(cf, len(final), len(nxt))
            if cfs <= 0:
                final = np.concatenate([final, nxt])
                continue
            overlap = final[-cfs:] * np.linspace(1, 0, cfs) + nxt[:cfs] * np.linspace(0, 1, cfs)
            final = np.concatenate([final[:-cfs], overlap, nxt[cfs:]])
        got = ops.crossfade_concat([w_.cuda() for w_ in waves], cf).cpu().numpy()
        assert got.shape == final.shape, (lens, cf)
        if all(n >= 2 * cf for n in lens[1:-1]):  # no sample is blended twice: one fp64 blend, rounded to fp32 once
            assert np.array_equal(got, final.astype(np.float32)), (lens, cf)
        else:  # chained blends: the device keeps the running wave in fp32 between folds, numpy in fp64
            assert float(np.abs(got - final).max()) <= 1e-6, (lens, cf)
    x = torch.tensor([0.0, 0.5, -0.5, 0.99999, -1.0, 1.0, 1.7, -2.0, 3.0517578e-05, -3.0517578e-05])
    want = np.int16(np.clip(x.numpy() * np.float32(32767), -32768, 32767))
    assert np.array_equal(ops.pcm16(x.cuda()).cpu().numpy(), want)


def test_cuda_graph_step_matches_eager_launches(monkeypatch):
    """the captured-graph ODE step (launch-bound small batches) must be bit-identical to the eager launch sequence,
    on first use (capture) and on replay with new inputs"""
    cfg = O.DiTConfig.tiny()
    model, sd = build_cfm(cfg, 0)
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("F5B_CUDA_GRAPH", mode)
        res = []
        for seed in (1234, 77):
            cond, text, duration, lens = synthetic_inputs(cfg, 2, 40, [96, 83], seed=seed)
            out, traj = model.sample(cond=cond.cuda(), text=text.cuda(), duration=duration.cuda(), lens=lens.cuda(), steps=3,
                                     cfg_strength=2.0, sway_sampling_coef=-1.0, seed=0)
            res.append((out.clone(), traj.clone()))
        outs[mode] = res
    torch.cuda.synchronize()
    for (o0, t0), (o1, t1) in zip(outs["0"], outs["1"]):
        assert torch.equal(o0, o1) and torch.equal(t0, t1)
    assert not torch.equal(outs["1"][0][0], outs["1"][1][0])


def test_cfm_forward_loss_vs_oracle():
    """CFM.forward (flow-matching loss, forward only) with the random draws pinned: per-row times, span mask, CFG drops"""
    cfg = O.DiTConfig.tiny()
    model, sd = build_cfm(cfg, 0)
    g = torch.Generator().manual_seed(5)
    B, n = 3, 77
    x1 = (torch.randn(B, n, cfg.mel_dim, generator=g) * 2 - 1.5).clamp(-11.5, 5)
    x0 = torch.randn(B, n, cfg.mel_dim, generator=g)
    time = torch.rand(B, generator=g)
    text = torch.randint(0, cfg.text_num_embeds, (B, 12), generator=g)
    lens = torch.tensor([77, 60, 33])
    span = torch.zeros(B, n, dtype=torch.bool)
    span[0, 20:70] = True
    span[1, 5:55] = True
    span[2, 10:30] = True
    for da, dt in ((False, False), (True, False), (True, True)):
        ref_loss, ref_cond, ref_pred = O.cfm_loss(sd, cfg, x1, text, span & (torch.arange(n)[None] < lens[:, None]), x0, time, da, dt)
        loss, cond, pred = model(x1.cuda(), text.cuda(), lens=lens.cuda(),
                                 draws=dict(rand_span_mask=span, x0=x0, time=time, drop_audio_cond=da, drop_text=dt))
        torch.cuda.synchronize()
        assert maxabs(cond, ref_cond) == 0.0
        assert maxabs(pred, ref_pred) <= VEL_TOL
        assert abs(float(loss) - float(ref_loss)) <= 2e-2 * float(ref_loss), (float(loss), float(ref_loss))


def test_dit_forward_full_sequence_length_vs_oracle_on_gpu():
    """BASELINE-size sequences (cfg-2: 1875 frames, Base width, 16 heads; 4 ragged utterances x {cond, uncond}): thousands of attention
    CTAs and several per SM over a launch — the regime the tiny-config goldens cannot reach.  The oracle (fp32 torch restatement of
    the reference) runs on the same GPU in fp32; bar: velocity max-abs <= 2e-2 (north_star's bf16 tolerance)."""
    cfg = O.DiTConfig(depth=4)
    model, sd = build_cfm(cfg, 0)
    dev = torch.device("cuda", 0)
    B, n = 4, 1875
    cond, text, _, _ = synthetic_inputs(cfg, B, n, n, seed=9)
    cond[:, 563:] = 0
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, n, cfg.mel_dim, generator=g)
    time = torch.tensor(0.37)
    lens = torch.tensor([1875, 1500, 1874, 700])
    mask = torch.arange(n)[None, :] < lens[:, None]
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    rope_cpu = O.rotary_freqs
    O.rotary_freqs = lambda n_, d=64, theta=10000.0: rope_cpu(n_, d, theta).to(dev)
    try:
        with torch.no_grad():
            for da, dt in ((False, False), (True, True)):
                ref = O.dit_forward(sd_dev, cfg, x.to(dev), cond.to(dev), text.to(dev), time.to(dev), da, dt, mask.to(dev))
                outs = []
                for _ in range(2):
                    out = model.transformer(x=x.to(dev), cond=cond.to(dev), text=text.to(dev), time=time.to(dev), drop_audio_cond=da,
                                            drop_text=dt, mask=mask.to(dev))
                    torch.cuda.synchronize()
                    outs.append(out.clone())
                assert torch.isfinite(outs[0]).all()
                assert torch.equal(outs[0], outs[1])  # bit-exact run to run
                assert maxabs(outs[0], ref) <= VEL_TOL * max(1.0, float(ref.abs().max())), (da, maxabs(outs[0], ref), float(ref.abs().max()))
    finally:
        O.rotary_freqs = rope_cpu
