"""Pins the CPU oracle (oracle/f5_oracle.py) against golden vectors produced by the REAL reference
modules (oracle/gen_golden.py), and — when /root/reference is present — against the reference live."""
import pytest
import torch

from oracle import f5_oracle as O
from oracle import ref_shim
from oracle.weights import make_dit_state_dict, state_dict_digest, synthetic_inputs


def _cfg(d):
    return O.DiTConfig(**d)


def test_melspec_matches_reference(golden):
    g = golden("melspec.pt")
    mel = O.melspec(g["wav"])
    assert mel.shape == g["mel"].shape == (2, 100, 1 + g["wav"].shape[1] // 256)
    assert (mel - g["mel"]).abs().max() < 2e-4


def test_mel_filterbank_matches_torchaudio():
    import torchaudio
    fb = torchaudio.functional.melscale_fbanks(513, 0.0, 12000.0, 100, 24000, norm=None, mel_scale="htk")
    assert (O.mel_filterbank() - fb).abs().max() < 1e-6


@pytest.mark.parametrize("tag", ["tiny", "tiny_v1"])
def test_dit_forward_matches_reference(golden, tag):
    g = golden(f"dit_{tag}.pt")
    cfg = _cfg(g["cfg"])
    sd = make_dit_state_dict(cfg, g["seed"])
    assert state_dict_digest(sd) == g["digest"]
    n = g["x"].shape[1]
    assert (O.timestep_embedding(sd, g["time"].repeat(2)) - g["t_emb"]).abs().max() < 1e-4
    te = O.text_embedding(sd, cfg, g["text"], n, False)
    tu = O.text_embedding(sd, cfg, g["text"], n, True)
    assert (te - g["text_cond"]).abs().max() < 1e-4
    assert (tu - g["text_unc"]).abs().max() < 1e-4
    h0 = O.input_embedding(sd, cfg, g["x"], g["cond"], te, False)
    assert (h0 - g["h0"]).abs().max() < 1e-4
    h1 = O.dit_block(sd, cfg, 0, h0, g["t_emb"], g["mask"], O.rotary_freqs(n, cfg.dim_head))
    assert (h1 - g["h1"]).abs().max() < 1e-4
    for name, (da, dt, m) in dict(cond=(False, False, g["mask"]), uncond=(True, True, g["mask"]),
                                  nomask=(False, False, None)).items():
        out = O.dit_forward(sd, cfg, g["x"], g["cond"], g["text"], g["time"], da, dt, m)
        ref = g["out"][name]
        assert (out - ref).abs().max() <= 1e-3 * ref.abs().max(), name


@pytest.mark.parametrize("tag", ["tiny_b2", "tiny_b1", "tiny_mid", "tiny_dup"])
def test_cfm_sample_matches_reference(golden, tag):
    g = golden(f"sample_{tag}.pt")
    cfg = _cfg(g["cfg"])
    sd = make_dit_state_dict(cfg, g["seed"])
    assert state_dict_digest(sd) == g["digest"]
    out, traj = O.cfm_sample(sd, cfg, g["cond"], g["text"], g["duration"], lens=g["lens"], steps=g["steps"],
                             cfg_strength=g["cfg_strength"], sway_sampling_coef=g["sway"], seed=g["sample_seed"],
                             method=g["method"], **g.get("extra", {}))
    assert out.shape == g["out"].shape and traj.shape[0] == g.get("traj_len", g["steps"] + 1)
    assert (traj[1] - g["traj_1"]).abs().max() < 1e-3
    assert (out - g["out"]).abs().max() < 2e-3


def test_istft_matches_torch():
    g = torch.Generator().manual_seed(3)
    spec = torch.randn(2, 513, 21, generator=g) + 1j * torch.randn(2, 513, 21, generator=g)
    ref = torch.istft(spec, 1024, 256, 1024, torch.hann_window(1024), center=True)
    got = O.istft_center(spec)
    assert got.shape == ref.shape == (2, 256 * 20)
    assert (got - ref).abs().max() < 1e-4


def test_vocos_shapes():
    vc = O.VocosConfig.tiny()
    from oracle.weights import make_vocos_state_dict
    vsd = make_vocos_state_dict(vc)
    wav = O.vocos_decode(vsd, vc, torch.randn(2, 100, 12))
    assert wav.shape == (2, 256 * 11) and torch.isfinite(wav).all()


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
def test_oracle_vs_live_reference_sample():
    cfg = O.DiTConfig.tiny(depth=1)
    sd = make_dit_state_dict(cfg, 3)
    model = ref_shim.build_reference_cfm(cfg, sd)
    cond, text, duration, lens = synthetic_inputs(cfg, 2, 30, [61, 50], seed=9)
    with torch.no_grad():
        ref_out, ref_traj = model.sample(cond=cond, text=text, duration=duration, lens=lens, steps=2,
                                         cfg_strength=2.0, sway_sampling_coef=-1.0, seed=0)
    out, traj = O.cfm_sample(sd, cfg, cond, text, duration, lens=lens, steps=2, cfg_strength=2.0,
                             sway_sampling_coef=-1.0, seed=0)
    assert (out - ref_out).abs().max() < 1e-3


@pytest.mark.parametrize("tag", ["tiny_cond", "tiny_uncond"])
def test_cfm_loss_and_autograd_match_reference_cfm_forward(tag):
    """oracle.cfm_loss (the checker of the GPU training step) against the reference's own CFM.forward + loss.backward() on the draws
    that call made (tests/golden/cfm_forward_*.pt, oracle/gen_golden.py::gen_cfm_forward): loss, cond, pred, the norm of every
    parameter gradient and the leading entries of a representative subset."""
    import os
    d = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"cfm_forward_{tag}.pt"))
    cfg = O.DiTConfig(**d["cfg"])
    sd = make_dit_state_dict(cfg, d["seed"])
    leaf = {k: v.clone().float().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    full = dict(sd)
    full.update(leaf)
    loss, cond, pred = O.cfm_loss(full, cfg, d["x1"], d["text"], d["span"], d["x0"], d["time"], d["drop_audio_cond"], d["drop_text"])
    assert abs(float(loss) - float(d["loss"])) <= 1e-5 * float(d["loss"])
    assert torch.equal(cond, d["cond"])
    assert float((pred - d["pred"]).abs().max()) <= 1e-4
    loss.backward()
    for k, n in d["grad_norms"].items():
        g = leaf[k].grad
        if n < 1e-12:
            assert g is None or float(g.norm()) < 1e-9, k
            continue
        assert g is not None and abs(float(g.norm()) - n) <= 2e-3 * n + 1e-9, (k, float(g.norm()), n)
    for k, r in d["grads"].items():
        g = leaf[k].grad.flatten()[: r.numel()]
        assert float((g - r).norm()) <= 2e-3 * float(r.norm()) + 1e-9, k
