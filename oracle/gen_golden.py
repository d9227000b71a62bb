"""Generate tests/golden/*.pt by running the REAL reference modules (oracle/ref_shim.py) in the build
container.  Run:  python -m oracle.gen_golden      (TEST INFRASTRUCTURE)

Each fixture stores the inputs, the reference's outputs and the digest of the deterministic weights
(oracle/weights.py) it was produced with; weights themselves are regenerated from (config, seed).
"""
from __future__ import annotations

import os
import sys
from dataclasses import asdict

import torch

from . import ref_shim
from .f5_oracle import DiTConfig
from .weights import make_dit_state_dict, state_dict_digest, synthetic_inputs

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    torch.save(obj, path)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)")


def gen_melspec():
    ref = ref_shim.load()
    g = torch.Generator().manual_seed(7)
    wav = 0.1 * torch.randn(2, 256 * 37 + 19, generator=g)
    mel = ref.modules.MelSpec()(wav)
    _save("melspec.pt", dict(wav=wav, mel=mel))


def gen_dit(cfg: DiTConfig, tag: str, batch=2, n=96, seed=0):
    sd = make_dit_state_dict(cfg, seed)
    model = ref_shim.build_reference_cfm(cfg, sd)
    tr = model.transformer
    g = torch.Generator().manual_seed(11)
    x = torch.randn(batch, n, cfg.mel_dim, generator=g)
    cond, text, _, _ = synthetic_inputs(cfg, batch, n, n, seed=5)
    cond[:, n // 2:] = 0.0
    time = torch.tensor(0.37)
    lens = torch.tensor([n, n - 17][:batch])
    mask = torch.arange(n)[None, :] < lens[:, None]
    cases = {}
    with torch.no_grad():
        for name, (da, dt, m) in dict(cond=(False, False, mask), uncond=(True, True, mask),
                                      nomask=(False, False, None)).items():
            cases[name] = tr(x=x, cond=cond, text=text, time=time, drop_audio_cond=da, drop_text=dt, mask=m)
        t_emb = tr.time_embed(time.repeat(batch))
        text_cond = tr.text_embed(text, n, drop_text=False)
        text_unc = tr.text_embed(text, n, drop_text=True)
        h0 = tr.input_embed(x, cond, text_cond, drop_audio_cond=False)
        rope = tr.rotary_embed.forward_from_seq_len(n)
        h1 = tr.transformer_blocks[0](h0, t_emb, mask=mask, rope=rope)
    _save(f"dit_{tag}.pt", dict(cfg=asdict(cfg), seed=seed, digest=state_dict_digest(sd), x=x, cond=cond, text=text,
                                time=time, mask=mask, out=cases, t_emb=t_emb, text_cond=text_cond,
                                text_unc=text_unc, h0=h0, h1=h1))


def gen_sample(cfg: DiTConfig, tag: str, batch, ref_frames, total, steps, method="euler", seed=0):
    sd = make_dit_state_dict(cfg, seed)
    model = ref_shim.build_reference_cfm(cfg, sd, method=method)
    cond, text, duration, lens = synthetic_inputs(cfg, batch, ref_frames, total, seed=1234)
    with torch.no_grad():
        out, traj = model.sample(cond=cond, text=text, duration=duration, lens=lens, steps=steps, cfg_strength=2.0,
                                 sway_sampling_coef=-1.0, seed=0)
    _save(f"sample_{tag}.pt", dict(cfg=asdict(cfg), seed=seed, digest=state_dict_digest(sd), cond=cond, text=text,
                                   duration=duration, lens=lens, steps=steps, method=method, cfg_strength=2.0,
                                   sway=-1.0, sample_seed=0, out=out, traj_last=traj[-1], traj_1=traj[1]))


def main():
    if not ref_shim.available():
        sys.exit("reference tree not present; golden vectors can only be generated in the build container")
    torch.set_num_threads(os.cpu_count() or 1)
    gen_melspec()
    gen_dit(DiTConfig.tiny(), "tiny")
    gen_dit(DiTConfig.tiny(pe_attn_head=None, text_mask_padding=True), "tiny_v1")
    gen_sample(DiTConfig.tiny(), "tiny_b2", batch=2, ref_frames=40, total=[96, 83], steps=4)
    gen_sample(DiTConfig.tiny(), "tiny_b1", batch=1, ref_frames=40, total=90, steps=4)
    gen_sample(DiTConfig.tiny(), "tiny_mid", batch=2, ref_frames=40, total=[70, 64], steps=3, method="midpoint")


if __name__ == "__main__":
    main()
