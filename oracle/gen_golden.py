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


def gen_sample(cfg: DiTConfig, tag: str, batch, ref_frames, total, steps, method="euler", seed=0, **extra):
    """`extra`: further keyword arguments of the reference's CFM.sample (duplicate_test / t_inter, cfm.py:96-97), stored in the file"""
    sd = make_dit_state_dict(cfg, seed)
    model = ref_shim.build_reference_cfm(cfg, sd, method=method)
    cond, text, duration, lens = synthetic_inputs(cfg, batch, ref_frames, total, seed=1234)
    with torch.no_grad():
        out, traj = model.sample(cond=cond, text=text, duration=duration, lens=lens, steps=steps, cfg_strength=2.0,
                                 sway_sampling_coef=-1.0, seed=0, **extra)
    _save(f"sample_{tag}.pt", dict(cfg=asdict(cfg), seed=seed, digest=state_dict_digest(sd), cond=cond, text=text,
                                   duration=duration, lens=lens, steps=steps, method=method, cfg_strength=2.0,
                                   sway=-1.0, sample_seed=0, out=out, traj_last=traj[-1], traj_1=traj[1], extra=extra,
                                   traj_len=traj.shape[0]))


GRAD_KEYS = ("transformer.proj_out.weight", "transformer.norm_out.linear.weight", "transformer.transformer_blocks.0.attn.to_q.weight",
             "transformer.transformer_blocks.0.attn.to_k.bias", "transformer.transformer_blocks.1.attn.to_out.0.weight",
             "transformer.transformer_blocks.0.attn_norm.linear.weight", "transformer.transformer_blocks.1.ff.ff.0.0.weight",
             "transformer.transformer_blocks.1.ff.ff.2.bias", "transformer.input_embed.proj.weight",
             "transformer.input_embed.conv_pos_embed.conv1d.0.weight", "transformer.text_embed.text_embed.weight",
             "transformer.text_embed.text_blocks.0.grn.gamma", "transformer.time_embed.time_mlp.0.weight")


def gen_cfm_forward(cfg: DiTConfig, tag: str, drop_audio_cond: bool, drop_text: bool, batch=3, n=88, seed=0):
    """The reference's own CFM.forward (cfm.py:210-283) + loss.backward(), with its internal random draws RECORDED (span mask, x0,
    time) and its two `random()` calls forced, so the oracle's cfm_loss can be pinned on identical draws.  eval() mode: the DiT's
    nn.Dropout sites are off (parity is defined at dropout 0).  Stored: loss, cond, pred, the L2 norm of EVERY parameter
    gradient and the first 4096 entries of a representative subset of them."""
    sd = make_dit_state_dict(cfg, seed)
    model = ref_shim.build_reference_cfm(cfg, sd)
    cfmmod = ref_shim.load().cfm
    g = torch.Generator().manual_seed(21)
    x1 = (torch.randn(batch, n, cfg.mel_dim, generator=g) * 2 - 1.5).clamp(-11.5, 5)
    text = torch.randint(0, cfg.text_num_embeds, (batch, 14), generator=g)
    text[1, 9:] = -1
    lens = torch.tensor([n, n - 13, n - 40][:batch])
    rec = {}
    orig_mask, orig_randn_like, orig_rand, orig_random = cfmmod.mask_from_frac_lengths, torch.randn_like, torch.rand, cfmmod.random
    draws = iter([0.0 if drop_audio_cond else 0.99, 0.0 if drop_text else 0.99])

    def mask_fn(seq_len, frac):
        m = orig_mask(seq_len, frac)
        rec["span_raw"] = m.clone()
        return m

    def randn_like(t, *a, **k):
        r = orig_randn_like(t, *a, **k)
        if t.shape == x1.shape:
            rec["x0"] = r.clone()
        return r

    def rand(*a, **k):
        r = orig_rand(*a, **k)
        if tuple(r.shape) == (batch,):
            rec["time"] = r.clone()
        return r

    cfmmod.mask_from_frac_lengths, torch.randn_like, torch.rand, cfmmod.random = mask_fn, randn_like, rand, lambda: next(draws)
    try:
        torch.manual_seed(123)
        for p_ in model.parameters():
            p_.requires_grad_(True)
        loss, cond, pred = model(x1, text, lens=lens)
        loss.backward()
    finally:
        cfmmod.mask_from_frac_lengths, torch.randn_like, torch.rand, cfmmod.random = orig_mask, orig_randn_like, orig_rand, orig_random
    mask = torch.arange(n)[None, :] < lens[:, None]
    grads = {k: p_.grad.detach().clone() for k, p_ in model.named_parameters() if p_.grad is not None}
    assert all(k in grads for k in GRAD_KEYS), [k for k in GRAD_KEYS if k not in grads]
    _save(f"cfm_forward_{tag}.pt", dict(cfg=asdict(cfg), seed=seed, digest=state_dict_digest(sd), x1=x1, text=text, lens=lens,
                                        span=rec["span_raw"] & mask, x0=rec["x0"], time=rec["time"], drop_audio_cond=drop_audio_cond,
                                        drop_text=drop_text, loss=loss.detach(), cond=cond.detach(), pred=pred.detach(),
                                        grads={k: grads[k].flatten()[:4096].clone() for k in GRAD_KEYS}, grad_norms={k: float(v.norm()) for k, v in grads.items()}))


def gen_dup():
    # duplicate_test corner (cfm.py:139-140, 188-191): y0 blended with the shifted reference mel, t_start = t_inter, fewer steps
    gen_sample(DiTConfig.tiny(), "tiny_dup", batch=2, ref_frames=30, total=[96, 83], steps=5, duplicate_test=True, t_inter=0.2)


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
    gen_dup()
    gen_cfm_forward(DiTConfig.tiny(), "tiny_cond", False, False)
    gen_cfm_forward(DiTConfig.tiny(), "tiny_uncond", True, True)


if __name__ == "__main__":
    if "--only-dup" in sys.argv:
        gen_dup()
    else:
        main()
