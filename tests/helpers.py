"""Shared helpers of the GPU parity tests: build the product CFM / Vocos from the oracle's deterministic weights."""
import torch

from oracle import f5_oracle as O
from oracle.weights import make_dit_state_dict, make_vocos_state_dict


def build_cfm(cfg: O.DiTConfig, seed=0, method="euler", device="cuda"):
    from eraxvif5tts_b200.model import CFM, DiT
    sd = make_dit_state_dict(cfg, seed)
    tr = DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, dim_head=cfg.dim_head, ff_mult=cfg.ff_mult, mel_dim=cfg.mel_dim,
             text_num_embeds=cfg.text_num_embeds, text_dim=cfg.text_dim, text_mask_padding=cfg.text_mask_padding,
             conv_layers=cfg.conv_layers, pe_attn_head=cfg.pe_attn_head)
    model = CFM(transformer=tr, odeint_kwargs=dict(method=method),
                mel_spec_kwargs=dict(n_fft=1024, hop_length=256, win_length=1024, n_mel_channels=cfg.mel_dim,
                                     target_sample_rate=24000, mel_spec_type="vocos"))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("inv_freq" in k for k in missing), (missing, unexpected)
    return model.to(device).eval(), sd


def build_vocos(vc: O.VocosConfig, seed=1, device="cuda"):
    from eraxvif5tts_b200.vocoder import Vocos
    vsd = make_vocos_state_dict(vc, seed)
    voc = Vocos(n_mels=vc.n_mels, dim=vc.dim, intermediate_dim=vc.intermediate_dim, num_layers=vc.num_layers, n_fft=vc.n_fft,
                hop_length=vc.hop_length)
    missing, unexpected = voc.load_state_dict(vsd, strict=False)
    assert not unexpected and all("window" in k for k in missing), (missing, unexpected)
    return voc.to(device).eval(), vsd


def relerr(a, b):
    return float((a.float().cpu() - b.float().cpu()).abs().max() / (b.float().abs().max() + 1e-12))


def maxabs(a, b):
    return float((a.float().cpu() - b.float().cpu()).abs().max())
