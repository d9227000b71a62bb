"""Deterministic synthetic weights with the reference's state_dict key names
(TEST / BENCH INFRASTRUCTURE — checkpoints are not available offline).

Values are a pure function of (config, seed) drawn from a CPU torch.Generator in a fixed key order, so
the build container (where the real reference is loaded to make the golden vectors) and the GPU box
(where /root/reference does not exist) construct bit-identical weights.  Zero-initialised reference
tensors (AdaLN linears, norm_out, proj_out — model/backbones/dit.py:162-172) are drawn N(0, 0.02^2)
instead, otherwise a random-init DiT returns exactly 0 (documented oracle adjustment #2).
With `bf16_exact=True` every value is rounded to a bf16-representable fp32 so that the CUDA path's
bf16 weight copy is lossless and parity measures arithmetic, not weight quantisation.
"""
from __future__ import annotations

import hashlib
import math

import torch

from .f5_oracle import DiTConfig, VocosConfig


def _draw(g, shape, std):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def dit_key_shapes(cfg: DiTConfig):
    """(key, shape, std) in a fixed order.  Shapes per SURVEY.md §10."""
    D, T = cfg.dim, cfg.text_dim
    inner = cfg.heads * cfg.dim_head
    ks = cfg.conv_pos_kernel
    cpg = D // cfg.conv_pos_groups
    out = []

    def lin(name, o, i, wstd=None, bstd=0.02):
        out.append((name + ".weight", (o, i), wstd if wstd is not None else 1.0 / math.sqrt(i)))
        out.append((name + ".bias", (o,), bstd))

    p = "transformer."
    lin(p + "time_embed.time_mlp.0", D, 256)
    lin(p + "time_embed.time_mlp.2", D, D)
    out.append((p + "text_embed.text_embed.weight", (cfg.text_num_embeds + 1, T), 1.0))
    for j in range(cfg.conv_layers):
        q = f"{p}text_embed.text_blocks.{j}."
        out.append((q + "dwconv.weight", (T, 1, 7), 1.0 / math.sqrt(7)))
        out.append((q + "dwconv.bias", (T,), 0.02))
        out.append((q + "norm.weight", (T,), None))          # 1 + N(0, .1)
        out.append((q + "norm.bias", (T,), 0.02))
        lin(q + "pwconv1", 2 * T, T)
        out.append((q + "grn.gamma", (1, 1, 2 * T), 0.2))
        out.append((q + "grn.beta", (1, 1, 2 * T), 0.02))
        lin(q + "pwconv2", T, 2 * T)
    lin(p + "input_embed.proj", D, 2 * cfg.mel_dim + T)
    for c in (0, 2):
        out.append((f"{p}input_embed.conv_pos_embed.conv1d.{c}.weight", (D, cpg, ks), 1.0 / math.sqrt(cpg * ks)))
        out.append((f"{p}input_embed.conv_pos_embed.conv1d.{c}.bias", (D,), 0.02))
    for i in range(cfg.depth):
        q = f"{p}transformer_blocks.{i}."
        lin(q + "attn_norm.linear", 6 * D, D, wstd=0.02)
        lin(q + "attn.to_q", inner, D)
        lin(q + "attn.to_k", inner, D)
        lin(q + "attn.to_v", inner, D)
        lin(q + "attn.to_out.0", D, inner)
        lin(q + "ff.ff.0.0", cfg.ff_mult * D, D)
        lin(q + "ff.ff.2", D, cfg.ff_mult * D)
    lin(p + "norm_out.linear", 2 * D, D, wstd=0.02)
    lin(p + "proj_out", cfg.mel_dim, D, wstd=0.02)
    return out


def make_dit_state_dict(cfg: DiTConfig, seed: int = 0, bf16_exact: bool = True) -> dict:
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for key, shape, std in dit_key_shapes(cfg):
        if std is None:
            w = 1.0 + _draw(g, shape, 0.1)
        else:
            w = _draw(g, shape, std)
        if bf16_exact:
            w = w.to(torch.bfloat16).float()
        sd[key] = w
    return sd


def vocos_key_shapes(vc: VocosConfig):
    out = []
    D, I = vc.dim, vc.intermediate_dim
    out.append(("backbone.embed.weight", (D, vc.n_mels, 7), 1.0 / math.sqrt(7 * vc.n_mels)))
    out.append(("backbone.embed.bias", (D,), 0.02))
    out.append(("backbone.norm.weight", (D,), None))
    out.append(("backbone.norm.bias", (D,), 0.02))
    for i in range(vc.num_layers):
        p = f"backbone.convnext.{i}."
        out.append((p + "dwconv.weight", (D, 1, 7), 1.0 / math.sqrt(7)))
        out.append((p + "dwconv.bias", (D,), 0.02))
        out.append((p + "norm.weight", (D,), None))
        out.append((p + "norm.bias", (D,), 0.02))
        out.append((p + "pwconv1.weight", (I, D), 1.0 / math.sqrt(D)))
        out.append((p + "pwconv1.bias", (I,), 0.02))
        out.append((p + "pwconv2.weight", (D, I), 1.0 / math.sqrt(I)))
        out.append((p + "pwconv2.bias", (D,), 0.02))
        out.append((p + "gamma", (D,), "gamma"))
    out.append(("backbone.final_layer_norm.weight", (D,), None))
    out.append(("backbone.final_layer_norm.bias", (D,), 0.02))
    # head.out drives exp(mag) (clipped at 1e2) and the phase; keep magnitudes log-mel-like
    out.append(("head.out.weight", (vc.n_fft + 2, D), 0.5 / math.sqrt(D)))
    out.append(("head.out.bias", (vc.n_fft + 2,), 0.1))
    return out


def make_vocos_state_dict(vc: VocosConfig, seed: int = 1, bf16_exact: bool = True) -> dict:
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for key, shape, std in vocos_key_shapes(vc):
        if std is None:
            w = 1.0 + _draw(g, shape, 0.1)
        elif std == "gamma":
            w = 0.125 + _draw(g, shape, 0.02)
        else:
            w = _draw(g, shape, std)
        if bf16_exact:
            w = w.to(torch.bfloat16).float()
        sd[key] = w
    return sd


def state_dict_digest(sd: dict) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def synthetic_inputs(cfg: DiTConfig, batch: int, ref_frames: int, total_frames, seed: int = 1234,
                     text_frac: float = 0.16):
    """SURVEY.md §8d synthetic workload: log-mel-like reference mel, random token ids padded with -1.
    `total_frames` may be an int or a per-item list (ragged batch)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    if isinstance(total_frames, int):
        total_frames = [total_frames] * batch
    cond = (torch.randn(batch, ref_frames, cfg.mel_dim, generator=g) * 2.0 - 1.5).clamp(-11.5, 5.0)
    nt = max(2, int(text_frac * max(total_frames)))
    text = torch.randint(0, cfg.text_num_embeds, (batch, nt), generator=g)
    for b in range(batch):      # ragged text lengths, padded with -1 like list_str_to_idx
        keep = max(1, int(nt * (0.6 + 0.4 * (b + 1) / batch)))
        text[b, keep:] = -1
    duration = torch.tensor(total_frames, dtype=torch.long)
    lens = torch.full((batch,), ref_frames, dtype=torch.long)
    return cond, text, duration, lens
