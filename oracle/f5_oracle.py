"""CPU oracle (TEST INFRASTRUCTURE — see oracle/__init__.py): a functional torch-fp32 restatement of the
reference's F5-TTS hot path.  Every function cites the reference file:line it follows
(paths relative to /root/reference/src/f5_tts/).  Weights come in as a flat `state_dict` with the
reference's key names so the same dict drives the reference, this oracle and the
CUDA path.

Documented oracle adjustments:
  (1) SDPA dropout_p is 0.0 here; the reference hard-codes 0.1 at model/modules.py:490.
  (2) zero-initialised AdaLN / proj_out tensors are re-drawn by `oracle.weights` (a zero-init DiT
      outputs exactly 0, model/backbones/dit.py:162-172).
Third-party restatements ("parity unpinned", nothing in the reference tests them):
  torchdiffeq fixed-grid euler/midpoint, x_transformers rotary embedding, vocos.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F


@dataclass
class DiTConfig:
    """model.arch constants, configs/F5TTS_Base.yaml:25-34 and DiT.__init__ model/backbones/dit.py:104-122."""
    dim: int = 1024
    depth: int = 22
    heads: int = 16
    dim_head: int = 64
    ff_mult: int = 2
    mel_dim: int = 100
    text_num_embeds: int = 2545
    text_dim: int = 512
    text_mask_padding: bool = False
    conv_layers: int = 4
    pe_attn_head: int | None = 1
    conv_pos_kernel: int = 31
    conv_pos_groups: int = 16

    @staticmethod
    def base(**kw):
        return DiTConfig(**kw)

    @staticmethod
    def small(**kw):
        return DiTConfig(dim=768, depth=18, heads=12, **kw)

    @staticmethod
    def tiny(**kw):
        """CPU-fast config for parity tests (same structure, every code path exercised)."""
        d = dict(dim=128, depth=2, heads=2, text_num_embeds=40, text_dim=64, conv_layers=2)
        d.update(kw)
        return DiTConfig(**d)


@dataclass
class MelConfig:
    """model.mel_spec constants, configs/F5TTS_Base.yaml:35-41."""
    n_fft: int = 1024
    hop_length: int = 256
    win_length: int = 1024
    n_mel_channels: int = 100
    target_sample_rate: int = 24000


@dataclass
class VocosConfig:
    """charactr/vocos-mel-24khz hyper-parameters (third-party vocos package)."""
    n_mels: int = 100
    dim: int = 512
    intermediate_dim: int = 1536
    num_layers: int = 8
    n_fft: int = 1024
    hop_length: int = 256

    @staticmethod
    def tiny():
        return VocosConfig(dim=64, intermediate_dim=128, num_layers=2)


# --------------------------------------------------------------------------------------
# MelSpec  (model/modules.py:75-101 get_vocos_mel_spectrogram)
# --------------------------------------------------------------------------------------

def hz_to_mel_htk(f):
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank(n_freqs=513, f_min=0.0, f_max=12000.0, n_mels=100, sample_rate=24000) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') restated: triangular filters
    on an HTK mel grid; returns fb[n_freqs, n_mels] (what MelSpectrogram multiplies by)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min, m_max = hz_to_mel_htk(f_min), hz_to_mel_htk(f_max)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)


def melspec(wav: torch.Tensor, mc: MelConfig = MelConfig()) -> torch.Tensor:
    """wav [b, nw] -> log-mel [b, n_mels, 1 + nw // hop]   (model/modules.py:83-101,130-143).
    reflect-pad n_fft/2, periodic Hann, |rFFT| (power=1), HTK filterbank no norm, log(clamp 1e-5)."""
    if wav.ndim == 3:
        wav = wav.squeeze(1)
    assert wav.ndim == 2
    pad = mc.n_fft // 2
    x = F.pad(wav.unsqueeze(1).float(), (pad, pad), mode="reflect").squeeze(1)
    frames = x.unfold(-1, mc.n_fft, mc.hop_length)                       # [b, T, n_fft]
    win = torch.hann_window(mc.win_length, periodic=True, dtype=torch.float32, device=wav.device)
    spec = torch.fft.rfft(frames * win, dim=-1).abs()                     # [b, T, 513]
    fb = mel_filterbank(mc.n_fft // 2 + 1, 0.0, mc.target_sample_rate / 2, mc.n_mel_channels,
                        mc.target_sample_rate).to(wav.device)
    mel = spec @ fb                                                       # [b, T, n_mels]
    return mel.clamp(min=1e-5).log().transpose(1, 2)


# --------------------------------------------------------------------------------------
# small helpers  (model/utils.py:42-47, 88-95)
# --------------------------------------------------------------------------------------

def lens_to_mask(t: torch.Tensor, length: int | None = None) -> torch.Tensor:
    if length is None:
        length = int(t.amax())
    return torch.arange(length, device=t.device)[None, :] < t[:, None]


def list_str_to_idx(text, vocab_char_map, padding_value=-1) -> torch.Tensor:
    rows = [torch.tensor([vocab_char_map.get(c, 0) for c in t], dtype=torch.long) for t in text]
    return torch.nn.utils.rnn.pad_sequence(rows, padding_value=padding_value, batch_first=True)


# --------------------------------------------------------------------------------------
# DiT pieces
# --------------------------------------------------------------------------------------

def timestep_embedding(sd, time: torch.Tensor, p="transformer.time_embed.") -> torch.Tensor:
    """model/modules.py:149-161 (SinusPositionEmbedding, dim 256, scale 1000, denominator half_dim-1,
    order cat(sin, cos)) and :721-731 (Linear -> SiLU -> Linear)."""
    half = 128
    k = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, device=time.device).float() * -k)
    emb = 1000.0 * time.float().unsqueeze(1) * freqs.unsqueeze(0)
    emb = torch.cat((emb.sin(), emb.cos()), dim=-1).to(time.dtype)
    h = F.linear(emb, sd[p + "time_mlp.0.weight"], sd[p + "time_mlp.0.bias"])
    return F.linear(F.silu(h), sd[p + "time_mlp.2.weight"], sd[p + "time_mlp.2.bias"])


def text_pos_table(dim: int, end: int = 4096, theta: float = 10000.0) -> torch.Tensor:
    """precompute_freqs_cis, model/modules.py:196-207: cat(cos, sin) of outer(pos, theta^(-2j/dim))."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    ang = torch.outer(torch.arange(end).float(), freqs)
    return torch.cat([ang.cos(), ang.sin()], dim=-1)


def grn(x, gamma, beta):
    """model/modules.py:225-234 — L2 norm over dim=1 (the SEQUENCE axis), mean over channels."""
    gx = torch.norm(x, p=2, dim=1, keepdim=True)
    nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
    return gamma * (x * nx) + beta + x


def convnext_v2_block(sd, p, x):
    """model/modules.py:241-269: dwconv k7 -> LN(affine, 1e-6) -> Linear -> GELU(erf) -> GRN -> Linear -> +res."""
    dim = x.shape[-1]
    h = F.conv1d(x.transpose(1, 2), sd[p + "dwconv.weight"], sd[p + "dwconv.bias"], padding=3, groups=dim)
    h = h.transpose(1, 2)
    h = F.layer_norm(h, (dim,), sd[p + "norm.weight"], sd[p + "norm.bias"], eps=1e-6)
    h = F.gelu(F.linear(h, sd[p + "pwconv1.weight"], sd[p + "pwconv1.bias"]))
    h = grn(h, sd[p + "grn.gamma"], sd[p + "grn.beta"])
    h = F.linear(h, sd[p + "pwconv2.weight"], sd[p + "pwconv2.bias"])
    return x + h


def text_embedding(sd, cfg: DiTConfig, text: torch.Tensor, seq_len: int, drop_text: bool) -> torch.Tensor:
    """TextEmbedding.forward, model/backbones/dit.py:49-79."""
    p = "transformer.text_embed."
    text = (text + 1)[:, :seq_len]
    text = F.pad(text, (0, seq_len - text.shape[1]), value=0)
    text_mask = text == 0
    if drop_text:
        text = torch.zeros_like(text)
    h = F.embedding(text, sd[p + "text_embed.weight"])
    if cfg.conv_layers > 0:
        pos = torch.arange(seq_len).clamp(max=4095)
        h = h + text_pos_table(cfg.text_dim)[pos].to(h)
        if cfg.text_mask_padding:
            m = text_mask.unsqueeze(-1)
            h = h.masked_fill(m, 0.0)
            for j in range(cfg.conv_layers):
                h = convnext_v2_block(sd, f"{p}text_blocks.{j}.", h).masked_fill(m, 0.0)
        else:
            for j in range(cfg.conv_layers):
                h = convnext_v2_block(sd, f"{p}text_blocks.{j}.", h)
    return h


def conv_pos_embed(sd, cfg: DiTConfig, x):
    """ConvPositionEmbedding.forward with mask=None, model/modules.py:167-190: 2x [grouped conv -> Mish]."""
    p = "transformer.input_embed.conv_pos_embed.conv1d."
    pad = cfg.conv_pos_kernel // 2
    h = x.transpose(1, 2)
    h = F.mish(F.conv1d(h, sd[p + "0.weight"], sd[p + "0.bias"], padding=pad, groups=cfg.conv_pos_groups))
    h = F.mish(F.conv1d(h, sd[p + "2.weight"], sd[p + "2.bias"], padding=pad, groups=cfg.conv_pos_groups))
    return h.transpose(1, 2)


def input_embedding(sd, cfg, x, cond, text_embed, drop_audio_cond: bool):
    """InputEmbedding.forward, model/backbones/dit.py:91-97 (conv_pos_embed gets NO mask)."""
    if drop_audio_cond:
        cond = torch.zeros_like(cond)
    h = F.linear(torch.cat((x, cond, text_embed), dim=-1),
                 sd["transformer.input_embed.proj.weight"], sd["transformer.input_embed.proj.bias"])
    return conv_pos_embed(sd, cfg, h) + h


def rotary_freqs(seq_len: int, dim_head: int = 64, theta: float = 10000.0) -> torch.Tensor:
    """x_transformers.RotaryEmbedding.forward_from_seq_len (third party, parity unpinned):
    freqs[n, dim_head] with every frequency repeated in adjacent pairs."""
    inv = 1.0 / (theta ** (torch.arange(0, dim_head, 2).float() / dim_head))
    f = torch.outer(torch.arange(seq_len).float(), inv)
    return torch.stack((f, f), dim=-1).flatten(-2)


def apply_rotary(t: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """x_transformers.apply_rotary_pos_emb (third party): t*cos + rotate_half(t)*sin with the
    interleaved-pair rotate_half (x0,x1)->(-x1,x0), fp32 math, cast back."""
    tf = t.float()
    x = tf.reshape(*tf.shape[:-1], -1, 2)
    rot = torch.stack((-x[..., 1], x[..., 0]), dim=-1).flatten(-2)
    return (tf * freqs.cos() + rot * freqs.sin()).to(t.dtype)


def dropout_multipliers(p: float, seed: int, layer: int, site: int, shape) -> torch.Tensor:
    """The product's train-mode dropout masks (include/f5b200.h: f5b_train_set_dropout), restated so that autograd through this
    oracle sees the SAME masks as the CUDA path: torch's nn.Dropout (modules.py:349, :439) draws from Philox, whose stream is an
    implementation detail no other implementation can reproduce, so the mask generator is the one thing on this path that is
    defined by the product (a splitmix64 hash of (seed, layer, site, element // 4), 16 bits per element) and only its
    distribution -- Bernoulli(1 - p) keeps scaled by 1 / (1 - p) -- by the reference.  site 0 = FeedForward's Dropout,
    site 1 = the Dropout behind to_out.  Returns a float32 tensor of `shape` holding 0 or 1 / (1 - p)."""
    import numpy as np
    numel = 1
    for d in shape:
        numel *= int(d)
    assert numel % 4 == 0
    thr = min(65535, int(np.float32(p) * np.float32(65536.0) + np.float32(0.5)))
    M = (1 << 64) - 1
    key = np.uint64((int(seed) * 0xD1342543DE82EF95 + (layer * 8 + site + 1) * 0x9E3779B97F4A7C15) & M)
    with np.errstate(over="ignore"):
        z = np.arange(numel // 4, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + key
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    lanes = np.stack([(z >> np.uint64(16 * k)) & np.uint64(0xFFFF) for k in range(4)], axis=1).reshape(-1)
    scale = np.float32(65536.0) / (np.float32(65536.0) - np.float32(thr))
    return torch.from_numpy(np.where(lanes >= thr, scale, np.float32(0.0)).astype(np.float32)).reshape(tuple(shape))


def attention_dropout_multipliers(p: float, seed: int, layer: int, BH: int, n: int) -> torch.Tensor:
    """The product's mask of the dropout inside scaled_dot_product_attention (modules.py:490; include/f5b200.h:
    f5b_train_set_attn_dropout, eraxvif5tts_b200/csrc/dropout.cuh: AttnDrop), restated for the same reason as dropout_multipliers.
    One Philox-2x32 block of 7 rounds per 8 consecutive keys of a query row: counter = group index ((bh * n + query) * ceil(n/8)
    + key // 8), key / counter-high xor from the (seed, layer) key; element key % 8 owns one byte, kept iff (byte & 0x7f) >= t7 with
    t7 = round(p * 128); kept values are scaled by 128 / (128 - t7).  Returns float32 [BH, n, n]."""
    import numpy as np
    n8 = (n + 7) // 8
    t7 = min(127, max(1, int(np.float32(p) * np.float32(128.0) + np.float32(0.5))))
    M = (1 << 64) - 1
    key = (int(seed) * 0xD1342543DE82EF95 + (layer * 8 + 3) * 0x9E3779B97F4A7C15) & M
    g = np.arange(BH * n * n8, dtype=np.uint64)
    c0 = (g & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    c1 = (g >> np.uint64(32)) ^ np.uint64(key >> 32)
    k = key & 0xFFFFFFFF
    for _ in range(7):
        pr = c0 * np.uint64(0xD256D193)  # 32 x 32 -> 64 bits: no overflow
        c0 = (pr >> np.uint64(32)) ^ np.uint64(k) ^ c1
        c1 = pr & np.uint64(0xFFFFFFFF)
        k = (k + 0x9E3779B9) & 0xFFFFFFFF
    lanes = np.stack([((c0 if j < 4 else c1) >> np.uint64(8 * (j & 3))) & np.uint64(0x7F) for j in range(8)], axis=1)  # [groups, 8]
    scale = np.float32(128.0) / np.float32(128 - t7)
    m = np.where(lanes >= t7, scale, np.float32(0.0)).astype(np.float32).reshape(BH, n, n8 * 8)[:, :, :n]
    return torch.from_numpy(np.ascontiguousarray(m))


def attention(sd, cfg: DiTConfig, p, x, mask, rope, drop=None, attn_drop=None):
    """AttnProcessor.__call__, model/modules.py:442-503 (dropout_p 0.0 unless attn_drop is given; reference :490 says 0.1).
    drop: optional callable applied to to_out's result (the Dropout of to_out, :439-440) before the padding mask.
    attn_drop: optional callable (b, H, n) -> multipliers [b, H, n, n] applied to the normalised probabilities (SDPA's dropout)."""
    b, n, _ = x.shape
    H, d = cfg.heads, cfg.dim_head
    q = F.linear(x, sd[p + "to_q.weight"], sd[p + "to_q.bias"]).view(b, n, H, d).transpose(1, 2)
    k = F.linear(x, sd[p + "to_k.weight"], sd[p + "to_k.bias"]).view(b, n, H, d).transpose(1, 2)
    v = F.linear(x, sd[p + "to_v.weight"], sd[p + "to_v.bias"]).view(b, n, H, d).transpose(1, 2)
    pn = cfg.pe_attn_head if cfg.pe_attn_head is not None else H
    q = torch.cat((apply_rotary(q[:, :pn], rope), q[:, pn:]), dim=1)
    k = torch.cat((apply_rotary(k[:, :pn], rope), k[:, pn:]), dim=1)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    if mask is not None:
        s = s.masked_fill(~mask[:, None, None, :], float("-inf"))
    pr = torch.softmax(s.float(), dim=-1)
    if attn_drop is not None:
        pr = pr * attn_drop(b, H, n).to(pr.device)
    o = pr.to(v.dtype) @ v
    o = o.transpose(1, 2).reshape(b, n, H * d)
    o = F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])
    if drop is not None:
        o = drop(o)
    if mask is not None:
        o = o.masked_fill(~mask.unsqueeze(-1), 0.0)
    return o


def dit_block(sd, cfg: DiTConfig, i: int, x, t, mask, rope, dropout=None):
    """DiTBlock.forward, model/modules.py:627-641 with AdaLayerNorm :310-315 (chunk order
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp) and FeedForward(GELU tanh) :342-353.
    dropout: None (eval) or (p, seed[, attn_p]) -- train mode with the masks of dropout_multipliers; attn_p (default 0) is the
    probability of SDPA's own dropout (:490), whose mask is attention_dropout_multipliers."""
    def drop(site):
        if dropout is None or not dropout[0] > 0:
            return None
        return lambda a: a * dropout_multipliers(dropout[0], dropout[1], i, site, a.shape).to(a.device)

    attn_drop = None
    if dropout is not None and len(dropout) > 2 and dropout[2] > 0:
        def attn_drop(b, H, n):
            return attention_dropout_multipliers(dropout[2], dropout[1], i, b * H, n).reshape(b, H, n, n)
    p = f"transformer.transformer_blocks.{i}."
    D = cfg.dim
    emb = F.linear(F.silu(t), sd[p + "attn_norm.linear.weight"], sd[p + "attn_norm.linear.bias"])
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = torch.chunk(emb, 6, dim=1)
    h = F.layer_norm(x, (D,), eps=1e-6) * (1 + scale_msa[:, None]) + shift_msa[:, None]
    akw = {} if attn_drop is None else {"attn_drop": attn_drop}  # (bench.py's eager comparator swaps `attention` for an SDPA version)
    x = x + gate_msa.unsqueeze(1) * attention(sd, cfg, p + "attn.", h, mask, rope, drop(1), **akw)
    h = F.layer_norm(x, (D,), eps=1e-6) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
    h = F.gelu(F.linear(h, sd[p + "ff.ff.0.0.weight"], sd[p + "ff.ff.0.0.bias"]), approximate="tanh")
    if drop(0) is not None:
        h = drop(0)(h)
    h = F.linear(h, sd[p + "ff.ff.2.weight"], sd[p + "ff.ff.2.bias"])
    return x + gate_mlp.unsqueeze(1) * h


def dit_forward(sd, cfg: DiTConfig, x, cond, text, time, drop_audio_cond, drop_text, mask=None,
                text_embed=None, return_hidden=False, dropout=None):
    """DiT.forward, model/backbones/dit.py:185-233 (final AdaLN chunk order is (scale, shift),
    model/modules.py:331-336)."""
    b, n = x.shape[:2]
    if time.ndim == 0:
        time = time.repeat(b)
    t = timestep_embedding(sd, time)
    if text_embed is None:
        text_embed = text_embedding(sd, cfg, text, n, drop_text)
    h = input_embedding(sd, cfg, x, cond, text_embed, drop_audio_cond)
    rope = rotary_freqs(n, cfg.dim_head).to(x.device)
    hidden = [h]
    for i in range(cfg.depth):
        h = dit_block(sd, cfg, i, h, t, mask, rope, dropout)
        if return_hidden:
            hidden.append(h)
    emb = F.linear(F.silu(t), sd["transformer.norm_out.linear.weight"], sd["transformer.norm_out.linear.bias"])
    scale, shift = torch.chunk(emb, 2, dim=1)
    h = F.layer_norm(h, (cfg.dim,), eps=1e-6) * (1 + scale)[:, None, :] + shift[:, None, :]
    out = F.linear(h, sd["transformer.proj_out.weight"], sd["transformer.proj_out.bias"])
    return (out, hidden) if return_hidden else out


# --------------------------------------------------------------------------------------
# CFM.sample  (model/cfm.py:82-208)  +  torchdiffeq fixed-grid solvers (third party, unpinned)
# --------------------------------------------------------------------------------------

def odeint_fixed(fn, y0, t, method="euler"):
    """torchdiffeq.odeint on the user grid `t`: euler  y += dt*f(t_i, y);
    midpoint  y += dt*f(t_i + dt/2, y + dt/2*f(t_i, y)).  Returns the stacked states."""
    ys = [y0]
    y = y0
    for i in range(len(t) - 1):
        t0, t1 = t[i], t[i + 1]
        dt = t1 - t0
        if method == "euler":
            y = y + dt * fn(t0, y)
        elif method == "midpoint":
            half = 0.5 * dt
            y = y + dt * fn(t0 + half, y + half * fn(t0, y))
        else:
            raise ValueError(method)
        ys.append(y)
    return torch.stack(ys)


def sway_time_grid(steps, sway_sampling_coef, dtype=torch.float32, t_start=0.0, device="cpu"):
    """model/cfm.py:193-195, evaluated in the model dtype exactly like the reference."""
    t = torch.linspace(t_start, 1, steps + 1, device=device, dtype=dtype)
    if sway_sampling_coef is not None:
        t = t + sway_sampling_coef * (torch.cos(torch.pi / 2 * t) - 1 + t)
    return t


def cfm_sample(sd, cfg: DiTConfig, cond, text, duration, *, lens=None, steps=32, cfg_strength=1.0,
               sway_sampling_coef=None, seed=None, max_duration=4096, method="euler",
               no_ref_audio=False, edit_mask=None, mel_cfg: MelConfig = MelConfig(), vocab_char_map=None,
               duplicate_test=False, t_inter=0.1):
    """CFM.sample, model/cfm.py:82-208, incl. the duplicate_test / t_inter corner (:139-140, :188-191).  Runs on the device of
    `cond` (fp32; the noise is drawn on the CPU generator and moved, so CPU and GPU runs of the oracle share y0).
    Returns (out, trajectory)."""
    if cond.ndim == 2:
        cond = melspec(cond, mel_cfg).permute(0, 2, 1)
    cond = cond.float()
    dev = cond.device
    batch, cond_seq_len = cond.shape[:2]
    if lens is None:
        lens = torch.full((batch,), cond_seq_len, dtype=torch.long, device=dev)
    if isinstance(text, list):
        text = list_str_to_idx(text, vocab_char_map).to(dev)
    cond_mask = lens_to_mask(lens)
    if edit_mask is not None:
        cond_mask = cond_mask & edit_mask
    if isinstance(duration, int):
        duration = torch.full((batch,), duration, dtype=torch.long, device=dev)
    duration = torch.maximum(torch.maximum((text != -1).sum(dim=-1), lens) + 1, duration).clamp(max=max_duration)
    max_dur = int(duration.amax())
    if duplicate_test:
        test_cond = F.pad(cond, (0, 0, cond_seq_len, max_dur - 2 * cond_seq_len), value=0.0)
    cond = F.pad(cond, (0, 0, 0, max_dur - cond_seq_len), value=0.0)
    if no_ref_audio:
        cond = torch.zeros_like(cond)
    cond_mask = F.pad(cond_mask, (0, max_dur - cond_mask.shape[-1]), value=False).unsqueeze(-1)
    step_cond = torch.where(cond_mask, cond, torch.zeros_like(cond))
    mask = lens_to_mask(duration) if batch > 1 else None

    te_cond = text_embedding(sd, cfg, text, max_dur, drop_text=False)      # text cache, dit.py:202-210
    te_unc = text_embedding(sd, cfg, text, max_dur, drop_text=True)

    def fn(t, x):
        pred = dit_forward(sd, cfg, x, step_cond, text, t, False, False, mask, text_embed=te_cond)
        if cfg_strength < 1e-5:
            return pred
        null = dit_forward(sd, cfg, x, step_cond, text, t, True, True, mask, text_embed=te_unc)
        return pred + (pred - null) * cfg_strength

    y0 = []
    for dur in duration.tolist():
        if seed is not None:
            torch.manual_seed(seed)
        y0.append(torch.randn(int(dur), cfg.mel_dim))
    y0 = torch.nn.utils.rnn.pad_sequence(y0, padding_value=0, batch_first=True).to(dev)
    t_start = 0.0
    if duplicate_test:
        t_start = t_inter
        y0 = (1 - t_start) * y0 + t_start * test_cond
        steps = int(steps * (1 - t_start))
    t = sway_time_grid(steps, sway_sampling_coef, t_start=t_start, device=dev)
    traj = odeint_fixed(fn, y0, t, method)
    out = torch.where(cond_mask, cond, traj[-1])
    return out, traj


# --------------------------------------------------------------------------------------
# CFM.forward loss (model/cfm.py:210-283), deterministic form: the random draws are arguments
# --------------------------------------------------------------------------------------

def cfm_loss(sd, cfg: DiTConfig, x1, text, rand_span_mask, x0, time, drop_audio_cond, drop_text, dropout=None):
    """CFM.forward, model/cfm.py:210-283, with the call's random draws as arguments.  Pinned (loss, pred and autograd gradients)
    against the reference's own CFM.forward + loss.backward() by tests/test_oracle_golden.py on tests/golden/cfm_forward_*.pt."""
    t = time[:, None, None]
    phi = (1 - t) * x0 + t * x1
    flow = x1 - x0
    cond = torch.where(rand_span_mask[..., None], torch.zeros_like(x1), x1)
    pred = dit_forward(sd, cfg, phi, cond, text, time, drop_audio_cond, drop_text, mask=None, dropout=dropout)
    loss = F.mse_loss(pred, flow, reduction="none")[rand_span_mask]
    return loss.mean(), cond, pred


def distill_losses(student_sd, student_cfg: DiTConfig, teacher_sd, teacher_cfg: DiTConfig, x1, text, rand_span_mask, x0, time,
                   drop_audio_cond, drop_text, alpha=0.5, loss_type="mse", spec_l1_weight=0.0, dropout=None):
    """One distillation step's losses, train/distil_reload.py:1044-1093 (the step is inline in that script's loop, not a function,
    so this restatement is checked by reading, not by calling the reference): teacher forward without gradient and always
    conditioned (:1054-1057), student forward with the CFG drops (:1060-1065), and -- unlike CFM.forward -- losses summed over
    channels and divided by the number of masked FRAMES (:1069-1071, :1092-1093).  Returns (total, student, distill, spec_l1, pred)."""
    t = time[:, None, None]
    xt = (1 - t) * x0 + t * x1
    flow = x1 - x0
    cond = torch.where(rand_span_mask[..., None], torch.zeros_like(x1), x1)
    with torch.no_grad():
        teacher = dit_forward(teacher_sd, teacher_cfg, xt, cond, text, time, False, False, mask=None)
    pred = dit_forward(student_sd, student_cfg, xt, cond, text, time, drop_audio_cond, drop_text, mask=None, dropout=dropout)
    m = rand_span_mask.unsqueeze(-1)
    cnt = rand_span_mask.sum().clamp(min=1)
    student = (F.mse_loss(pred, flow, reduction="none") * m).sum() / cnt
    if loss_type == "mse":
        full = F.mse_loss(pred, teacher, reduction="none")
    elif loss_type == "l1":
        full = F.l1_loss(pred, teacher, reduction="none")
    else:
        raise ValueError(f"Unsupported distill_loss_type: {loss_type}")
    distill = (full * m).sum() / cnt
    spec = (F.l1_loss(pred, teacher, reduction="none") * m).sum() / cnt if spec_l1_weight > 0 else torch.zeros(())
    total = (1.0 - alpha) * student + alpha * distill + spec * spec_l1_weight
    return total, student, distill, spec, pred


# --------------------------------------------------------------------------------------
# Vocos (third-party package `vocos`, config charactr/vocos-mel-24khz; unpinned)
# --------------------------------------------------------------------------------------

def vocos_backbone(sd, vc: VocosConfig, mel):
    h = F.conv1d(mel, sd["backbone.embed.weight"], sd["backbone.embed.bias"], padding=3)
    h = F.layer_norm(h.transpose(1, 2), (vc.dim,), sd["backbone.norm.weight"], sd["backbone.norm.bias"], 1e-6)
    for i in range(vc.num_layers):
        p = f"backbone.convnext.{i}."
        r = h
        y = F.conv1d(h.transpose(1, 2), sd[p + "dwconv.weight"], sd[p + "dwconv.bias"], padding=3, groups=vc.dim)
        y = F.layer_norm(y.transpose(1, 2), (vc.dim,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
        y = F.gelu(F.linear(y, sd[p + "pwconv1.weight"], sd[p + "pwconv1.bias"]))
        y = F.linear(y, sd[p + "pwconv2.weight"], sd[p + "pwconv2.bias"]) * sd[p + "gamma"]
        h = r + y
    return F.layer_norm(h, (vc.dim,), sd["backbone.final_layer_norm.weight"],
                        sd["backbone.final_layer_norm.bias"], 1e-6)


def istft_center(spec: torch.Tensor, n_fft=1024, hop=256) -> torch.Tensor:
    """torch.istft(center=True, window=hann) restated: irFFT per frame, window, overlap-add, divide by the
    overlap-added squared window, drop n_fft/2 from each side.  spec complex [b, n_fft/2+1, T] -> [b, hop*(T-1)]."""
    b, _, T = spec.shape
    win = torch.hann_window(n_fft, periodic=True, dtype=torch.float32, device=spec.device)
    frames = torch.fft.irfft(spec.transpose(1, 2), n=n_fft, dim=-1) * win       # [b, T, n_fft]
    out_len = n_fft + hop * (T - 1)
    y = F.fold(frames.transpose(1, 2), (1, out_len), (1, n_fft), stride=(1, hop)).reshape(b, out_len)
    env = F.fold((win * win).expand(1, T, n_fft).transpose(1, 2), (1, out_len), (1, n_fft),
                 stride=(1, hop)).reshape(out_len)
    pad = n_fft // 2
    return y[:, pad:out_len - pad] / env[pad:out_len - pad]


def vocos_decode(sd, vc: VocosConfig, mel: torch.Tensor) -> torch.Tensor:
    """Vocos.decode(mel[b,100,T]) -> wav[b, 256*(T-1)]: backbone -> Linear(dim, n_fft+2) -> exp/clip(1e2) magnitude,
    cos/sin phase -> iSTFT."""
    h = vocos_backbone(sd, vc, mel.float())
    o = F.linear(h, sd["head.out.weight"], sd["head.out.bias"]).transpose(1, 2)
    mag, ph = o.chunk(2, dim=1)
    mag = torch.clip(torch.exp(mag), max=1e2)
    spec = mag * (torch.cos(ph) + 1j * torch.sin(ph))
    return istft_center(spec, vc.n_fft, vc.hop_length)


# --------------------------------------------------------------------------------------
# F5TTSWrapper.generate numeric tail (infer/f5tts_wrapper.py:517-531)
# --------------------------------------------------------------------------------------

def generate_tail(vsd, vc: VocosConfig, out_mel, ref_audio_len: int, ref_wav, target_rms=0.1):
    gen = out_mel.float()[:, ref_audio_len:, :].permute(0, 2, 1)
    wave = vocos_decode(vsd, vc, gen)
    rms = torch.sqrt(torch.mean(torch.square(ref_wav)))
    if rms < target_rms:
        wave = wave * rms / target_rms
    return wave, gen
