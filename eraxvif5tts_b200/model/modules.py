"""MelSpec with the reference's interface (/root/reference/src/f5_tts/model/modules.py:104-143), computed by the CUDA kernel
`f5b_melspec` (framing + in-smem FFT + mel + log in one pass)."""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import ops


def hz_to_mel_htk(f: float) -> float:
    return 2595.0 * math.log10(1.0 + f / 700.0)


def melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk") — the filterbank torchaudio's MelSpectrogram builds for
    get_vocos_mel_spectrogram (modules.py:83-93); returns fb[n_freqs, n_mels]."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(hz_to_mel_htk(f_min), hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)


class MelSpec(nn.Module):
    """modules.py:104-143.  forward(wav [b, nw]) -> log-mel [b, n_mel_channels, 1 + nw // hop]."""

    def __init__(self, n_fft=1024, hop_length=256, win_length=1024, n_mel_channels=100, target_sample_rate=24_000,
                 mel_spec_type="vocos"):
        super().__init__()
        if mel_spec_type != "vocos":
            raise NotImplementedError("only the vocos mel (the reference's configured mel_spec_type) is built; bigvgan is out of scope")
        if (n_fft, hop_length, win_length) != (1024, 256, 1024):
            raise NotImplementedError("the CUDA mel kernel is built for n_fft=1024, hop=256, win=1024 (configs/F5TTS_Base.yaml:35-41)")
        self.n_fft, self.hop_length, self.win_length = n_fft, hop_length, win_length
        self.n_mel_channels, self.target_sample_rate = n_mel_channels, target_sample_rate
        self.mel_spec_type = mel_spec_type
        self.register_buffer("dummy", torch.tensor(0), persistent=False)
        fb = melscale_fbanks_htk(n_fft // 2 + 1, 0.0, target_sample_rate / 2, n_mel_channels, target_sample_rate)
        nz = fb > 0
        f0 = torch.where(nz.any(0), nz.float().argmax(0), torch.zeros(n_mel_channels, dtype=torch.long))
        f1 = torch.where(nz.any(0), fb.shape[0] - nz.flip(0).float().argmax(0), torch.zeros(n_mel_channels, dtype=torch.long))
        self.register_buffer("fb", fb.contiguous(), persistent=False)
        self.register_buffer("fb_ranges", torch.stack((f0, f1), -1).to(torch.int32).contiguous(), persistent=False)

    def forward_token_major(self, wav: torch.Tensor) -> torch.Tensor:
        """wav [b, nw] -> [b, T, n_mels] (the layout CFM consumes; saves the permute of cfm.py:105)."""
        if wav.ndim == 3:
            wav = wav.squeeze(1)
        assert wav.ndim == 2
        if self.fb.device != wav.device:  # the reference moves itself lazily too (modules.py:131-132)
            self.to(wav.device)
        with torch.cuda.device(wav.device):  # the library launches on the current device: make it the input's
            return ops.melspec(wav.float().contiguous(), self.fb, self.fb_ranges, self.n_mel_channels)

    def forward(self, wav: torch.Tensor) -> torch.Tensor:
        return self.forward_token_major(wav).permute(0, 2, 1)


_MELSPEC_CACHE: dict = {}


def get_vocos_mel_spectrogram(waveform, n_fft=1024, n_mel_channels=100, target_sample_rate=24000, hop_length=256, win_length=1024):
    """The functional form the reference's MelSpec delegates to (model/modules.py:75-101): waveform [b, nw] (or [b, 1, nw]) on the
    GPU -> log-mel [b, n_mels, frames], through the same `f5b_melspec` kernel as `MelSpec`."""
    key = (n_fft, hop_length, win_length, n_mel_channels, target_sample_rate)
    if key not in _MELSPEC_CACHE:
        _MELSPEC_CACHE[key] = MelSpec(n_fft=n_fft, hop_length=hop_length, win_length=win_length, n_mel_channels=n_mel_channels,
                                      target_sample_rate=target_sample_rate, mel_spec_type="vocos")
    if waveform.ndim == 3:
        waveform = waveform.squeeze(1)
    return _MELSPEC_CACHE[key](waveform)


def precompute_freqs_cis(dim: int, end: int, theta: float = 10000.0, theta_rescale_factor=1.0):
    """model/modules.py:196-207 (the TextEmbedding position table); the DiT engine's own copy lives in backbones/dit.py"""
    theta = theta * theta_rescale_factor ** (dim / (dim - 2))
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
    freqs = torch.outer(torch.arange(end), freqs).float()
    return torch.cat([torch.cos(freqs), torch.sin(freqs)], dim=-1)


def get_pos_embed_indices(start, length, max_pos, scale=1.0):
    """model/modules.py:210-219: start [b] -> position indices [b, length], start + floor(arange * scale), clamped below max_pos"""
    scale = scale * torch.ones_like(start, dtype=torch.float32)
    pos = start.unsqueeze(1) + (torch.arange(length, device=start.device, dtype=torch.float32).unsqueeze(0) * scale.unsqueeze(1)).long()
    return torch.where(pos < max_pos, pos, max_pos - 1)
