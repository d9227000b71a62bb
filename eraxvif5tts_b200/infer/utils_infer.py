"""Inference helpers with the reference's names (/root/reference/src/f5_tts/infer/utils_infer.py): defaults, chunk_text,
load_vocoder, load_checkpoint, load_model.  Host-side glue only."""
from __future__ import annotations

import os
import re

import torch

from ..model import CFM, DiT
from ..model.utils import get_tokenizer
from ..vocoder import Vocos

# utils_infer.py:49-62
target_sample_rate = 24000
n_mel_channels = 100
hop_length = 256
win_length = 1024
n_fft = 1024
mel_spec_type = "vocos"
target_rms = 0.1
cross_fade_duration = 0.15
ode_method = "euler"
nfe_step = 32
cfg_strength = 2.0
sway_sampling_coef = -1.0
speed = 1.0
fix_duration = None

# model.arch of the reference's hydra configs (src/f5_tts/configs/*.yaml) — constants only, no omegaconf needed
MODEL_ARCHS = {
    "F5TTS_Base": dict(dim=1024, depth=22, heads=16, ff_mult=2, text_dim=512, text_mask_padding=False, conv_layers=4, pe_attn_head=1),
    "F5TTS_Small": dict(dim=768, depth=18, heads=12, ff_mult=2, text_dim=512, text_mask_padding=False, conv_layers=4, pe_attn_head=1),
    "F5TTS_v1_Base": dict(dim=1024, depth=22, heads=16, ff_mult=2, text_dim=512, text_mask_padding=True, conv_layers=4, pe_attn_head=None),
    "F5TTS_v1_Pruned_14": dict(dim=1024, depth=14, heads=16, ff_mult=2, text_dim=512, text_mask_padding=True, conv_layers=4, pe_attn_head=None),
    "F5TTS_v1_Pruned_12": dict(dim=1024, depth=12, heads=16, ff_mult=2, text_dim=512, text_mask_padding=True, conv_layers=4, pe_attn_head=None),
}


def resolve_arch(model_name: str) -> dict:
    """`model_name` -> model.arch dict.  Names containing "custom" are paths to a YAML file with the reference's layout
    (f5tts_wrapper.py:128-135)."""
    if "custom" in model_name.lower() or model_name.endswith((".yaml", ".yml")):
        import yaml
        with open(model_name, "r", encoding="utf-8") as f:
            cfg = yaml.safe_load(f)
        arch = dict(cfg["model"]["arch"])
        if cfg["model"].get("backbone", "DiT") != "DiT":
            raise NotImplementedError("only the DiT backbone is on the north-star path (UNetT / MMDiT are out of scope)")
        arch.pop("qk_norm", None) if arch.get("qk_norm") is None else None
        arch.pop("checkpoint_activations", None)
        return arch
    if model_name not in MODEL_ARCHS:
        raise ValueError(f"unknown model {model_name!r}; known: {sorted(MODEL_ARCHS)} or a path to a custom YAML")
    return dict(MODEL_ARCHS[model_name])


def chunk_text(text, max_chars=135):
    """utils_infer.py:70-95: split on punctuation, pack sentences into chunks of <= max_chars UTF-8 bytes."""
    chunks = []
    current_chunk = ""
    sentences = re.split(r"(?<=[;:,.!?])\s+|(?<=[；：，。！？])", text)
    for sentence in sentences:
        piece = sentence + " " if sentence and len(sentence[-1].encode("utf-8")) == 1 else sentence
        if len(current_chunk.encode("utf-8")) + len(sentence.encode("utf-8")) <= max_chars:
            current_chunk += piece
        else:
            if current_chunk:
                chunks.append(current_chunk.strip())
            current_chunk = piece
    if current_chunk:
        chunks.append(current_chunk.strip())
    return chunks


def load_state_dict_file(path: str, device="cpu") -> dict:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device=str(device))
    return torch.load(path, map_location=device, weights_only=True)


def load_vocoder(vocoder_name="vocos", is_local=False, local_path="", device="cuda", hf_cache_dir=None, state_dict=None):
    """utils_infer.py:101-124.  There is no network here: weights come from `local_path/pytorch_model.bin` (the upstream
    vocos-mel-24khz layout) or an explicit `state_dict`; without either the vocoder keeps its random init (benchmarks)."""
    if vocoder_name != "vocos":
        raise NotImplementedError("only the vocos vocoder (the reference default) is on the north-star path")
    voc = Vocos()
    if state_dict is None and is_local and local_path:
        p = os.path.join(local_path, "pytorch_model.bin")
        if os.path.isfile(p):
            state_dict = torch.load(p, map_location="cpu", weights_only=True)
    if state_dict is not None:
        voc.load_state_dict(state_dict, strict=False)
    return voc.eval().to(device)


def load_checkpoint(model, ckpt_path, device, dtype=None, use_ema=True):
    """utils_infer.py:184-226 / f5tts_wrapper.py:201-254: .safetensors or .pt, EMA prefix strip, legacy mel keys dropped,
    pruned checkpoints (`model_state_dict` + `pruning_info`) accepted.  The masters stay fp32; bf16 packing happens in the engine."""
    ckpt_type = ckpt_path.split(".")[-1]
    checkpoint = load_state_dict_file(ckpt_path, "cpu")
    if use_ema and (ckpt_type == "safetensors" or "ema_model_state_dict" in checkpoint):
        if ckpt_type == "safetensors":
            checkpoint = {"ema_model_state_dict": checkpoint}
        sd = {k.replace("ema_model.", ""): v for k, v in checkpoint["ema_model_state_dict"].items() if k not in ["initted", "step"]}
    else:
        if ckpt_type == "safetensors":
            checkpoint = {"model_state_dict": checkpoint}
        sd = checkpoint["model_state_dict"]
    for key in ["mel_spec.mel_stft.mel_scale.fb", "mel_spec.mel_stft.spectrogram.window"]:
        sd.pop(key, None)
    sd = {re.sub(r"^(module\.|model\.|_orig_mod\.)+", "", k): v for k, v in sd.items()}
    model.load_state_dict(sd, strict=False)
    return model.to(device)


def load_model(model_cls, model_cfg, ckpt_path, mel_spec_type=mel_spec_type, vocab_file="", ode_method=ode_method, use_ema=True,
               device="cuda"):
    """utils_infer.py:232-286"""
    if model_cls is not DiT:
        raise NotImplementedError("only the DiT backbone is built")
    vocab_char_map, vocab_size = get_tokenizer(vocab_file, "custom")
    model = CFM(transformer=DiT(**model_cfg, text_num_embeds=vocab_size, mel_dim=n_mel_channels),
                mel_spec_kwargs=dict(n_fft=n_fft, hop_length=hop_length, win_length=win_length, n_mel_channels=n_mel_channels,
                                     target_sample_rate=target_sample_rate, mel_spec_type=mel_spec_type),
                odeint_kwargs=dict(method=ode_method), vocab_char_map=vocab_char_map).to(device)
    if ckpt_path:
        model = load_checkpoint(model, ckpt_path, device, use_ema=use_ema)
    return model
