"""Inference helpers with the reference's names (/root/reference/src/f5_tts/infer/utils_infer.py): defaults, chunk_text,
load_vocoder, load_checkpoint, load_model, preprocess_ref_audio_text, remove_silence_edges, remove_silence_for_generated_wav,
infer_process / infer_batch_process.  Host-side glue only."""
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


# ------------------------------------------------------------------------------------------------------------------------------
# infer_process / infer_batch_process (utils_infer.py:366-563): the call sites the reference's api.py, CLI, Gradio app and socket
# server go through.  Same arguments and yields; `batch_chunks` (extension) samples all chunks as ONE ragged batch
# instead of the reference's thread pool of B=1 sample() calls, and the cross-fade fold runs on the device.

def infer_process(ref_audio, ref_text, gen_text, model_obj, vocoder, mel_spec_type=mel_spec_type, show_info=print, progress=None,
                  target_rms=target_rms, cross_fade_duration=cross_fade_duration, nfe_step=nfe_step, cfg_strength=cfg_strength,
                  sway_sampling_coef=sway_sampling_coef, speed=speed, fix_duration=fix_duration, device="cuda", batch_chunks=True):
    """utils_infer.py:366-410.  ref_audio: a wav path, or (audio [channels, samples], sample_rate)."""
    if isinstance(ref_audio, (str, os.PathLike)):
        from .f5tts_wrapper import _read_wav
        audio, sr = _read_wav(str(ref_audio))
    else:
        audio, sr = ref_audio
        audio = torch.as_tensor(audio, dtype=torch.float32)
        if audio.ndim == 1:
            audio = audio.unsqueeze(0)
    max_chars = int(len(ref_text.encode("utf-8")) / (audio.shape[-1] / sr) * (22 - audio.shape[-1] / sr))
    gen_text_batches = chunk_text(gen_text, max_chars=max_chars)
    show_info(f"Generating audio in {len(gen_text_batches)} batches...")
    return next(infer_batch_process((audio, sr), ref_text, gen_text_batches, model_obj, vocoder, mel_spec_type=mel_spec_type,
                                    progress=progress, target_rms=target_rms, cross_fade_duration=cross_fade_duration, nfe_step=nfe_step,
                                    cfg_strength=cfg_strength, sway_sampling_coef=sway_sampling_coef, speed=speed,
                                    fix_duration=fix_duration, device=device, batch_chunks=batch_chunks))


def infer_batch_process(ref_audio, ref_text, gen_text_batches, model_obj, vocoder, mel_spec_type="vocos", progress=None, target_rms=0.1,
                        cross_fade_duration=0.15, nfe_step=32, cfg_strength=2.0, sway_sampling_coef=-1, speed=1, fix_duration=None,
                        device="cuda", streaming=False, chunk_size=2048, batch_chunks=True, seed=None):
    """utils_infer.py:417-563 (a generator, like the reference).  streaming=False yields (final_wave, sample_rate,
    combined_spectrogram) once; streaming=True yields (chunk of <= chunk_size samples, sample_rate) per piece, text chunk by text
    chunk.  `progress` is accepted for signature compatibility (tqdm wrappers are control plane)."""
    import numpy as np
    from .. import ops
    from .f5tts_wrapper import convert_char_to_pinyin
    if mel_spec_type != "vocos":
        raise NotImplementedError("only the vocos mel / vocoder pair is built (BigVGAN is an absent submodule of the reference)")
    audio, sr = ref_audio
    audio = torch.as_tensor(audio, dtype=torch.float32)
    if audio.ndim == 1:
        audio = audio.unsqueeze(0)
    if audio.shape[0] > 1:
        audio = torch.mean(audio, dim=0, keepdim=True)
    rms = torch.sqrt(torch.mean(torch.square(audio)))
    if rms < target_rms:
        audio = audio * target_rms / rms
    if sr != target_sample_rate:
        import torchaudio
        audio = torchaudio.transforms.Resample(sr, target_sample_rate)(audio)
    audio = audio.to(device)
    if len(ref_text[-1].encode("utf-8")) == 1:
        ref_text = ref_text + " "
    ref_audio_len = audio.shape[-1] // hop_length

    def chunk_duration(gen_text):
        local_speed = 0.3 if len(gen_text.encode("utf-8")) < 10 else speed
        if fix_duration is not None:
            return int(fix_duration * target_sample_rate / hop_length)
        ref_text_len, gen_text_len = len(ref_text.encode("utf-8")), len(gen_text.encode("utf-8"))
        return ref_audio_len + int(ref_audio_len / ref_text_len * gen_text_len / local_speed)

    def finish(gen_mel):  # [1, n, mel] with the reference part attached -> (wave on the device, mel [mel, n_gen] on the host)
        g = gen_mel.to(torch.float32)[:, ref_audio_len:, :].permute(0, 2, 1)
        wave = vocoder.decode(g)
        if rms < target_rms:
            wave = wave * rms / target_rms
        return wave.reshape(-1), g[0].cpu().numpy()

    def sample(texts, durations):
        with torch.inference_mode():
            out, _ = model_obj.sample(cond=audio.expand(len(texts), -1) if len(texts) > 1 else audio, text=convert_char_to_pinyin(texts),
                                      duration=durations, steps=nfe_step, cfg_strength=cfg_strength,
                                      sway_sampling_coef=sway_sampling_coef, seed=seed, return_trajectory=False)
        return out

    def results():
        """(wave, mel) per text chunk, in order"""
        if batch_chunks and len(gen_text_batches) > 1 and not streaming:
            texts = [ref_text + t for t in gen_text_batches]
            durs = torch.tensor([chunk_duration(t) for t in gen_text_batches], dtype=torch.long)
            out = sample(texts, durs.to(device))
            cond_frames = ref_audio_len + 1
            durs_eff = torch.maximum(durs, torch.tensor([len(t) for t in convert_char_to_pinyin(texts)]).clamp(min=cond_frames) + 1)
            for i in range(len(texts)):
                yield finish(out[i:i + 1, : int(min(durs_eff[i], out.shape[1]))])
        else:
            for t in gen_text_batches:
                yield finish(sample([ref_text + t], chunk_duration(t)))

    if streaming:
        for wave, _ in results():
            w = wave.cpu().numpy()
            for j in range(0, len(w), chunk_size):
                yield w[j:j + chunk_size], target_sample_rate
        return
    waves, mels = [], []
    for wave, mel in results():
        waves.append(wave)
        mels.append(mel)
    if not waves:
        yield None, target_sample_rate, None
        return
    final = ops.crossfade_concat(waves, int(cross_fade_duration * target_sample_rate) if cross_fade_duration > 0 else 0)
    yield final.cpu().numpy(), target_sample_rate, np.concatenate(mels, axis=1)


# ---- reference-audio / output-file helpers (utils_infer.py:273-360, 569-578); pydub is absent offline, so these work on PCM wav files
# and float tensors through the restated silence search of infer/f5tts_wrapper.py -------------------------------------------------
_ref_audio_cache: dict = {}


def remove_silence_edges(audio, silence_threshold=-42, sample_rate: int = target_sample_rate):
    """utils_infer.py:273-286 on a mono float tensor [samples] (the reference takes a pydub AudioSegment)"""
    from .f5tts_wrapper import remove_silence_edges as _edges
    return _edges(torch.as_tensor(audio, dtype=torch.float32).reshape(-1), sample_rate, silence_threshold)


def preprocess_ref_audio_text(ref_audio_orig, ref_text, clip_short=True, show_info=print):
    """utils_infer.py:292-360: clip the reference at a silence so that it stays under 12 s, trim the silent edges, append 50 ms of
    silence, write the result to a temporary wav and make sure the text ends with sentence punctuation.  Returns (wav path, text).
    The Whisper transcription of an empty `ref_text` is outside this path (no ASR model offline): it raises."""
    import hashlib
    import tempfile
    from .f5tts_wrapper import _read_wav, _write_wav, clip_reference, remove_silence_edges as _edges
    show_info("Converting audio...")
    audio, sr = _read_wav(str(ref_audio_orig))
    x = audio.float().mean(dim=0) if audio.shape[0] > 1 else audio[0].float()
    if clip_short:
        n0 = x.numel()
        x = clip_reference(x, sr)
        if x.numel() < n0:
            show_info("Audio is over 12s, clipping short.")
    x = torch.cat((_edges(x, sr), torch.zeros(int(0.05 * sr))))
    with tempfile.NamedTemporaryFile(delete=False, suffix=".wav") as f:
        ref_audio = f.name
    _write_wav(ref_audio, x.numpy(), sr)
    with open(ref_audio, "rb") as fh:
        audio_hash = hashlib.md5(fh.read()).hexdigest()
    if not ref_text.strip():
        if audio_hash in _ref_audio_cache:
            show_info("Using cached reference text...")
            ref_text = _ref_audio_cache[audio_hash]
        else:
            raise RuntimeError("auto-transcription needs the Whisper pipeline, which is not available offline; pass ref_text")
    else:
        show_info("Using custom reference text...")
    if not ref_text.endswith(". ") and not ref_text.endswith("。"):
        ref_text += " " if ref_text.endswith(".") else ". "
    return ref_audio, ref_text


def remove_silence_for_generated_wav(filename):
    """utils_infer.py:569-578: drop every silence of >= 1 s below -50 dBFS from a generated wav file in place, keeping 500 ms on
    each side of the cuts"""
    from .f5tts_wrapper import _read_wav, _write_wav, split_on_silence
    audio, sr = _read_wav(str(filename))
    x = audio.float().mean(dim=0) if audio.shape[0] > 1 else audio[0].float()
    pieces = split_on_silence(x, sr, min_silence_len=1000, silence_thresh=-50, keep_silence=500, seek_step=10)
    out = torch.cat(pieces) if pieces else x[:0]
    _write_wav(str(filename), out.numpy(), sr)
