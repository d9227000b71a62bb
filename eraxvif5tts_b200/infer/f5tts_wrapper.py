"""F5TTSWrapper with the reference's public surface (/root/reference/src/f5_tts/infer/f5tts_wrapper.py:28-621):
__init__ / preprocess_reference / generate / get_current_audio_length and the attributes its callers touch
(ref_audio_processed, ref_text, ref_audio_len, target_sample_rate, device, use_duration_predictor, model, vocoder).

Differences, all host-side and documented in DESIGN.md: no network / hydra / pydub / Whisper in this image, so the checkpoint
path is optional (random init without it), model configs are constants, reference audio may be given as a tensor, and the
auto-transcription branch raises.  `generate(..., batch_chunks=True)` (extension) runs all text chunks of a
request as ONE ragged sample() batch instead of the reference's serial B=1 loop."""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from ..model import CFM, DiT
from ..model.utils import convert_char_to_pinyin, get_tokenizer  # noqa: F401  (convert_char_to_pinyin: re-exported, older import path)
from .. import _lib as L
from .utils_infer import chunk_text, load_checkpoint, load_vocoder, resolve_arch



def _read_wav(path: str):
    """PCM wav -> (float32 [channels, samples], sample_rate) without torchaudio backends."""
    try:
        import torchaudio
        return torchaudio.load(path)
    except Exception:  # noqa: BLE001
        import wave
        with wave.open(path, "rb") as w:
            sr, ch, sw, nf = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
            raw = w.readframes(nf)
        if sw == 2:
            a = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        elif sw == 4:
            a = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
        elif sw == 1:
            a = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        else:
            raise ValueError(f"unsupported wav sample width {sw}")
        return torch.from_numpy(a.reshape(-1, ch).T.copy()), sr


def _ms_to_samples(ms: float, sr: int) -> int:
    return int(ms * sr / 1000.0)  # pydub: AudioSegment._parse_position -> int(frame_count(ms=...))


def detect_silence(x: torch.Tensor, sr: int, min_silence_len: int = 1000, silence_thresh: float = -16.0, seek_step: int = 1):
    """pydub.silence.detect_silence (third party, absent offline; restated from its published algorithm) on a mono float signal
    `x` [samples] in [-1, 1]: windows of `min_silence_len` ms starting every `seek_step` ms whose RMS is at or below
    10^(silence_thresh / 20) of full scale, merged into [start_ms, end_ms] ranges exactly like pydub does.  The window RMS comes from
    one cumulative sum of squares (pydub: audioop.rms per slice)."""
    seg_len = int(round(1000.0 * x.numel() / sr))  # len(AudioSegment) in ms
    if seg_len < min_silence_len:
        return []
    thresh = 10.0 ** (silence_thresh / 20.0)
    last = seg_len - min_silence_len
    starts = list(range(0, last + 1, seek_step))
    if last % seek_step:
        starts.append(last)
    csum = torch.cat((torch.zeros(1, dtype=torch.float64), torch.cumsum(x.double().square(), 0)))
    s0 = torch.tensor([_ms_to_samples(i, sr) for i in starts]).clamp(max=x.numel())
    s1 = torch.tensor([_ms_to_samples(i + min_silence_len, sr) for i in starts]).clamp(max=x.numel())
    cnt = (s1 - s0).clamp(min=1).double()
    rms = torch.sqrt((csum[s1] - csum[s0]) / cnt)
    silent = [starts[k] for k in torch.nonzero(rms <= thresh).flatten().tolist()]
    if not silent:
        return []
    ranges = []
    prev = silent.pop(0)
    cur = prev
    for st in silent:
        continuous = st == prev + seek_step
        has_gap = st > prev + min_silence_len
        if not continuous and has_gap:
            ranges.append([cur, prev + min_silence_len])
            cur = st
        prev = st
    ranges.append([cur, prev + min_silence_len])
    return ranges


def split_on_silence(x: torch.Tensor, sr: int, min_silence_len: int = 1000, silence_thresh: float = -16.0, keep_silence: int = 100,
                     seek_step: int = 1):
    """pydub.silence.split_on_silence restated (see detect_silence): the non-silent stretches, each padded by `keep_silence` ms of the
    surrounding silence (overlapping paddings meet half way).  Returns the list of signal pieces."""
    seg_len = int(round(1000.0 * x.numel() / sr))
    silent = detect_silence(x, sr, min_silence_len, silence_thresh, seek_step)
    if not silent:
        nonsilent = [[0, seg_len]]
    elif silent[0][0] == 0 and silent[0][1] == seg_len:
        nonsilent = []
    else:
        prev_end, nonsilent = 0, []
        for st, en in silent:
            nonsilent.append([prev_end, st])
            prev_end = en
        if en != seg_len:
            nonsilent.append([prev_end, seg_len])
        if nonsilent[0] == [0, 0]:
            nonsilent.pop(0)
    out = [[st - keep_silence, en + keep_silence] for st, en in nonsilent]
    for a_, b_ in zip(out, out[1:]):
        if b_[0] < a_[1]:
            a_[1] = (a_[1] + b_[0]) // 2
            b_[0] = a_[1]
    return [x[_ms_to_samples(max(st, 0), sr): _ms_to_samples(min(en, seg_len), sr)] for st, en in out]


def clip_reference(x: torch.Tensor, sr: int) -> torch.Tensor:
    """The reference's clip_short rule (f5tts_wrapper.py:272-301): cut at a long silence (>= 1 s below -50 dBFS) once more than 6 s
    are collected and the next piece would pass 12 s; if that still leaves more than 12 s, the same with short silences (>= 100 ms
    below -40 dBFS); if that fails too, a hard cut at 12 s."""
    def ms(t):
        return int(round(1000.0 * t.numel() / sr))

    def collect(pieces):
        wave = x[:0]
        for seg in pieces:
            if ms(wave) > 6000 and ms(wave) + ms(seg) > 12000:
                break
            wave = torch.cat((wave, seg))
        return wave
    wave = collect(split_on_silence(x, sr, min_silence_len=1000, silence_thresh=-50, keep_silence=1000, seek_step=10))
    if ms(wave) > 12000:
        wave = collect(split_on_silence(x, sr, min_silence_len=100, silence_thresh=-40, keep_silence=1000, seek_step=10))
    if ms(wave) > 12000:
        wave = wave[: _ms_to_samples(12000, sr)]
    return wave


def remove_silence_edges(x: torch.Tensor, sr: int, silence_threshold: float = -42.0) -> torch.Tensor:
    """remove_silence_edges / F5TTSWrapper._remove_silence_edges (utils_infer.py:273-286, f5tts_wrapper.py:356-378) restated on a mono
    float signal `x` [samples]: pydub's detect_leading_silence walks 10 ms chunks from the start while their level is below the
    threshold; the tail is walked in 1 ms slices from the end until one is louder than the threshold, and the cut lands on
    int((duration_seconds - 0.001 * slices) * 1000) ms.  Levels are RMS dBFS of the slice (pydub: audioop.rms on the PCM)."""
    def levels(sig, step_ms):
        seg_len = int(round(1000.0 * sig.numel() / sr))
        if seg_len == 0:
            return seg_len, torch.empty(0, dtype=torch.float64)
        csum = torch.cat((torch.zeros(1, dtype=torch.float64), torch.cumsum(sig.double().square(), 0)))
        starts = torch.arange(0, seg_len, step_ms)
        s0 = torch.tensor([_ms_to_samples(int(i), sr) for i in starts]).clamp(max=sig.numel())
        s1 = torch.tensor([_ms_to_samples(min(int(i) + step_ms, seg_len), sr) for i in starts]).clamp(max=sig.numel())
        ms_ = (csum[s1] - csum[s0]) / (s1 - s0).clamp(min=1).double()
        return seg_len, 10.0 * torch.log10(ms_.clamp(min=1e-30))  # an all-zero slice is -inf dBFS in pydub: far below any threshold

    seg_len, db10 = levels(x, 10)
    loud = (db10 >= silence_threshold).nonzero()  # the walk continues while dBFS < threshold
    lead_ms = min(int(loud[0]) * 10 if loud.numel() else ((seg_len + 9) // 10) * 10, seg_len)
    x = x[_ms_to_samples(lead_ms, sr):]
    seg_len, db1 = levels(x, 1)
    loud = (db1 > silence_threshold).nonzero()
    quiet_tail = seg_len - 1 - int(loud[-1]) if loud.numel() else seg_len
    dur = x.numel() / sr
    for _ in range(quiet_tail):  # the reference subtracts 0.001 per slice in floating point; int(dur * 1000) depends on that rounding
        dur -= 0.001
    end_ms = int(dur * 1000)
    return x[: _ms_to_samples(max(end_ms, 0), sr)]


def _trim_silence_edges(audio: torch.Tensor, sr: int, threshold_db: float = -42.0) -> torch.Tensor:
    """[1, samples] form of remove_silence_edges"""
    return remove_silence_edges(audio[0], sr, threshold_db).unsqueeze(0)


class F5TTSWrapper:
    def __init__(self, model_name: str = "F5TTS_v1_Base", ckpt_path: Optional[str] = None, vocab_file: Optional[str] = None,
                 vocoder_name: str = "vocos", use_local_vocoder: bool = False, vocoder_path: Optional[str] = None,
                 device: Optional[str] = None, hf_cache_dir: Optional[str] = None, target_sample_rate: int = 24000,
                 n_mel_channels: int = 100, hop_length: int = 256, win_length: int = 1024, n_fft: int = 1024,
                 ode_method: str = "euler", use_ema: bool = True, use_duration_predictor: bool = False,
                 vocab_char_map: Optional[dict] = None):
        if device is None:
            device = "cuda"
        if not str(device).startswith("cuda"):
            raise RuntimeError("eraxvif5tts_b200 runs on a B200 only (device must be cuda); there is no CPU path")
        self.device = device
        self.target_sample_rate, self.n_mel_channels = target_sample_rate, n_mel_channels
        self.hop_length, self.win_length, self.n_fft = hop_length, win_length, n_fft
        self.mel_spec_type = vocoder_name
        self.ode_method = ode_method
        self.use_duration_predictor = use_duration_predictor
        arch = resolve_arch(model_name)
        if vocab_char_map is not None:
            self.vocab_char_map, vocab_size = vocab_char_map, len(vocab_char_map)
        elif vocab_file is not None:
            self.vocab_char_map, vocab_size = get_tokenizer(vocab_file, "custom")
        else:
            raise ValueError("vocab_file (or vocab_char_map) is required: the reference's bundled vocab.txt is not shipped here")
        self.model = CFM(
            transformer=DiT(**arch, text_num_embeds=vocab_size, mel_dim=n_mel_channels),
            mel_spec_kwargs=dict(n_fft=n_fft, hop_length=hop_length, win_length=win_length, n_mel_channels=n_mel_channels,
                                 target_sample_rate=target_sample_rate, mel_spec_type=vocoder_name),
            odeint_kwargs=dict(method=ode_method), vocab_char_map=self.vocab_char_map).to(self.device)
        if ckpt_path is not None:
            load_checkpoint(self.model, ckpt_path, self.device, use_ema=use_ema)
        # the duration-predictor variant of the wrapper (model/f5tts_wrapper-dur_pred.py:166-230): attach one with
        # attach_duration_predictor(); checkpoints that bundle it are not auto-detected
        self.has_duration_predictor = False
        self.duration_predictor, self._dp_tokenizer = None, None
        if self.use_duration_predictor:
            print("Warning: Duration predictor requested but not found in model. Using fallback duration calculation.")
            self.use_duration_predictor = False
        self.vocoder = load_vocoder(vocoder_name=vocoder_name, is_local=use_local_vocoder or vocoder_path is not None,
                                    local_path=vocoder_path or "", device=self.device, hf_cache_dir=hf_cache_dir)
        self.ref_audio_processed = None
        self.ref_text = None
        self.ref_audio_len = None
        self.target_rms = 0.1
        self.cross_fade_duration = 0.15
        self.nfe_step = 32
        self.cfg_strength = 2.0
        self.sway_sampling_coef = -1.0
        self.speed = 1.0
        self.fix_duration = None

    # ------------------------------------------------------------------------------------------------------------------
    def preprocess_reference(self, ref_audio_path, ref_text: str = "", clip_short: bool = True, sample_rate: Optional[int] = None):
        """f5tts_wrapper.py:256-354.  `ref_audio_path` may also be a float tensor / ndarray [samples] or [channels, samples]
        (then pass `sample_rate`)."""
        if isinstance(ref_audio_path, (str, os.PathLike)):
            audio, sr = _read_wav(str(ref_audio_path))
        else:
            audio = torch.as_tensor(ref_audio_path, dtype=torch.float32)
            if audio.ndim == 1:
                audio = audio.unsqueeze(0)
            sr = sample_rate or self.target_sample_rate
        audio = audio.float()
        if audio.shape[0] > 1:
            audio = torch.mean(audio, dim=0, keepdim=True)
        if clip_short:  # silence-aware clipping, f5tts_wrapper.py:272-301 (pydub's split_on_silence restated on the tensor)
            audio = clip_reference(audio[0], sr).unsqueeze(0)
        audio = _trim_silence_edges(audio, sr)
        audio = torch.cat((audio, torch.zeros(1, int(0.05 * sr))), dim=-1)  # + AudioSegment.silent(duration=50)
        if not ref_text.strip():
            raise RuntimeError("auto-transcription needs the Whisper pipeline, which is not available offline; pass ref_text")
        if not ref_text.endswith(". ") and not ref_text.endswith("。"):
            ref_text += " " if ref_text.endswith(".") else ". "
        rms = torch.sqrt(torch.mean(torch.square(audio)))
        if rms < self.target_rms:
            audio = audio * self.target_rms / rms
        if sr != self.target_sample_rate:
            import torchaudio
            audio = torchaudio.transforms.Resample(sr, self.target_sample_rate)(audio)
        audio = audio.to(self.device)
        self.ref_audio_processed = audio
        self.ref_text = ref_text
        self.ref_audio_len = audio.shape[-1] // self.hop_length
        return audio, ref_text

    # ------------------------------------------------------------------------------------------------------------------
    def _remove_silence_edges(self, audio, silence_threshold=-42, sample_rate: Optional[int] = None):
        """f5tts_wrapper.py:356-378 on a float tensor [samples] or [1, samples] instead of a pydub AudioSegment"""
        a = torch.as_tensor(audio, dtype=torch.float32)
        sr = sample_rate or self.target_sample_rate
        return _trim_silence_edges(a, sr, silence_threshold) if a.ndim == 2 else remove_silence_edges(a, sr, silence_threshold)

    def attach_duration_predictor(self, duration_predictor, tokenizer=None, use: bool = True):
        """model/f5tts_wrapper-dur_pred.py:169-230.  `tokenizer(text) -> (ids: list[int], is_phoneme: bool)` turns a text chunk into
        the predictor's input; the reference phonemizes with espeak (alignment_utils.py:39-58, not available offline) and calls
        phoneme_forward.  Default: the model's own character tokenizer (vocab_char_map) and DurationPredictor.forward, the input
        the predictor is trained on in train/distil_reload.py:1096-1110."""
        self.duration_predictor = duration_predictor.to(self.device).eval()
        self._dp_tokenizer = tokenizer
        self.has_duration_predictor = True
        self.use_duration_predictor = bool(use)

    def calculate_duration_with_predictor(self, text_input, local_speed=1.0):
        """model/f5tts_wrapper-dur_pred.py:441-519: ref frames + int(sum(exp(clamp(logw, -20, 20))) / local_speed), falling back to
        the text-length ratio when the predictor path raises"""
        if not self.has_duration_predictor:
            raise ValueError("Duration predictor not available")
        try:
            if self._dp_tokenizer is not None:
                ids, is_phoneme = self._dp_tokenizer(text_input)
                ids = torch.tensor([list(ids)], dtype=torch.long, device=self.device)
            else:
                from ..model.utils import list_str_to_idx
                ids = list_str_to_idx([text_input], self.vocab_char_map).to(self.device)
                is_phoneme = False
            mask = torch.ones_like(ids)
            fwd = self.duration_predictor.phoneme_forward if is_phoneme else self.duration_predictor
            log_durations = fwd(ids, mask)
            if log_durations.dim() == 3 and log_durations.size(1) == 1:
                log_durations = log_durations.squeeze(1)
            durations = torch.exp(torch.clamp(log_durations, -20, 20)).sum(dim=1)
            return self.ref_audio_len + int(durations[0].item() / local_speed)
        except Exception as e:  # the reference falls back to the ratio rule on any failure (:504-519)
            print(f"Error in duration prediction: {e}")
            text_len = len(text_input.encode("utf-8")) if isinstance(text_input, str) else sum(len(str(t).encode("utf-8")) for t in text_input)
            ref_text_len = len(self.ref_text.encode("utf-8"))
            return self.ref_audio_len + int(self.ref_audio_len / ref_text_len * text_len / local_speed)

    def _chunk_duration(self, text_batch: str, speed: float, fix_duration, use_predictor: bool = False):
        local_speed = 0.3 if len(text_batch.encode("utf-8")) < 10 else speed
        if fix_duration is not None:
            return int(fix_duration * self.target_sample_rate / self.hop_length)
        if use_predictor:
            return self.calculate_duration_with_predictor(text_batch, local_speed)
        ref_text_len = len(self.ref_text.encode("utf-8"))
        gen_text_len = len(text_batch.encode("utf-8"))
        return self.ref_audio_len + int(self.ref_audio_len / ref_text_len * gen_text_len / local_speed)

    @L.on_own_device
    def generate(self, text: str, output_path: Optional[str] = None, nfe_step: Optional[int] = None,
                 cfg_strength: Optional[float] = None, sway_sampling_coef: Optional[float] = None, speed: Optional[float] = None,
                 fix_duration: Optional[float] = None, cross_fade_duration: Optional[float] = None,
                 use_duration_predictor: Optional[bool] = None, return_numpy: bool = False, return_spectrogram: bool = False,
                 batch_chunks: bool = False, seed: Optional[int] = None, device_crossfade: bool = False, return_pcm16: bool = False):
        """f5tts_wrapper.py:408-607.  Beyond the reference: batch_chunks samples all chunks as one ragged batch;
        device_crossfade keeps the chunk waveforms on the GPU and cross-fades them there (float32 result; the reference's numpy
        fold returns float64 after the first blend); return_pcm16 (implies device_crossfade) returns np.int16(wave * 32767)
        packed on the device, as the socket server streams it (socket_server.py:54)."""
        device_crossfade = device_crossfade or return_pcm16
        if self.ref_audio_processed is None or self.ref_text is None:
            raise ValueError("Reference audio not preprocessed. Call preprocess_reference() first.")
        nfe_step = nfe_step if nfe_step is not None else self.nfe_step
        cfg_strength = cfg_strength if cfg_strength is not None else self.cfg_strength
        sway_sampling_coef = sway_sampling_coef if sway_sampling_coef is not None else self.sway_sampling_coef
        speed = speed if speed is not None else self.speed
        fix_duration = fix_duration if fix_duration is not None else self.fix_duration
        cross_fade_duration = cross_fade_duration if cross_fade_duration is not None else self.cross_fade_duration
        use_predictor = use_duration_predictor if use_duration_predictor is not None else self.use_duration_predictor
        can_use_predictor = bool(use_predictor and self.has_duration_predictor)

        audio_len = self.ref_audio_processed.shape[-1] / self.target_sample_rate
        max_chars = int(len(self.ref_text.encode("utf-8")) / audio_len * (22 - audio_len))
        text_batches = chunk_text(text, max_chars=max_chars)
        if not text_batches:
            raise RuntimeError("No audio generated")
        rms = torch.sqrt(torch.mean(torch.square(self.ref_audio_processed)))
        generated_waves, spectrograms = [], []

        def finish(gen_mel):  # gen_mel [1, n, mel] (reference part still attached)
            g = gen_mel.to(torch.float32)[:, self.ref_audio_len:, :].permute(0, 2, 1)
            wave = self.vocoder.decode(g)
            if rms < self.target_rms:
                wave = wave * rms / self.target_rms
            generated_waves.append(wave.reshape(-1) if device_crossfade else wave.squeeze().cpu().numpy())
            if return_spectrogram or output_path is not None:
                spectrograms.append(g.squeeze().cpu().numpy())

        with torch.inference_mode():
            if batch_chunks and len(text_batches) > 1:
                texts = convert_char_to_pinyin([self.ref_text + tb for tb in text_batches])
                durs = torch.tensor([self._chunk_duration(tb, speed, fix_duration, can_use_predictor) for tb in text_batches], dtype=torch.long)
                cond = self.ref_audio_processed.expand(len(text_batches), -1)
                generated, _ = self.model.sample(cond=cond, text=texts, duration=durs, steps=nfe_step, cfg_strength=cfg_strength,
                                                 sway_sampling_coef=sway_sampling_coef, seed=seed, return_trajectory=False)
                cond_frames = self.ref_audio_len + 1  # mel frames of the reference (1 + L // hop)
                durs_eff = torch.maximum(durs, torch.tensor([len(t) for t in texts]).clamp(min=cond_frames) + 1)  # cfm.py:132-136
                for i in range(len(text_batches)):
                    finish(generated[i:i + 1, : int(min(durs_eff[i], generated.shape[1]))])
            else:
                for text_batch in text_batches:
                    final_text_list = convert_char_to_pinyin([self.ref_text + text_batch])
                    duration = self._chunk_duration(text_batch, speed, fix_duration, can_use_predictor)
                    generated, _ = self.model.sample(cond=self.ref_audio_processed, text=final_text_list, duration=duration,
                                                     steps=nfe_step, cfg_strength=cfg_strength,
                                                     sway_sampling_coef=sway_sampling_coef, seed=seed, return_trajectory=False)
                    finish(generated)

        # cross-fade (f5tts_wrapper.py:542-575)
        if device_crossfade:
            from .. import ops
            final_dev = ops.crossfade_concat(generated_waves, int(cross_fade_duration * self.target_sample_rate) if cross_fade_duration > 0 else 0)
            pcm = ops.pcm16(final_dev).cpu().numpy() if return_pcm16 else None
            final_wave = final_dev.cpu().numpy() if (pcm is None or output_path is not None) else None
        elif cross_fade_duration <= 0:
            final_wave = np.concatenate(generated_waves)
        else:
            final_wave = generated_waves[0]
            for i in range(1, len(generated_waves)):
                prev_wave, next_wave = final_wave, generated_waves[i]
                cfs = min(int(cross_fade_duration * self.target_sample_rate), len(prev_wave), len(next_wave))
                if cfs <= 0:
                    final_wave = np.concatenate([prev_wave, next_wave])
                    continue
                fade_out, fade_in = np.linspace(1, 0, cfs), np.linspace(0, 1, cfs)
                overlap = prev_wave[-cfs:] * fade_out + next_wave[:cfs] * fade_in
                final_wave = np.concatenate([prev_wave[:-cfs], overlap, next_wave[cfs:]])
        combined_spectrogram = np.concatenate(spectrograms, axis=1) if spectrograms else None
        if output_path is not None:
            output_dir = os.path.dirname(output_path)
            if output_dir and not os.path.exists(output_dir):
                os.makedirs(output_dir)
            _write_wav(output_path, final_wave, self.target_sample_rate)
            if not return_numpy:
                return output_path
        if return_pcm16:
            final_wave = pcm
        if return_spectrogram:
            return final_wave, self.target_sample_rate, combined_spectrogram
        return final_wave, self.target_sample_rate

    @L.on_own_device
    def generate_many(self, texts, nfe_step: Optional[int] = None, cfg_strength: Optional[float] = None,
                      sway_sampling_coef: Optional[float] = None, speed: Optional[float] = None, fix_duration: Optional[float] = None,
                      cross_fade_duration: Optional[float] = None, use_duration_predictor: Optional[bool] = None,
                      seed: Optional[int] = None, return_pcm16: bool = False, max_batch_frames: int = 65536):
        """Cross-request batcher (no counterpart in the reference, whose servers call `generate` one request at a
        time): the chunks of ALL `texts` (same voice = this wrapper's preprocessed reference) are sorted by duration and packed into
        ragged `CFM.sample` batches of at most `max_batch_frames` padded frames (length bucketing bounds the padding waste), vocoded,
        and cross-faded per request on the GPU.  Returns a list of (wave, sample_rate), float32 or — `return_pcm16` — int16, in the
        order of `texts`.  Every chunk is sampled exactly as `generate(..., batch_chunks=True)` samples it (same duration rule, same
        per-item noise for a given seed); only the composition of the batches differs."""
        from .. import ops
        if self.ref_audio_processed is None or self.ref_text is None:
            raise ValueError("Reference audio not preprocessed. Call preprocess_reference() first.")
        if isinstance(texts, str):
            raise TypeError("generate_many takes a list of request texts; use generate() for one request")
        nfe_step = nfe_step if nfe_step is not None else self.nfe_step
        cfg_strength = cfg_strength if cfg_strength is not None else self.cfg_strength
        sway_sampling_coef = sway_sampling_coef if sway_sampling_coef is not None else self.sway_sampling_coef
        speed = speed if speed is not None else self.speed
        fix_duration = fix_duration if fix_duration is not None else self.fix_duration
        cross_fade_duration = cross_fade_duration if cross_fade_duration is not None else self.cross_fade_duration
        use_predictor = use_duration_predictor if use_duration_predictor is not None else self.use_duration_predictor
        can_use_predictor = bool(use_predictor and self.has_duration_predictor)
        audio_len = self.ref_audio_processed.shape[-1] / self.target_sample_rate
        max_chars = int(len(self.ref_text.encode("utf-8")) / audio_len * (22 - audio_len))
        items = []  # (request, chunk index, chunk text, duration)
        for r, text in enumerate(texts):
            chunks = chunk_text(text, max_chars=max_chars)
            if not chunks:
                raise RuntimeError(f"No audio generated for request {r}")
            for ci, tb in enumerate(chunks):
                items.append((r, ci, tb, int(self._chunk_duration(tb, speed, fix_duration, can_use_predictor))))
        batches = plan_ragged_batches([it[3] for it in items], max_batch_frames)
        rms = torch.sqrt(torch.mean(torch.square(self.ref_audio_processed)))
        cond_frames = self.ref_audio_len + 1
        waves = [dict() for _ in texts]
        with torch.inference_mode():
            for idxs in batches:
                tbs = [items[i][2] for i in idxs]
                tx = convert_char_to_pinyin([self.ref_text + tb for tb in tbs])
                durs = torch.tensor([items[i][3] for i in idxs], dtype=torch.long)
                generated, _ = self.model.sample(cond=self.ref_audio_processed.expand(len(idxs), -1), text=tx, duration=durs, steps=nfe_step,
                                                 cfg_strength=cfg_strength, sway_sampling_coef=sway_sampling_coef, seed=seed,
                                                 return_trajectory=False)
                durs_eff = torch.maximum(durs, torch.tensor([len(t) for t in tx]).clamp(min=cond_frames) + 1)  # cfm.py:132-136
                for k, i in enumerate(idxs):
                    g = generated[k:k + 1, : int(min(durs_eff[k], generated.shape[1]))].to(torch.float32)[:, self.ref_audio_len:, :]
                    wave = self.vocoder.decode(g.permute(0, 2, 1))
                    if rms < self.target_rms:
                        wave = wave * rms / self.target_rms
                    waves[items[i][0]][items[i][1]] = wave.reshape(-1)
        cfs = int(cross_fade_duration * self.target_sample_rate) if cross_fade_duration > 0 else 0
        out = []
        for w in waves:
            final = ops.crossfade_concat([w[ci] for ci in range(len(w))], cfs)
            out.append(((ops.pcm16(final) if return_pcm16 else final).cpu().numpy(), self.target_sample_rate))
        return out

    def get_current_audio_length(self):
        if self.ref_audio_processed is None:
            return 0
        return self.ref_audio_processed.shape[-1] / self.target_sample_rate


def plan_ragged_batches(durations, max_batch_frames: int):
    """Length-bucketed packing for the cross-request batcher: items sorted by duration (descending, stable) and cut greedily so that
    a batch's PADDED size, len(batch) * its longest duration, stays within `max_batch_frames` (a single item longer than the budget
    is a batch of its own).  Returns lists of item indices; every index appears exactly once."""
    order = sorted(range(len(durations)), key=lambda i: -int(durations[i]))
    batches, cur = [], []
    for i in order:
        longest = int(durations[cur[0]]) if cur else int(durations[i])
        if cur and (len(cur) + 1) * longest > max_batch_frames:
            batches.append(cur)
            cur = []
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


def _write_wav(path: str, wave_f32: np.ndarray, sr: int):
    import wave
    pcm = (np.clip(wave_f32, -1.0, 1.0) * 32767.0).astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(pcm.tobytes())
