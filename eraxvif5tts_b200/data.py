"""Host side of the training data path (/root/reference/src/f5_tts/model/dataset.py:168-239 `DynamicBatchSampler`, :308-333
`collate_fn`; consumed by the training loop at /root/reference/src/f5_tts/model/trainer.py:1191-1262).

Same batching rule and arguments as the reference: utterances sorted by frame count, greedily packed while the batch stays
within `frames_threshold` frames (and `max_samples` utterances), batches shuffled per epoch from `random_seed + epoch`.
What is re-designed for this package:
  * `collate_token_major` pads straight into ONE pinned host buffer in the token-major [b, n, n_mels] layout the kernels read
    (the reference pads channel-major and the trainer permutes on the device, trainer.py:1252), so the H2D copy of a step is a
    single async memcpy that overlaps the previous step;
  * raw-audio datasets hand over waveforms; the mel spectrogram is computed on the GPU by `MelSpec` inside
    `TrainEngine.loss_and_grads` (a 2-D input is taken as audio, like `CFM.forward`, cfm.py:221-224);
  * `shard_batches` gives each data-parallel rank every world-th batch of an epoch's order, truncated to equal length on all ranks
    (accelerate's BatchSamplerShard with even_batches under `drop_last`, which is what the reference relies on, dataset.py:221)."""
from __future__ import annotations

import torch
from torch.utils.data import Sampler


class DynamicBatchSampler(Sampler):
    """frame-budget batches; `sampler.data_source.get_frame_len(idx)` supplies the lengths (dataset.py:84-90, 137-140)"""

    def __init__(self, sampler, frames_threshold: int, max_samples: int = 0, random_seed=None, drop_residual: bool = False):
        self.sampler = sampler
        self.frames_threshold = frames_threshold
        self.max_samples = max_samples
        self.random_seed = random_seed
        self.epoch = 0
        data_source = sampler.data_source
        indices = sorted(((idx, data_source.get_frame_len(idx)) for idx in sampler), key=lambda e: e[1])  # stable, like list.sort
        batches, batch, batch_frames = [], [], 0
        for idx, frame_len in indices:
            if batch_frames + frame_len <= frames_threshold and (max_samples == 0 or len(batch) < max_samples):
                batch.append(idx)
                batch_frames += frame_len
            else:
                if batch:
                    batches.append(batch)
                if frame_len <= frames_threshold:
                    batch, batch_frames = [idx], frame_len
                else:  # an utterance longer than the budget is dropped
                    batch, batch_frames = [], 0
        if not drop_residual and batch:
            batches.append(batch)
        self.batches = batches
        self.drop_last = True

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def __iter__(self):
        if self.random_seed is not None:
            g = torch.Generator()
            g.manual_seed(self.random_seed + self.epoch)
            order = torch.randperm(len(self.batches), generator=g).tolist()
            return iter([self.batches[i] for i in order])
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


def shard_batches(batches: list, rank: int, world: int) -> list:
    """rank's share of one epoch's batch order: batches rank, rank + world, ...; every rank gets the same count (the tail that
    does not fill a round is dropped), so the per-step all-reduce never waits on a rank without data"""
    if world <= 1:
        return list(batches)
    rounds = len(batches) // world
    return [batches[r * world + rank] for r in range(rounds)]


def collate_fn(batch):
    """the reference's collate (dataset.py:308-333): mel channel-major [b, n_mels, max_len], zero padded"""
    mel_specs = [item["mel_spec"].squeeze(0) for item in batch]
    mel_lengths = torch.LongTensor([spec.shape[-1] for spec in mel_specs])
    max_len = int(mel_lengths.amax())
    mel = torch.stack([torch.nn.functional.pad(spec, (0, max_len - spec.size(-1)), value=0) for spec in mel_specs])
    text = [item["text"] for item in batch]
    return dict(mel=mel, mel_lengths=mel_lengths, text=text, text_lengths=torch.LongTensor([len(t) for t in text]),
                phoneme=[item.get("phoneme", []) for item in batch])


def collate_token_major(batch, pin: bool = True):
    """same items, but padded into one (pinned) token-major buffer [b, max_len, n_mels] — what `TrainEngine.loss_and_grads` reads —
    so no permute / contiguous pass runs on the device and the host-to-device copy is one async memcpy"""
    mel_specs = [item["mel_spec"].squeeze(0) for item in batch]  # [n_mels, T_i]
    mel_lengths = torch.LongTensor([spec.shape[-1] for spec in mel_specs])
    max_len = int(mel_lengths.amax())
    n_mels = mel_specs[0].shape[0]
    mel = torch.zeros(len(batch), max_len, n_mels, dtype=torch.float32)
    # never pin inside a DataLoader worker: cudaHostAlloc in a forked child of a CUDA-initialised parent fails; there the batch is
    # pinned by the loader's pin-memory thread in the main process (Trainer passes pin_memory=True when num_workers > 0)
    if pin and torch.utils.data.get_worker_info() is None and torch.cuda.is_available():
        mel = mel.pin_memory()
    for i, spec in enumerate(mel_specs):
        mel[i, :spec.shape[-1]] = spec.transpose(0, 1)
    text = [item["text"] for item in batch]
    return dict(mel=mel, mel_lengths=mel_lengths, text=text, text_lengths=torch.LongTensor([len(t) for t in text]),
                phoneme=[item.get("phoneme", []) for item in batch])
