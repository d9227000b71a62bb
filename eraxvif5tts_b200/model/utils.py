"""Host-side helpers with the reference's names and behaviour (/root/reference/src/f5_tts/model/utils.py)."""
from __future__ import annotations

import os

import torch
from torch.nn.utils.rnn import pad_sequence


def exists(v):
    return v is not None


def default(v, d):
    return v if exists(v) else d


def lens_to_mask(t, length=None):
    """utils.py:42-47"""
    if not exists(length):
        length = t.amax()
    seq = torch.arange(length, device=t.device)
    return seq[None, :] < t[:, None]


def mask_from_start_end_indices(seq_len, start, end):
    """utils.py:50-55"""
    max_seq_len = seq_len.max().item()
    seq = torch.arange(max_seq_len, device=start.device).long()
    return (seq[None, :] >= start[:, None]) & (seq[None, :] < end[:, None])


def mask_from_frac_lengths(seq_len, frac_lengths):
    """utils.py:58-66"""
    lengths = (frac_lengths * seq_len).long()
    max_start = seq_len - lengths
    rand = torch.rand_like(frac_lengths)
    start = (max_start * rand).long().clamp(min=0)
    end = start + lengths
    return mask_from_start_end_indices(seq_len, start, end)


def list_str_to_tensor(text, padding_value=-1):
    """utils.py:81-84 — UTF-8 byte tokenizer"""
    rows = [torch.tensor([*bytes(t, "UTF-8")]) for t in text]
    return pad_sequence(rows, padding_value=padding_value, batch_first=True)


def list_str_to_idx(text, vocab_char_map, padding_value=-1):
    """utils.py:88-95 — unknown characters map to index 0"""
    rows = [torch.tensor([vocab_char_map.get(c, 0) for c in t], dtype=torch.long) for t in text]
    return pad_sequence(rows, padding_value=padding_value, batch_first=True)


def get_tokenizer(path_or_dataset_name, tokenizer_type="custom"):
    """utils.py:118-240, the vocab-file modes: index = line order, a first line that is a literal space is kept
    (quirk 16 in SURVEY.md §9.1).  `tokenizer_type` "custom" takes a path to vocab.txt; "byte" is the 256-symbol UTF-8 map."""
    if tokenizer_type == "byte":
        return None, 256
    path = path_or_dataset_name
    if not os.path.isfile(path):
        raise FileNotFoundError(f"vocab file not found: {path}")
    vocab_char_map = {}
    with open(path, "r", encoding="utf-8") as f:
        for i, line in enumerate(f):
            ch = line[:-1] if line.endswith("\n") else line
            vocab_char_map[ch] = i
    return vocab_char_map, len(vocab_char_map)
