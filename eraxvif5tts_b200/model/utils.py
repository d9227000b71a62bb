"""Host-side helpers with the reference's names and behaviour (/root/reference/src/f5_tts/model/utils.py)."""
from __future__ import annotations

import os
import random
from collections import defaultdict

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


def seed_everything(seed=0):
    """model/utils.py:18-25"""
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def maybe_masked_mean(t, mask=None):
    """model/utils.py:69-77: mean over the sequence axis of t [b, n, d], restricted to mask [b, n] when given"""
    if not exists(mask):
        return t.mean(dim=1)
    t = torch.where(mask[:, :, None], t, torch.zeros((), device=t.device, dtype=t.dtype))
    return t.sum(dim=1) / mask.float().sum(dim=1).clamp(min=1.0)[:, None]


def repetition_found(text, length=2, tolerance=10):
    """model/utils.py:290-298: does any `length`-gram occur more than `tolerance` times (the eval scripts' degenerate-output filter)"""
    counts = defaultdict(int)
    for i in range(len(text) - length + 1):
        counts[text[i: i + length]] += 1
    return any(c > tolerance for c in counts.values())


_CUSTOM_TRANS = str.maketrans({";": ",", "“": '"', "”": '"', "‘": "'", "’": "'"})


def convert_char_to_pinyin(text_list, polyphone=True):
    """model/utils.py:243-284.  With jieba + pypinyin present the reference algorithm runs unchanged; without them (this
    image) non-CJK text takes the same character-level path the reference produces for Latin / Vietnamese script."""
    try:
        import jieba
        from pypinyin import Style, lazy_pinyin
        if not hasattr(jieba, "cut"):
            jieba = None
    except ImportError:
        jieba = None
    out = []
    for text in text_list:
        text = text.translate(_CUSTOM_TRANS)
        if jieba is None:
            if any("㄀" <= c <= "鿿" for c in text):
                raise RuntimeError("Chinese text needs jieba + pypinyin, which are not installed in this image")
            out.append(list(text))
            continue
        char_list = []
        for seg in jieba.cut(text):
            seg_byte_len = len(bytes(seg, "UTF-8"))
            if seg_byte_len == len(seg):
                if char_list and seg_byte_len > 1 and char_list[-1] not in " :'\"":
                    char_list.append(" ")
                char_list.extend(seg)
            elif polyphone and seg_byte_len == 3 * len(seg):
                seg_ = lazy_pinyin(seg, style=Style.TONE3, tone_sandhi=True)
                for i, c in enumerate(seg):
                    if "㄀" <= c <= "鿿":
                        char_list.append(" ")
                    char_list.append(seg_[i])
            else:
                for c in seg:
                    if ord(c) < 256:
                        char_list.extend(c)
                    elif "㄀" <= c <= "鿿":
                        char_list.append(" ")
                        char_list.extend(lazy_pinyin(c, style=Style.TONE3, tone_sandhi=True))
                    else:
                        char_list.append(c)
        out.append(char_list)
    return out
