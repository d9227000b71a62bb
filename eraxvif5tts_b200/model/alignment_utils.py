"""Monotonic alignment search on the GPU, mirroring the reference's model/alignment_utils.py:123-133, :154-257, :337-355 (the
phonemizer front end of that file -- viphoneme / espeak -- is host text processing and is not rebuilt).

Same function names, arguments and return values: `similarity_matrix` [b, nt, mel_len] -> hard alignment [b, nt, mel_len] of 0 / 1
in the input's dtype.  The reference runs nt x mel_len tiny torch launches per call in a Python double loop; here the recurrence is
one kernel sweeping the lattice's anti-diagonals per batch item (csrc/align.cu), bit-exact with the reference in fp32."""
import torch

from .. import _lib as L

f32 = torch.float32


def _prep(similarity_matrix):
    if similarity_matrix.device.type != "cuda":
        raise L.F5bError("alignment search needs a CUDA tensor (B200); there is no CPU fallback")
    if similarity_matrix.ndim != 3:
        raise ValueError("similarity_matrix must be [b, nt, mel_len]")
    return similarity_matrix.to(f32).contiguous()


def viterbi_vectorized_alignment(similarity_matrix, return_durations: bool = False):
    """alignment_utils.py:154-212"""
    sim = _prep(similarity_matrix)
    b, nt, T = sim.shape
    path = torch.empty_like(sim)
    align = torch.empty_like(sim)
    dur = torch.empty(b, nt, dtype=torch.int32, device=sim.device)
    L.check(L.load().f5b_align_viterbi(sim.data_ptr(), path.data_ptr(), align.data_ptr(), dur.data_ptr(), b, nt, T, L.stream()),
            "f5b_align_viterbi")
    align = align.to(similarity_matrix.dtype)
    return (align, dur) if return_durations else align


def windowed_monotonic_alignment(similarity_matrix, window_size=0.2, return_durations: bool = False):
    """alignment_utils.py:214-257"""
    sim = _prep(similarity_matrix)
    b, nt, T = sim.shape
    actual_window = max(2, int(T * window_size))
    align = torch.empty_like(sim)
    dur = torch.empty(b, nt, dtype=torch.int32, device=sim.device)
    err = torch.empty(b, dtype=torch.int32, device=sim.device)
    L.check(L.load().f5b_align_window(sim.data_ptr(), align.data_ptr(), dur.data_ptr(), err.data_ptr(), b, nt, T, actual_window,
                                      L.stream()), "f5b_align_window")
    if bool(err.any()):  # the reference's torch.argmax raises on the empty window
        raise IndexError("windowed_monotonic_alignment: empty search window (argmax of an empty tensor in the reference)")
    align = align.to(similarity_matrix.dtype)
    return (align, dur) if return_durations else align


def progressive_monotonic_alignment(similarity_matrix):
    raise NotImplementedError("the 'progressive' refinement (alignment_utils.py:260-334) is not built; use 'viterbi' or 'window'")


def monotonic_alignment_search(similarity_matrix, algorithm="viterbi"):
    """alignment_utils.py:337-355"""
    if algorithm == "viterbi":
        return viterbi_vectorized_alignment(similarity_matrix)
    if algorithm == "window":
        return windowed_monotonic_alignment(similarity_matrix)
    if algorithm == "progressive":
        return progressive_monotonic_alignment(similarity_matrix)
    raise ValueError(f"unsupported algorithm: {algorithm}. Choose one of 'viterbi', 'window', 'progressive'")


def get_durations_from_alignment(alignment):
    """alignment_utils.py:123-133"""
    return alignment.sum(dim=2)
