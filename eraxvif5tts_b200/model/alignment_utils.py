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
    """alignment_utils.py:260-334: uniform segmentation refined by two greedy sweeps that move each boundary by up to
    min(5, mel_len // 20) frames when that raises the score.  The sweep is inherently sequential (every decision changes the state the
    next one reads) and touches O(nt) boundaries, so it runs on the HOST over a copy of the similarity matrix; the score of a trial
    is the item's current score plus the change over the moved columns (the reference re-sums the whole [nt, mel_len] product for each
    of the 11 trials of each boundary).  The reference's bookkeeping is kept literally: a trial of item i is compared against
    `total_score`, which starts as the sum over the WHOLE batch and becomes the accepted trial's item-level score afterwards
    (alignment_utils.py:281, 295, 315, 329-330); trial rows are overwritten, not exchanged, so a shift longer than a segment leaves a
    frame assigned to two tokens, exactly as in the reference."""
    import numpy as np
    sim = similarity_matrix.detach().to(torch.float32).cpu().numpy().astype(np.float64)
    b, nt, T = sim.shape
    align = np.zeros((b, nt, T), dtype=np.float64)
    bounds = torch.linspace(0, T, nt + 1).long().tolist()
    for n in range(nt):
        if bounds[n] < bounds[n + 1]:
            align[:, n, bounds[n]:bounds[n + 1]] = 1.0
    item = (sim * align).sum(axis=(1, 2))
    total = float(item.sum())
    shift_range = min(5, T // 20)
    for _ in range(2):
        for i in range(b):
            for n in range(nt - 1):
                row = align[i, n]
                edge = np.nonzero((row[:-1] == 1.0) & (row[1:] == 0.0))[0]
                if edge.size == 0:
                    continue
                boundary = int(edge[0])
                best_score, best_shift, best_delta = total, 0, 0.0
                for shift in range(-shift_range, shift_range + 1):
                    nb = boundary + shift
                    if not (0 <= nb < T - 1):
                        continue
                    if shift < 0:    # columns nb+1 .. boundary: row n -> 0, row n+1 -> 1
                        c = slice(nb + 1, boundary + 1)
                        delta = float(((0.0 - align[i, n, c]) * sim[i, n, c]).sum() + ((1.0 - align[i, n + 1, c]) * sim[i, n + 1, c]).sum())
                    elif shift > 0:  # columns boundary+1 .. nb: row n -> 1, row n+1 -> 0
                        c = slice(boundary + 1, nb + 1)
                        delta = float(((1.0 - align[i, n, c]) * sim[i, n, c]).sum() + ((0.0 - align[i, n + 1, c]) * sim[i, n + 1, c]).sum())
                    else:
                        delta = 0.0
                    new_score = float(item[i]) + delta
                    if new_score > best_score:
                        best_score, best_shift, best_delta = new_score, shift, delta
                if best_shift != 0:
                    nb = boundary + best_shift
                    if best_shift < 0:
                        align[i, n, nb + 1:boundary + 1] = 0.0
                        align[i, n + 1, nb + 1:boundary + 1] = 1.0
                    else:
                        align[i, n, boundary + 1:nb + 1] = 1.0
                        align[i, n + 1, boundary + 1:nb + 1] = 0.0
                    item[i] += best_delta
                    total = best_score
    return torch.from_numpy(align).to(device=similarity_matrix.device, dtype=similarity_matrix.dtype)


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
