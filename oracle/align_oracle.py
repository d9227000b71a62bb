"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg) -- never imported by the product path.

CPU restatement of the reference's monotonic alignment search and duration predictor:
  viterbi_alignment      model/alignment_utils.py:154-212  (viterbi_vectorized_alignment)
  windowed_alignment     model/alignment_utils.py:214-257  (windowed_monotonic_alignment)
  duration_predictor     model/duration_predictor.py:27-44 (DurationPredictor.forward, eval mode; id_shift=0: phoneme_forward :46-66)
Pinned: tests/golden/align_golden.pt holds outputs of the reference's own functions (tests/golden/make_align_golden.py runs them
from /root/reference), and tests/test_align_oracle.py checks this file against them bit for bit (alignments) / to 1e-5 (predictor)."""
import numpy as np
import torch
import torch.nn.functional as F


def viterbi_path_prob(sim: np.ndarray) -> np.ndarray:
    """fp32 cumulative path matrix, alignment_utils.py:159-175"""
    sim = np.asarray(sim, dtype=np.float32)
    b, nt, T = sim.shape
    path = np.zeros_like(sim)
    path[:, 0, :] = np.cumsum(sim[:, 0, :].astype(np.float32), axis=1, dtype=np.float32) if T > 0 else 0
    # np.cumsum in float32 adds left to right like the reference's loop
    for n in range(1, nt):
        path[:, n, 0] = path[:, n - 1, 0] + sim[:, n, 0]
        for t in range(1, T):
            path[:, n, t] = sim[:, n, t] + np.maximum(path[:, n - 1, t], path[:, n, t - 1])
    return path


def viterbi_alignment(sim: np.ndarray):
    """-> (alignment fp32 [b, nt, T], durations int [b, nt]); backtracking rule of alignment_utils.py:183-210"""
    sim = np.asarray(sim, dtype=np.float32)
    b, nt, T = sim.shape
    path = viterbi_path_prob(sim)
    align = np.zeros_like(sim)
    for i in range(b):
        curr = T - 1
        for n in range(nt - 1, -1, -1):
            boundary = 0
            if n > 0:
                costs = path[i, n, :curr + 1]
                cand = np.nonzero((costs[1:] - costs[:-1]) > 0)[0]
                if len(cand) > 0:
                    boundary = int(cand[-1])
            align[i, n, boundary:curr + 1] = 1
            curr = boundary - 1
            if curr < 0:
                break
    return align, align.sum(-1).astype(np.int64)


def windowed_alignment(sim: np.ndarray, window_size=0.2):
    """alignment_utils.py:214-257"""
    sim = np.asarray(sim, dtype=np.float32)
    b, nt, T = sim.shape
    align = np.zeros_like(sim)
    W = max(2, int(T * window_size))
    for i in range(b):
        fpp = T / nt
        start = 0
        for n in range(nt - 1):
            e = int((n + 1) * fpp)
            ws, we = max(start, e - W), min(T - 1, e + W)
            if ws > we:
                raise IndexError("empty search window")
            best_end = ws + int(np.argmax(sim[i, n, ws:we + 1]))
            align[i, n, start:best_end + 1] = 1
            start = best_end + 1
            if start >= T:
                break
        if start < T:
            align[i, -1, start:] = 1
    return align, align.sum(-1).astype(np.int64)


def duration_predictor(sd: dict, ids: torch.Tensor, mask: torch.Tensor, id_shift: int = 1, dropout=None) -> torch.Tensor:
    """sd: DurationPredictor.state_dict(); ids int [b, nt]; mask [b, nt] -> [b, 1, nt].  dropout None: eval mode (identity);
    (p, seed): train mode with the product's counter-based masks (f5_oracle.dropout_multipliers, layer 0, site 0 / 1, element index
    in [b, nt, F] order) so that autograd here sees the masks the CUDA path used."""
    def drop(x, site):
        if dropout is None or not dropout[0] > 0:
            return x
        from .f5_oracle import dropout_multipliers
        b, f, nt = x.shape
        return x * dropout_multipliers(dropout[0], dropout[1], 0, site, (b, nt, f)).transpose(1, 2)
    k = sd["conv_1.weight"].shape[-1]
    x = F.embedding(ids + id_shift, sd["text_embed.weight"].float()).transpose(1, 2)
    m = mask.float().unsqueeze(1)
    x = F.conv1d(x * m, sd["conv_1.weight"].float(), sd["conv_1.bias"].float(), padding=k // 2)
    x = drop(F.group_norm(torch.relu(x), 1, sd["norm_1.weight"].float(), sd["norm_1.bias"].float(), eps=1e-5), 0)
    x = F.conv1d(x * m, sd["conv_2.weight"].float(), sd["conv_2.bias"].float(), padding=k // 2)
    x = drop(F.group_norm(torch.relu(x), 1, sd["norm_2.weight"].float(), sd["norm_2.bias"].float(), eps=1e-5), 1)
    x = F.conv1d(x * m, sd["proj.weight"].float(), sd["proj.bias"].float())
    return x * m


def duration_loss(logw: torch.Tensor, attn: torch.Tensor, mask: torch.Tensor, per_item: bool = False) -> torch.Tensor:
    """train/distil_reload.py:1104-1115, written as the script writes it: logw [b, 1, nt] minus logw_ [b, nt] broadcasts to
    [b, b, nt] (cross terms between batch items for b > 1 -- the reference's literal behaviour).  per_item=True: the intended loss."""
    m = mask.float()
    logw_ = torch.log(attn.float().sum(dim=2) + 1e-6) * m
    if per_item:
        logw_ = logw_.unsqueeze(1)
    l_length = torch.sum((logw - logw_) ** 2, [1, 2]) / torch.sum(m)
    return torch.sum(l_length.float())
