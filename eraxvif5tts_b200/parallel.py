"""Utterance-level data parallelism for inference: one process per GPU, independent utterances, NO collective on the sampling
path (precedent: accelerator.split_between_processes in /root/reference/src/f5_tts/eval/eval_infer_batch.py:160-196).
The only cross-rank operations are a start/end barrier and gathering small per-rank results."""
from __future__ import annotations

import os
from typing import Sequence


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_indices(n_items: int, rank: int, world: int, batch_size: int = 1) -> list[list[int]]:
    """Batches of utterance indices for `rank`: the item list is cut into batches of `batch_size` (in order, so a caller that
    sorted by length keeps its buckets, cf. eval/utils_eval.py:154,201-202) and batches are dealt round-robin to ranks."""
    batches = [list(range(s, min(s + batch_size, n_items))) for s in range(0, n_items, batch_size)]
    return batches[rank::world]


def shard(items: Sequence, rank: int, world: int, batch_size: int = 1) -> list[list]:
    return [[items[i] for i in b] for b in shard_indices(len(items), rank, world, batch_size)]


def gather_counts(local_count: int, group=None) -> list[int]:
    """all ranks' item counts (host-side, tiny): used to compute whole-job throughput = sum(frames) / max(rank time)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [local_count]
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([local_count], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return [int(x) for x in out]


def max_over_ranks(value: float, group=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return value
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t)


def allreduce_flat_(flat, group=None) -> float:
    """Gradient averaging of data-parallel training (DDP all-reduce behind accelerator.backward, trainer.py:1280) over ONE flat
    buffer: a single NCCL all-reduce (NVLink / NVSwitch, NVLS when available) instead of per-bucket hooks.  Sums in place and
    returns the factor (1 / world) the optimizer kernel folds into its gradient scale."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1.0
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / dist.get_world_size(group)


def broadcast_flat_(flat, src: int = 0, group=None) -> None:
    """initial parameter broadcast of accelerator.prepare (trainer.py:324)"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
