"""Multi-GPU plumbing (one process per GPU, torch.distributed).

Decoding shards utterances across ranks with no data-path collective.  Segmental K-means
has exactly one exchange per iteration: the packed float64 sufficient statistics and the
int32 transition counts are summed over ranks (NCCL all-reduce over NVLink on GPUs, gloo in
the CPU tests), after which every rank runs the identical M-step.  The
reference has no counterpart: its only parallelism is a process pool (hidden_markov_model.py:300-305).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple, TypeVar

T = TypeVar("T")


def _dist():
    import torch.distributed as dist
    return dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(items: Sequence[T]) -> List[T]:
    """Round-robin shard of a list for this rank (the whole list when not distributed)."""
    rank, n = world()
    return list(items) if n == 1 else list(items[rank::n])


def shard_by_frames(lengths: Sequence[int], n_ranks: int) -> List[List[int]]:
    """Balanced partition of utterance indices by total frames (longest first, greedy)."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    loads = [0] * n_ranks
    parts: List[List[int]] = [[] for _ in range(n_ranks)]
    for i in order:
        r = loads.index(min(loads))
        parts[r].append(i)
        loads[r] += lengths[i]
    return [sorted(p) for p in parts]


def allreduce_stats(stats, counts):
    """Sum the statistics tensors over all ranks: ONE collective per iteration on a single
    packed float64 buffer (counts ride along as float64: exact below 2**53)."""
    rank, n = world()
    if n == 1:
        return stats, counts
    import torch

    dist = _dist()
    packed = torch.cat([stats.reshape(-1), counts.reshape(-1).to(torch.float64)])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    ns = stats.numel()
    return packed[:ns].reshape(stats.shape), packed[ns:].round().to(counts.dtype).reshape(counts.shape)
