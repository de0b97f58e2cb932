"""Multi-GPU plumbing of the path: utterances are independent, so ranks take contiguous shards of
the utterance list, weights are replicated and there is NO collective on the data path.  The only
communication is the timing protocol of the bench: a barrier on both sides of the timed region and a MAX
reduction of the per-rank device time (``torch.distributed``: NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int):
    """Contiguous shard [start, start + count) of ``n_items`` utterances for ``rank`` of ``world``; the first
    ``n_items % world`` ranks take one extra item."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def barrier(device=None):
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(value: float, device="cpu") -> float:
    """The job's time for a timed region = the slowest rank's device time."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
