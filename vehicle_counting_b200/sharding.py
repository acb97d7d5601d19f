"""Multi-GPU layout of the hot path: frames are independent units (the reference already processes one frame per
iteration with no cross-frame state on this path, /root/reference/modules/__init__.py:54-57), so rank r owns
frames f % world == r, weights are replicated, no activation crosses GPUs, and the only collective is one
all-gather of int64 counters at the end of the stream (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence

import torch


def frames_of_rank(num_frames: int, rank: int, world: int) -> range:
    """Round-robin ownership: frame f belongs to rank f % world."""
    return range(rank, num_frames, world)


def frames_per_rank(num_frames: int, world: int) -> List[int]:
    return [len(frames_of_rank(num_frames, r, world)) for r in range(world)]


def gather_counters(counters: Sequence[int], device=None) -> List[List[int]]:
    """All-gather one int64 vector per rank; returns the per-rank vectors on every rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [list(int(c) for c in counters)]
    t = torch.tensor(list(counters), dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def max_over_ranks(values: Sequence[float], device=None) -> List[float]:
    """Element-wise MAX all-reduce (timings are reported as the slowest rank's)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
