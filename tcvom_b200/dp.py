"""Data-parallel plumbing for the frame-window path: windows are independent units (SURVEY.md
section 8e), so ranks only share (a) the partition of the work list and (b) the timing reduction.
There is no data-path collective.  Mirrors the reference's contiguous per-process ranges
(pred_test.py:124-137) and its distributed helpers (utils/utils.py:41-59)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) of `n_items` work units owned by `rank` (pred_test.py:125-131:
    pproc = len // world + 1; range(i*pproc, min((i+1)*pproc, len)))."""
    if world <= 1:
        return 0, n_items
    pproc = n_items // world + 1
    return min(rank * pproc, n_items), min((rank + 1) * pproc, n_items)


def shard_round_robin(n_items: int, rank: int, world: int) -> List[int]:
    """window i -> rank i mod world (streaming inputs: keeps per-rank queues balanced)."""
    return list(range(rank, n_items, max(world, 1)))


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (device time of the timed region) over all ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def job_throughput(units_this_rank: int, ms_this_rank: float, device=None) -> float:
    """Whole-job units/s = units processed by all ranks / max-over-ranks time."""
    total = sum_over_ranks(float(units_this_rank), device)
    ms = max_over_ranks(ms_this_rank, device)
    return total / (ms / 1e3)


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
