"""Multi-GPU plumbing for the batch-sharded forward: every (batch, head) is an independent
attention problem, so ranks take disjoint batch slices and the forward has NO collective.  The only
communication the benchmark needs is the max-over-ranks reduction of the device-side timing and a barrier;
the reference's own helpers for the same job are source/utils/common.py:18-102 (init_ddp, reduce_dict)."""
from __future__ import annotations

from typing import Tuple

import torch


def batch_slice(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of the global batch owned by `rank`; remainders go to the lowest ranks (what
    DistributedSampler / train.py:110 amount to when the batch divides evenly)."""
    assert 0 <= rank < world and global_batch >= 0
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Whole-job time = slowest rank.  Works with NCCL (cuda tensor) and gloo (cpu tensor)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def job_throughput(units_per_rank: int, world: int, ms_max: float) -> float:
    """Aggregate units/s of a weak-scaled job: every rank processed `units_per_rank` in `ms_max` ms."""
    return units_per_rank * world / (ms_max * 1e-3)
