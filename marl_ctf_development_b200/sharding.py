"""Env sharding over ranks: one process per GPU, no collective inside the step.

Rank r owns the global env ids [r*B, (r+1)*B).  The counter RNG is keyed by the *global* id, so results do
not depend on the number of GPUs.  The only exchange is the sum of the episode-statistics counters.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def env_id_base(envs_per_rank: int, rank: int | None = None) -> int:
    r = rank_world()[0] if rank is None else rank
    return r * int(envs_per_rank)


def shard_bounds(total_envs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split of a fixed total (strong-scaling use): [lo, hi) of rank."""
    lo = total_envs * rank // world
    hi = total_envs * (rank + 1) // world
    return lo, hi


def all_reduce_stats(counters: torch.Tensor) -> torch.Tensor:
    """SUM over ranks of the int64 [13, N] counters (NCCL on GPUs, gloo on CPU); identity when not distributed."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters
