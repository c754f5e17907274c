"""Scene-parallel helpers: one process per GPU, scenes sharded by batch, no data-path
collective in the forward; the only reductions are bookkeeping (max-over-ranks timing, throughput sums)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .synth import SceneBatch, shard_scenes


def world() -> tuple:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_for_rank(batch: SceneBatch) -> SceneBatch:
    """This rank's scenes {r, r+W, ...} of a global batch, re-based to local node ids."""
    rank, size = world()
    return batch if size == 1 else shard_scenes(batch, rank, size)


def max_over_ranks(value: float, device="cpu") -> float:
    """Device/elapsed time of a step is the max over ranks (never wall clock of one rank)."""
    _, size = world()
    if size == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    _, size = world()
    if size == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
