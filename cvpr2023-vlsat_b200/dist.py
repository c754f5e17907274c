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


class GradientAllReducer:
    """Data-parallel training: scenes shard by batch, the only collective is one all-reduce (mean) of the
    gradients per step - NCCL over NVLink on the GPU box, gloo in the CPU tests.

    Gradients are packed into a few flat buckets (one collective launch each, sized for launch latency rather than link
    count), reduced asynchronously and unpacked; ``attach(optimizer)`` runs this from an optimizer pre-step hook, so the
    reference's own ``loss.backward(); optimizer.step()`` (SGFN_MMG/model.py:483-488) needs no change.
    Parameters without a gradient in this step (``triplet_projector_3d`` is never used by the forward) are skipped; the
    set is the same on every rank because it depends on the model graph only."""

    def __init__(self, params, bucket_bytes: int = 64 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.bucket_bytes = bucket_bytes
        self.last_bytes = 0

    def _buckets(self, grads):
        cur, size = [], 0
        for g in grads:
            nbytes = g.numel() * g.element_size()
            if cur and size + nbytes > self.bucket_bytes:
                yield cur
                cur, size = [], 0
            cur.append(g)
            size += nbytes
        if cur:
            yield cur

    def allreduce(self) -> None:
        _, size = world()
        grads = [p.grad for p in self.params if p.grad is not None]
        self.last_bytes = sum(g.numel() * g.element_size() for g in grads)
        if size == 1 or not grads:
            return
        pending = []
        for bucket in self._buckets(grads):
            flat = torch.cat([g.reshape(-1) for g in bucket])
            pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, bucket))
        for work, flat, bucket in pending:
            work.wait()
            flat.div_(size)
            off = 0
            for g in bucket:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()

    def attach(self, optimizer: torch.optim.Optimizer):
        """Average gradients across ranks right before every ``optimizer.step()``."""
        return optimizer.register_step_pre_hook(lambda opt, args, kwargs: self.allreduce())
