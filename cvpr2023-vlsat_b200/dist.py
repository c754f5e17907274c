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
    gradients per step - NCCL over NVLink on the GPU box, gloo in the CPU tests. The reference is single-process; the
    ordering follows its ``backward()`` (SGFN_MMG/model.py:483-488): between ``loss.backward()`` and ``optimizer.step()``.

    One persistent flat fp32 buffer holds every gradient of a step (16-byte aligned slots, allocated once per set of
    gradient addresses). Per step: ONE multi-tensor launch packs the gradients into it with the 1 / world_size of the mean
    folded in (``vlsat_pack_scale``), ONE all-reduce (sum) runs on the whole buffer, and every ``p.grad`` is re-pointed at
    its slot - no ``torch.cat``, no divide pass, no copy back; the optimiser reads the averaged gradients where NCCL left
    them. Parameters without a gradient in this step (``triplet_projector_3d`` is never used by the forward) are skipped;
    the set is the same on every rank because it depends on the model graph only. ``attach(optimizer)`` runs this from an
    optimizer pre-step hook, so the reference's own ``loss.backward(); optimizer.step()`` needs no change."""

    def __init__(self, params, chunk_elems: int = 16384):
        self.params = [p for p in params if p.requires_grad]
        self.chunk_elems = chunk_elems
        self.last_bytes = 0
        self._plan = None

    def _build(self, ps, grads):
        """(flat, views, table, chunk lists): the flat buffer and its views depend on the gradient SHAPES only (allocated once);
        the pointer table is rebuilt when the gradients live somewhere else (eager steps allocate new gradients; a replayed
        CUDA graph keeps them in place, so the steady state rebuilds nothing)."""
        shapes = tuple(tuple(g.shape) for g in grads)
        if self._plan is None or self._plan["shapes"] != shapes or self._plan["flat"].device != grads[0].device:
            offs, total = [], 0
            for g in grads:
                offs.append(total)
                total += (g.numel() + 3) // 4 * 4                  # 16-byte aligned slots
            flat = torch.zeros((total,), device=grads[0].device, dtype=torch.float32)
            views = [flat[o:o + g.numel()].view(g.shape) for o, g in zip(offs, grads)]
            cts, cis = [], []
            for i, g in enumerate(grads):
                nch = (g.numel() + self.chunk_elems - 1) // self.chunk_elems
                cts += [i] * nch
                cis += list(range(nch))
            dev = grads[0].device
            self._plan = dict(shapes=shapes, flat=flat, views=views, ptrs=None, table=None, n_chunks=len(cts),
                              ct=torch.tensor(cts, dtype=torch.int32).to(dev) if dev.type == "cuda" else None,
                              ci=torch.tensor(cis, dtype=torch.int32).to(dev) if dev.type == "cuda" else None)
        plan = self._plan
        ptrs = tuple(g.data_ptr() for g in grads)
        if plan["flat"].is_cuda and plan["ptrs"] != ptrs:
            from ._lib import CopyTensor
            arr = (CopyTensor * len(grads))()
            for i, (g, v) in enumerate(zip(grads, plan["views"])):
                arr[i].dst, arr[i].src, arr[i].n = v.data_ptr(), g.data_ptr(), g.numel()
            plan["table"] = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(plan["flat"].device)
            plan["ptrs"] = ptrs
        return plan

    def _pack_and_reduce(self, plan, grads, size: int) -> None:
        flat = plan["flat"]
        if flat.is_cuda:
            from . import _lib, ops
            _lib.check(ops._call("vlsat_pack_scale", plan["table"].data_ptr(), plan["ct"].data_ptr(), plan["ci"].data_ptr(), plan["n_chunks"],
                                 self.chunk_elems, 1.0 / size, ops._stream()), "vlsat_pack_scale")
        else:                                                      # CPU tensors: the gloo tests of the host-side logic
            for v, g in zip(plan["views"], grads):
                v.copy_(g).div_(size)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)

    def allreduce(self) -> None:
        _, size = world()
        ps = [p for p in self.params if p.grad is not None]
        grads = [p.grad for p in ps]
        self.last_bytes = sum(g.numel() * g.element_size() for g in grads)
        if size == 1 or not grads:
            return
        if self._plan is not None and len(grads) == len(self._plan["views"]) and all(g.data_ptr() == v.data_ptr() for g, v in zip(grads, self._plan["views"])):
            return                                                 # already this step's averaged gradients (called twice)
        for i, g in enumerate(grads):
            if g.dtype != torch.float32 or not g.is_contiguous():
                grads[i] = ps[i].grad = g.float().contiguous()
        plan = self._build(ps, grads)
        self._pack_and_reduce(plan, grads, size)
        for p, v in zip(ps, plan["views"]):
            p.grad = v

    def replay(self) -> None:
        """Pack + all-reduce once more from the gradient buffers of the last ``allreduce()`` call, without touching ``.grad``:
        the collective alone, for timing (bench.py's grad_allreduce_ms)."""
        _, size = world()
        if size > 1 and self._plan is not None and self._plan["flat"].is_cuda and self._plan["table"] is not None:
            self._pack_and_reduce(self._plan, None, size)

    def attach(self, optimizer: torch.optim.Optimizer):
        """Average gradients across ranks right before every ``optimizer.step()``."""
        return optimizer.register_step_pre_hook(lambda opt, args, kwargs: self.allreduce())
