"""Derived-weight cache: packed / concatenated / folded copies of parameters that the kernels consume.

A derived tensor is rebuilt whenever one of its source parameters changes storage (``.to(device)``) or
is written in place (``load_state_dict``, ``optimizer.step`` bump ``Tensor._version``).
"""
from __future__ import annotations

from typing import Callable, Dict, Sequence, Tuple

import torch


class DerivedCache:
    def __init__(self) -> None:
        self._store: Dict[str, Tuple[tuple, object]] = {}

    def get(self, key: str, sources: Sequence[torch.Tensor], build: Callable[[], object]):
        sig = tuple((t.data_ptr(), t._version, t.device) for t in sources)
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = build()
            if hit is not None:
                # refresh IN PLACE when shapes allow: stable addresses keep downstream caches (the tf32
                # splits keyed on data_ptr + version) bounded during training.
                val = _copy_into(hit[1], val)
        self._store[key] = (sig, val)
        return val

    def clear(self) -> None:
        self._store.clear()


def _copy_into(old, new):
    if isinstance(old, torch.Tensor) and isinstance(new, torch.Tensor):
        if old.shape == new.shape and old.device == new.device and old.dtype == new.dtype:
            old.copy_(new)
            return old
        return new
    if isinstance(old, tuple) and isinstance(new, tuple) and len(old) == len(new):
        return tuple(_copy_into(o, n) for o, n in zip(old, new))
    return new


def require_inference(module: torch.nn.Module, what: str) -> None:
    """Guard of the fused inference-only entry points (``attend_scenes``, ``fused`` ...): training mode and autograd
    go through the modules' ``forward`` (-> train_path.py), never through these. There is no PyTorch fallback."""
    if module.training:
        raise NotImplementedError(
            f"{what}: this is a fused inference entry point; in training mode call the module's forward "
            "(differentiable vlsat_b200 path, train_path.py).")
