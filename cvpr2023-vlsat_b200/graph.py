"""CUDA-graph replay of the forward pass.

The forward is ~150 kernel launches through the C ABI; at B200 speeds the Python / launch overhead of
issuing them one by one (tens of microseconds each) exceeds the GPU time of the small ones. Every launch
goes to ``torch.cuda.current_stream()``, allocations come from PyTorch's caching allocator and TMA
descriptors are plain kernel parameters, so the whole forward can be captured once per input shape and
replayed with a single launch. Inputs are copied into static buffers before each replay.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


class GraphedForward:
    """``GraphedForward(model)(obj_points, obj_2d_feats, edge_indices, descriptor, batch_ids)`` - same
    outputs as ``model(...)`` in eval mode; one CUDA graph per distinct input-shape signature."""

    def __init__(self, model: torch.nn.Module, istrain: bool = False, max_graphs: int = 16):
        self.model, self.istrain, self.max_graphs = model, istrain, max_graphs
        self._graphs: Dict[Tuple, tuple] = {}
        self.kernels_per_replay = 0

    @staticmethod
    def _sig(args) -> Tuple:
        return tuple((tuple(a.shape), a.dtype, a.device) for a in args)

    def capture(self, *args):
        sig = self._sig(args)
        static_in = [a.clone() for a in args]
        with torch.no_grad():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                 # warm-up: lazy init, weight splits, allocator pool
                for _ in range(2):
                    self.model(*static_in, istrain=self.istrain)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            from . import ops
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(graph):
                static_out = self.model(*static_in, istrain=self.istrain)
            self.kernels_per_replay = ops.launch_count() - n0      # vlsat kernel nodes in the graph
        if len(self._graphs) >= self.max_graphs:
            self._graphs.pop(next(iter(self._graphs)))
        self._graphs[sig] = (graph, static_in, static_out)
        return self._graphs[sig]

    def __call__(self, *args):
        entry = self._graphs.get(self._sig(args))
        if entry is None:
            entry = self.capture(*args)
        graph, static_in, static_out = entry
        for dst, src in zip(static_in, args):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        graph.replay()
        return static_out
