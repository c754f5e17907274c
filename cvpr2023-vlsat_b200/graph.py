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
        self._flag_host = None          # pinned copy of the sticky sorted-batch_ids flag, polled (never waited for)
        self._sticky = None

    def _weights_version(self) -> int:
        """Sum of the version counters of every parameter and buffer: the derived weights a captured graph reads (packed /
        folded / split copies) were built from these versions; a different sum means an optimiser step or a
        ``load_state_dict`` happened since the capture and the graphs must be re-captured."""
        v = 0
        for p in self.model.parameters():
            v += p._version
        for b in self.model.buffers():
            v += b._version
        return v

    def _check_previous_flag(self) -> None:
        """Never blocks: reads the pinned copy of a device-side sticky counter that every replay adds its flag to. The copy
        of the replay that set it may still be in flight, so the error surfaces at the next call or the one after - a host
        sync here would serialise the replays with the host (measured: 2.63 -> 2.92 ms per step)."""
        if self._flag_host is not None and int(self._flag_host.item()) != 0:
            self._flag_host.zero_()
            self._sticky.zero_()
            raise RuntimeError("vlsat_b200: the batch_ids of a previous replay were not non-decreasing scene ids "
                               "(src/dataset/DataLoader.py:153-176); the outputs of that replay are invalid")

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
            from .attention import _last_err_flag
            if self._flag_host is None:
                self._flag_host = torch.zeros((1,), dtype=torch.int32).pin_memory()
                self._sticky = torch.zeros((1,), dtype=torch.int32, device=static_in[0].device)
            with torch.cuda.graph(graph):
                static_out = self.model(*static_in, istrain=self.istrain)
                flag = _last_err_flag.get(static_in[0].device)      # written by the scene-range kernel inside the graph
                if flag is not None:
                    # input validation without a host sync: a sticky device counter and its 4-byte copy to pinned memory
                    # are nodes of the graph; __call__ polls the pinned value
                    self._sticky.add_(flag)
                    self._flag_host.copy_(self._sticky, non_blocking=True)
            self.kernels_per_replay = ops.launch_count() - n0      # vlsat kernel nodes in the graph
        if len(self._graphs) >= self.max_graphs:
            self._graphs.pop(next(iter(self._graphs)))
        self._graphs[sig] = (graph, static_in, static_out, self._weights_version(), flag)
        return self._graphs[sig]

    def __call__(self, *args):
        self._check_previous_flag()
        entry = self._graphs.get(self._sig(args))
        if entry is not None and entry[3] != self._weights_version():
            self._graphs.clear()                                 # weights changed since the capture: the derived copies are stale
            entry = None
        if entry is None:
            entry = self.capture(*args)
        graph, static_in, static_out, _, flag = entry
        for dst, src in zip(static_in, args):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        graph.replay()
        return static_out


class GraphedTrainStep:
    """``GraphedTrainStep(model, loss_fn)(*forward_args)`` -> ``(loss, outputs)``: ``model(..., istrain=True)``,
    ``loss_fn(outputs)`` and ``loss.backward()`` captured as ONE CUDA graph per input signature and replayed with a single
    launch (a training step is ~1800 launches through the C ABI; issued one by one the host is the bottleneck).

    After the call every trainable parameter's ``.grad`` holds this step's gradient (static buffers owned by the graph,
    re-attached after each replay, so ``optimizer.zero_grad(set_to_none=True)`` between steps is fine); the optimizer
    step stays the caller's. Dropout masks differ between replays: the kernels mix a device-side step counter, bumped
    inside the graph, into their seed. The signature includes the scene composition statistics that size buffers
    (number of same-scene pairs, largest scene): one small host sync per call unless ``scene_stats`` is passed."""

    def __init__(self, model: torch.nn.Module, loss_fn, max_graphs: int = 8):
        self.model, self.loss_fn, self.max_graphs = model, loss_fn, max_graphs
        self._graphs: Dict[Tuple, tuple] = {}
        self.kernels_per_replay = 0
        self._step = None

    def _run(self, static_in):
        from . import ops
        with ops.zero_arena(self, static_in[0].device):      # the backward's small accumulation targets: one fill per step
            outs = self.model(*static_in, istrain=True)
            loss = self.loss_fn(outs)
            loss.backward()
        return loss, outs

    def capture(self, args, stats):
        from . import autograd as A
        from . import ops
        from . import train_path as T
        dev = args[0].device
        if self._step is None:
            self._step = torch.zeros((1,), device=dev, dtype=torch.int64)
        static_in = [a.clone() for a in args]
        old_hint, old_step = T._scene_hint, A.DropoutState.device_step
        T._scene_hint, A.DropoutState.device_step = stats, self._step
        try:
            # The two warm-up passes are real training-mode forwards: without care the first step of every new input
            # signature would update the BatchNorm running statistics three times (and count three batches) and consume
            # two steps of dropout stream - the reference does each once per step (advisor finding, round 1). Snapshot the
            # buffers and the dropout offset, restore them after the warm-up.
            buffers = [(b, b.detach().clone()) for b in self.model.buffers()]
            drop_state = (A.DropoutState.seed, A.DropoutState.offset)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                 # warm-up: lazy init, allocator pool
                for _ in range(2):
                    self.model.zero_grad(set_to_none=True)
                    self._run(static_in)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.no_grad():
                for b, saved in buffers:
                    b.copy_(saved)
            A.DropoutState.seed, A.DropoutState.offset = drop_state
            self.model.zero_grad(set_to_none=True)
            ops._weight_splits.clear()                    # every weight split must be a kernel node of the graph
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(graph):
                self._step.add_(1)
                loss, outs = self._run(static_in)
            self.kernels_per_replay = ops.launch_count() - n0
        finally:
            T._scene_hint, A.DropoutState.device_step = old_hint, old_step
        grads = [(p, p.grad) for p in self.model.parameters() if p.grad is not None]
        if len(self._graphs) >= self.max_graphs:
            self._graphs.pop(next(iter(self._graphs)))
        entry = (graph, static_in, loss, outs, grads)
        self._graphs[(GraphedForward._sig(args), stats)] = entry
        return entry

    def __call__(self, *args, scene_stats=None):
        from . import train_path as T
        if not self.model.training and not torch.is_grad_enabled():
            raise RuntimeError("GraphedTrainStep needs autograd enabled")
        stats = tuple(scene_stats) if scene_stats is not None else T.scene_stats(args[4])
        entry = self._graphs.get((GraphedForward._sig(args), stats))
        if entry is None:
            entry = self.capture(args, stats)
        graph, static_in, loss, outs, grads = entry
        for dst, src in zip(static_in, args):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        graph.replay()
        for p, g in grads:
            p.grad = g
        return loss, outs


class StreamedInference:
    """Host-to-host inference loop around ``GraphedForward``: every batch comes from (pinned) host memory and every result
    goes back to pinned host memory, with the copies of neighbouring batches overlapped with the compute of the current one.

        pipe = StreamedInference(model)
        for batch in loader:                       # tuples of the five forward arguments, host tensors
            done = pipe.submit(batch)              # returns the PREVIOUS batch's outputs (host tensors) or None
        last = pipe.drain()

    This is the serving-side counterpart of ``MMGNet.validation`` (src/model/model.py:201-215: ``.cuda()`` of the collated
    batch, forward, ``.detach().cpu()`` of the logits). Three streams: H2D into one of two staging sets on the copy-in
    stream, a device-to-device hop into the graph's static inputs + the replay + a hop of the outputs into one of two
    staging sets on the compute stream, D2H on the copy-out stream. The result buffers returned for batch i are reused for
    batch i + 2: consume (or copy) them before submitting two more."""

    def __init__(self, model: torch.nn.Module):
        self.graphed = GraphedForward(model)
        self.device = next(model.parameters()).device
        self.h2d, self.d2h = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self._in = [None, None]        # device staging sets of the inputs
        self._out_dev = [None, None]   # device staging sets of the outputs
        self._out_host = [None, None]
        self._in_free = [None, None]   # events: the compute stream has consumed staging set k
        self._out_done = [None, None]  # events: D2H of staging set k has finished
        self._i = 0
        self._pending = None

    def submit(self, host_args):
        k = self._i & 1
        comp = torch.cuda.current_stream()
        with torch.cuda.stream(self.h2d):
            if self._in_free[k] is not None:
                self.h2d.wait_event(self._in_free[k])
            if self._in[k] is None or any(a.shape != b.shape for a, b in zip(self._in[k], host_args)):
                self._in[k] = [torch.empty(a.shape, dtype=a.dtype, device=self.device) for a in host_args]
            for dst, src in zip(self._in[k], host_args):
                dst.copy_(src, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.h2d)
        comp.wait_event(ready)
        outs = self.graphed(*self._in[k])                       # device-to-device hop into the static inputs + replay
        self._in_free[k] = torch.cuda.Event()
        self._in_free[k].record(comp)
        if self._out_dev[k] is None or any(a.shape != b.shape for a, b in zip(self._out_dev[k], outs)):
            self._out_dev[k] = [torch.empty_like(o) for o in outs]
            self._out_host[k] = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
        if self._out_done[k] is not None:
            comp.wait_event(self._out_done[k])                  # the D2H that last read this staging set is done
        for dst, src in zip(self._out_dev[k], outs):
            dst.copy_(src, non_blocking=True)
        computed = torch.cuda.Event()
        computed.record(comp)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(computed)
            for dst, src in zip(self._out_host[k], self._out_dev[k]):
                dst.copy_(src, non_blocking=True)
            self._out_done[k] = torch.cuda.Event()
            self._out_done[k].record(self.d2h)
        prev, self._pending = self._pending, k
        self._i += 1
        if prev is None:
            return None
        self._out_done[prev].synchronize()                      # the caller consumes batch i - 1 while batch i computes
        return self._out_host[prev]

    def drain(self):
        if self._pending is None:
            return None
        k, self._pending = self._pending, None
        self._out_done[k].synchronize()
        return self._out_host[k]
