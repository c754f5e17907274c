#!/usr/bin/env python
"""bench.py - scenes/s of the VL-SAT hot path (Mmgnet.forward, and the training step around it) on synthetic scenes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference|reference-gpu] [--workload cfg2]
                  [--metric fwd|train_step]

Workload (BASELINE.json configs[1]): 16 synthetic scenes x 40 objects x 256 points, 600 edges per scene
(sum_N 640, sum_E 9600), mmgnet.json model (2 layers, 8 heads, 512/512/256), fp32, eval-mode forward.
Under torchrun every rank runs its own 16-scene batch (scenes shard by batch: weak scaling, no
data-path collective in the forward; the training step all-reduces its gradients over NCCL). Rank 0 prints ONE JSON line.

  value    : scenes/s with the batch resident in HBM, CUDA events per step, L2 flushed between steps
  e2e      : scenes/s through the public module API from pinned HOST buffers (H2D of the batch and D2H
             of the results inside the timed region, wall clock between device syncs)
  roofline : dominant KERNEL of the step by device time (CUDA events around each C-ABI launch in a separate pass; the
             projection entry point is split into its two tile kernels, 128x128 and 128x64, as the ncu launch list shows
             them; linear_all_* keeps the entry point as a whole); the GAT scatter kernel BASELINE.json names and the other
             tensor kernels are carried as scalars (gat_frac, flash_frac, ...)
  config   : besides the workload, the secondary measurements as SCALARS (the driver's records keep `config` and
             `roofline` whole): fwd_bwd_ms, train_step_ms, grad_allreduce_ms, ref_gpu_ms / ref_gpu_speedup (the
             reference's PyTorch forward on the same B200 at 16 and 64 scenes - north_star's ">= 10x" denominator)
  cpu_baseline : the reference's algorithm on this box's host cores

--metric train_step makes the full training step (forward, the reference's losses, backward, NCCL gradient all-reduce at
N > 1, fused AdamW) the headline `value` / `e2e` instead of the forward; it is the leg with a collective.

--impl reference times the reference's own CPU implementation on the host cores: the UNMODIFIED reference modules staged
under baseline/_ref (oracle/stage_reference.py; kind "reference") when present, else the oracle port (kind "port").
--impl reference-gpu runs the same on cuda:0 (not part of the driver's contract; the default run reports it in `config`).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

import torch  # noqa: E402

UNIT = "scenes/s"
METRICS = {"fwd": "scenes_per_sec_fwd", "train_step": "scenes_per_sec_train_step"}

# stdout carries exactly ONE line (the JSON): everything else a library may print there (NCCL's version banner, ...) is
# sent to stderr by pointing fd 1 at fd 2 for the duration of the run
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions: NVML polled from a thread every 2 ms (an
    nvidia-smi child needs longer to start than a 60 ms timed region lasts - round 1 recorded 0 or 1 samples), with
    nvidia-smi as the fallback. Samples are kept only between ``start()`` and ``stop()``."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.sm, self.mx, self.reasons = index, [], 0, set()
        self._on, self._quit, self._thread, self._h, self._nv = False, False, None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.mx = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        except Exception:
            self._nv = None

    @staticmethod
    def _physical_index(local: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip()]
            if local < len(ids) and ids[local].strip().isdigit():
                return int(ids[local])
        return local

    def _poll(self):
        nv = self._nv
        while not self._quit:
            if self._on:
                try:
                    self.sm.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.002)

    def start(self):
        self._on = True

    def stop(self):
        self._on = False

    def close(self):
        self._quit = True

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            return float(out[0]), float(out[1])
        except Exception:
            return None, None

    def summary(self):
        if not self.sm:
            sm, mx = self._smi_once()
            return {"sm_mhz": sm, "sm_max_mhz": mx, "reasons": [], "samples": 0 if sm is None else 1, "source": "nvidia-smi after the run"}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "NVML polled every 2 ms inside the timed regions"}


def workload_kwargs(name: str):
    from vlsat_b200 import synth
    return dict(synth.CONFIGS[name])


def describe_objects(kw) -> str:
    o = kw["objects_per_scene"]
    return f"{o} obj" if isinstance(o, int) else f"{min(o)}..{max(o)} obj (mean {sum(o) / len(o):.1f}, 3RScan-shaped)"


def describe_edges(kw) -> str:
    return f"{kw['edges_per_scene']} edges/scene" if kw["edges_per_scene"] is not None else "fully connected"


def workload_text(name: str, kw, scenes: int, metric: str, dtype: str = "f32") -> str:
    what = "forward (eval)" if metric == "fwd" else "training step (forward, losses, backward, AdamW)"
    what = ("fp32 " if dtype == "f32" else "single-pass bf16 ") + what
    return (f"{name}: {scenes} scenes/GPU x {describe_objects(kw)} x {kw['points_per_object']} pts, {describe_edges(kw)}, "
            f"mmgnet.json model (L=2, H=8), {what}")


def build_model(device):
    import vlsat_b200 as V
    from vlsat_b200 import synth
    model = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)
    synth.load_seeded(model, 0)
    return model.to(device).eval()


def make_targets(batch, gen, device):
    """Synthetic supervision of one batch: object classes, multi-label relationships, unit-norm text embedding (stands for
    the CLIP text encoder of get_rel_emb, SGFN_MMG/model.py:221-255)."""
    n_, e_ = batch.obj_points.shape[0], batch.edge_indices.shape[1]
    text = torch.randn(e_, 512, generator=gen)
    return (torch.randint(0, 160, (n_,), generator=gen).to(device), (torch.rand(e_, 26, generator=gen) < 1.0 / 26).float().to(device),
            (text / text.norm(dim=-1, keepdim=True)).to(device))


# ------------------------------------------------------------------------------------------ reference arms
def reference_staged() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "src", "model", "SGFN_MMG", "model.py"))


class ReferenceRunner:
    """The reference's own implementation of the path on ``device``: the unmodified modules staged under baseline/_ref
    when present (kind "reference"), otherwise the oracle port (kind "port"). Test / measurement infrastructure only."""

    def __init__(self, device: str, train: bool = False):
        import vlsat_b200 as V
        from vlsat_b200 import synth
        self.device, self.train = torch.device(device), train
        ours = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)
        synth.load_seeded(ours, 0)
        self.kind = "port"
        self.net = None
        if reference_staged():
            try:
                os.environ["VLSAT_REFERENCE_ROOT"] = REF_DIR
                from oracle import ref_shims
                if self.device.type == "cpu":
                    torch.Tensor.cuda = lambda self_, *a, **k: self_          # network_MMG.py:185-186 hard-code .cuda()
                net, _ = ref_shims.build_reference_mmgnet(0)
                net.load_state_dict(ours.state_dict(), strict=True)
                self.net = net.to(self.device)
                self.net.train() if train else self.net.eval()
                self.kind = "reference"
            except Exception as exc:                                         # report, then use the port
                print(f"[bench] staged reference unusable ({type(exc).__name__}: {exc}); using the oracle port", file=sys.stderr)
                self.net = None
        if self.net is None:
            self.sd = {k: v.detach().to(self.device).requires_grad_(train and v.is_floating_point() and not k.startswith("clip_adapter"))
                       for k, v in ours.state_dict().items()}
            self.opt = torch.optim.AdamW([v for v in self.sd.values() if v.requires_grad], lr=1e-4) if train else None

    def forward(self, batch):
        from oracle import vlsat_oracle as O
        with torch.no_grad():
            if self.net is not None:
                return self.net(*batch.forward_args(), istrain=False)
            return O.mmgnet_forward(self.sd, *batch.forward_args(), istrain=False)

    def train_step(self, batch, targets):
        """forward(istrain=True) -> the six loss terms of process_train (oracle restatement of SGFN_MMG/model.py:343-412:
        get_rel_emb needs CLIP weights) -> backward -> the reference's own AdamW (13 groups) + cosine scheduler."""
        from oracle import vlsat_oracle as O
        if self.net is not None:
            outs = self.net(*batch.forward_args(), istrain=True)
            loss, _ = O.train_losses(outs, *targets)
            self.net.backward(loss)                                          # SGFN_MMG/model.py:483-488
        else:
            outs = O.mmgnet_forward(self.sd, *batch.forward_args(), istrain=True)
            loss, _ = O.train_losses(outs, *targets)
            self.opt.zero_grad(set_to_none=True)
            loss.backward()
            self.opt.step()
        return loss


def time_reference(workload: str, metric: str, steps: int, warmup: int, budget_s: float, device: str = "cpu"):
    """Reference arm on ``device``. Returns dict(value, cores, sample, ms_per_step, scenes, kind)."""
    from vlsat_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    train = metric == "train_step"
    ref = ReferenceRunner(device, train=train)
    kw = workload_kwargs(workload)
    scenes = kw["num_scenes"]
    gen = torch.Generator().manual_seed(77)

    def make(kw_):
        b = synth.make_batch(seed=100, **kw_)
        t = make_targets(b, gen, ref.device) if train else None
        return b.to(ref.device), t
    # the reference materialises [8, sum_E, sum_E] fp32 score tensors (attention.py:55-66), about three alive at once (more
    # with autograd): bound the sample BEFORE the first run so that it fits the host (8 GB) / the GPU (100 GB)
    per_scene_edges = kw["edges_per_scene"]
    if per_scene_edges is None:
        o = kw["objects_per_scene"]
        per_scene_edges = max(o) * (max(o) - 1) if not isinstance(o, int) else o * (o - 1)
    limit = (8e9 if ref.device.type == "cpu" else 100e9) / (96.0 * (3.0 if train else 1.0))
    while scenes > 1 and (scenes * per_scene_edges) ** 2 > limit:
        scenes = max(1, scenes // 2)
    if scenes != kw["num_scenes"]:
        kw["num_scenes"] = scenes
        if not isinstance(kw["objects_per_scene"], int):
            kw["objects_per_scene"] = list(kw["objects_per_scene"])[:scenes]
    batch, tg = make(kw)
    run = (lambda: ref.train_step(batch, tg)) if train else (lambda: ref.forward(batch))
    sync = torch.cuda.synchronize if ref.device.type == "cuda" else (lambda: None)
    t0 = time.perf_counter()
    run(); sync()
    first = time.perf_counter() - t0
    total_steps = steps + max(warmup - 1, 0)
    if first * total_steps > budget_s and scenes > 4:
        # cross_attn_rel is O(sum_E^2): shrink the sample to a 4-scene batch of the same scene shape
        scenes = 4
        kw["num_scenes"] = 4
        if not isinstance(kw["objects_per_scene"], int):
            kw["objects_per_scene"] = list(kw["objects_per_scene"])[:4]
        batch, tg = make(kw)
        run(); sync()
    for _ in range(max(warmup - 1, 0)):
        run()
    sync()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        run(); sync()
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    where = f"torch CPU fp32 with {cores} threads" if ref.device.type == "cpu" else "stock PyTorch (cuBLAS / cuDNN / ATen) on cuda:0"
    what = "training steps" if train else "forwards"
    src = "unmodified reference modules (baseline/_ref, dependency stand-ins of oracle/ref_shims.py)" if ref.kind == "reference" else "oracle port of the reference"
    return dict(value=scenes / sec, cores=cores, ms_per_step=sec * 1e3, scenes=scenes, kind=ref.kind,
                sample=f"{scenes}-scene batch of the {workload} scene shape, {steps} timed {what} after {warmup} warm-up, {src}, {where}")


def reference_gpu_yardstick(model, graphed, dev, scene_counts=(16, 64)):
    """The reference's PyTorch forward on the SAME B200 next to ours, at the benchmark batch and at north_star's 64 scenes.
    Returns scalars for `config`. Any failure (reference absent and port OOM, ...) is reported as a string, never fatal."""
    from vlsat_b200 import synth
    out = {}
    try:
        ref = ReferenceRunner(str(dev), train=False)
        out["ref_gpu_kind"] = ref.kind
        for scenes in scene_counts:
            batch = synth.make_config_batch("cfg2", seed=1, num_scenes=scenes).to(dev)

            def med(fn, warm, iters):
                for _ in range(warm):
                    fn()
                torch.cuda.synchronize()
                ts = []
                for _ in range(iters):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fn(); b.record()
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b))
                return statistics.median(ts)
            ours = med(lambda: graphed(*batch.forward_args()), 3, 10)
            try:
                theirs = med(lambda: ref.forward(batch), 1, 3)
            except torch.OutOfMemoryError:
                torch.cuda.empty_cache()
                out[f"ref_gpu_ms_{scenes}"] = "reference out of memory"
                continue
            out[f"ours_ms_{scenes}"] = round(ours, 3)
            out[f"ref_gpu_ms_{scenes}"] = round(theirs, 2)
            out[f"ref_gpu_speedup_{scenes}"] = round(theirs / ours, 2)
            torch.cuda.empty_cache()
        del ref
        torch.cuda.empty_cache()
    except Exception as exc:
        out["ref_gpu_error"] = f"{type(exc).__name__}: {exc}"[:200]
    return out


# ---------------------------------------------------------------------------------------------- main arm
def run_b200(args):
    import torch.distributed as dist
    import vlsat_b200 as V
    from vlsat_b200 import autograd as A
    from vlsat_b200 import dist as vd
    from vlsat_b200 import ops, synth
    from vlsat_b200 import train_glue as G
    from vlsat_b200 import train_path as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    kw = workload_kwargs(args.workload)
    scenes = kw["num_scenes"]
    if args.dtype == "bf16":
        ops.set_precision("bf16")
    model = build_model(dev)
    n_batches = 4
    host = [synth.make_batch(seed=1 + rank * 100 + i, **kw).pin() for i in range(n_batches)]
    resident = [b.to(dev) for b in host]
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)     # > 126 MB L2
    clk = ClockSampler(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_device(fn, n, warm):
        """n steps of fn(i) after `warm` untimed ones: CUDA events per step, L2 flushed between steps, barrier + sync on
        both sides, max over ranks. Returns total ms."""
        for i in range(warm):
            fn(i)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        clk.start()
        for i in range(n):
            flush.zero_()
            ev[i][0].record()
            fn(i)
            ev[i][1].record()
        barrier()
        clk.stop()
        return max_ranks(sum(a.elapsed_time(b) for a, b in ev))

    # ================= forward =================
    def step_eager(b):
        with torch.no_grad():
            return model(*b.forward_args(), istrain=False)
    # One CUDA graph per input-shape signature: the ~150 C-ABI launches of a forward are replayed with a
    # single launch; inputs are copied into the graph's static buffers first (vlsat_b200/graph.py).
    graphed = V.GraphedForward(model)
    step = (lambda b: graphed(*b.forward_args())) if not args.eager else step_eager
    launches0 = ops.launch_count()
    fwd_ms_total = timed_device(lambda i: step(resident[i % n_batches]), args.steps, args.warmup)
    fwd_launches = (ops.launch_count() - launches0) if args.eager else graphed.kernels_per_replay * args.steps
    fwd_ms = fwd_ms_total / args.steps
    fwd_value = world * scenes * args.steps / (fwd_ms_total * 1e-3)

    # ---- forward end to end through the public API from pinned host buffers: every step's inputs are copied host -> device
    # and its four logit tensors device -> host inside the timed region. StreamedInference overlaps the copies of
    # neighbouring steps with the compute of the current one (what a serving / validation loop does); --eager and
    # VLSAT_E2E=serial keep the one-stream, sync-per-step loop of round 1.
    d2h_bytes = 0
    serial = args.eager or os.environ.get("VLSAT_E2E", "") == "serial"
    if serial:
        out_host = None
        for i in range(2):                       # warm the copy path
            outs = step(host[i % n_batches].to(dev, non_blocking=True))
            if out_host is None:
                out_host = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
                d2h_bytes = sum(o.numel() * o.element_size() for o in outs)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            outs = step(host[i % n_batches].to(dev, non_blocking=True))
            for o, h in zip(outs, out_host):
                h.copy_(o, non_blocking=True)
            torch.cuda.synchronize()             # the caller consumes the logits of this step
        barrier()
    else:
        pipe = V.StreamedInference(model)
        checksum = 0.0
        for i in range(3):                       # warm the copy path and both staging sets
            pipe.submit(host[i % n_batches].forward_args())
        d2h_bytes = sum(o.numel() * o.element_size() for o in pipe.drain())
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            done = pipe.submit(host[i % n_batches].forward_args())
            if done is not None:
                checksum += float(done[2][0, 0])  # the caller reads the previous step's logits (host memory)
        checksum += float(pipe.drain()[2][0, 0])
        barrier()
    fwd_e2e = world * scenes * args.steps / max_ranks(time.perf_counter() - t0)
    fwd_e2e_line = {"value": round(fwd_e2e, 2), "unit": UNIT, "h2d_bytes_per_step": host[0].nbytes(), "d2h_bytes_per_step": d2h_bytes}

    # ---- the reference's PyTorch forward on this GPU (north_star's ">= 10x" denominator); rank 0, N = 1 only
    yard = reference_gpu_yardstick(model, graphed, dev) if (rank == 0 and world == 1 and not args.no_ref_gpu and args.workload == "cfg2"
                                                            and args.dtype == "f32") else {}

    # ---- N2: device side of the loader's per-object loop at this workload's object / point counts
    prep = {}
    if rank == 0 and world == 1:
        try:
            from vlsat_b200 import data_prep
            n_obj, n_pts = resident[0].obj_points.shape[0], resident[0].obj_points.shape[2]
            gen = torch.Generator(device=dev).manual_seed(9)
            per = 2048                                              # points of an instance in the scan cloud
            cloud = torch.randn(n_obj * per, 3, device=dev, generator=gen)
            choice = (torch.randint(0, per, (n_obj, n_pts), device=dev, generator=gen) + torch.arange(n_obj, device=dev).view(-1, 1) * per)
            for _ in range(3):
                data_prep.prepare_objects(cloud, choice)
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(20):
                data_prep.prepare_objects(cloud, choice)
            b_.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b_) / 20 * 1e3
            prep = {"object_prep_us": round(us, 2), "object_prep_gbs": round(n_obj * n_pts * 32.0 / (us * 1e-6) / 1e9, 1)}
            del cloud, choice
        except Exception as exc:
            prep = {"object_prep_error": f"{type(exc).__name__}: {exc}"[:160]}

    # ================= training: forward + backward, then the full step =================
    fwd_bwd = train_line = None
    train_scalars = {}
    train_e2e_line = None
    train_launches = 0
    n_train = min(args.steps, 10) if args.metric == "fwd" else args.steps
    stats = {}

    def stats_of(b):
        if id(b) not in stats:      # scene composition of a resident batch is fixed: no per-step host sync
            stats[id(b)] = T.scene_stats(b.batch_ids)
        return stats[id(b)]
    if not args.no_train:
        A.DropoutState.manual_seed(1234 + rank)
        # ---- full training step: the same forward + backward driven by the reference's losses
        # (process_train, SGFN_MMG/model.py:343-412) and followed by its optimiser step (AdamW, 13 groups, cosine schedule)
        try:
            tgen = torch.Generator().manual_seed(77 + rank)
            tgt, tgt_host = {}, {}

            def targets_of(b):
                if id(b) not in tgt:
                    tgt[id(b)] = make_targets(b, tgen, dev)
                return tgt[id(b)]
            for b in resident:            # synthetic supervision and scene statistics of every batch BEFORE any timed region:
                targets_of(b); stats_of(b)  # generating them lazily inside a timed loop cost ~25 ms of host time per batch and
                                            # was the "second training graph replays slower" oddity of round 1
            # a second instance with the same seeded weights: its parameters are really updated by the optimiser below,
            # the forward legs and the per-kernel pass keep measuring the original weights
            tmodel = build_model(dev).train()
            treducer = vd.GradientAllReducer(tmodel.parameters())
            opt = G.build_optimizer(tmodel, lr=1e-4, max_iteration=1000)
            ts = G.TrainStep(tmodel, opt, reducer=treducer, graphed=not args.eager)
            full_step = lambda b: ts.step(*b.forward_args(), *targets_of(b), scene_stats=stats_of(b))
            first_loss = float(full_step(resident[0])[0].item())
            if args.metric == "fwd":
                # forward + backward alone: the same captured step (train-mode forward, the reference's six loss terms,
                # backward) without the collective and the optimiser - one training graph serves both legs
                fb = lambda i: ts.forward_backward(*resident[i % n_batches].forward_args(), *targets_of(resident[i % n_batches]),
                                                   scene_stats=stats_of(resident[i % n_batches]))
                fb_ms = timed_device(fb, n_train, 2)
                fwd_bwd = {"value": round(world * scenes * n_train / (fb_ms * 1e-3), 2), "unit": UNIT, "ms_per_step": round(fb_ms / n_train, 3),
                           "steps": n_train, "gpu_launches": (ts.kernels_per_step - 2) * n_train,
                           "mode": "train(): dropout on, BatchNorm batch statistics, forward(istrain=True) + the reference's six loss terms + backward, "
                                   + ("eager launches" if args.eager else "one CUDA graph replay per step") + "; no collective, no optimiser"}
                train_scalars["fwd_bwd_ms"] = fwd_bwd["ms_per_step"]
                train_scalars["fwd_bwd_scenes_per_s"] = fwd_bwd["value"]
            l0 = ops.launch_count()
            tms = timed_device(lambda i: full_step(resident[i % n_batches]), n_train, 2)
            last_loss = float(full_step(resident[0])[0].item())
            train_launches = (ops.launch_count() - l0) if args.eager else ts.kernels_per_step * n_train
            train_line = {"value": round(world * scenes * n_train / (tms * 1e-3), 2), "unit": UNIT, "ms_per_step": round(tms / n_train, 3),
                          "steps": n_train, "gpu_launches": train_launches, "grad_allreduce_bytes_per_step": treducer.last_bytes if world > 1 else 0,
                          "loss_first": round(first_loss, 5), "loss_last": round(last_loss, 5),
                          "mode": "process_train up to and including backward(): train-mode forward, the reference's six loss terms on synthetic "
                                  "targets (text embedding provided), backward, " + ("NCCL all-reduce (mean) of all gradients, " if world > 1 else "") +
                                  "fused multi-tensor AdamW with the reference's 13 parameter groups + cosine schedule; "
                                  + ("eager launches" if args.eager else "one CUDA graph replay + one optimiser launch per step")}
            train_scalars.update(train_step_ms=train_line["ms_per_step"], train_step_scenes_per_s=train_line["value"],
                                 train_loss_first=train_line["loss_first"], train_loss_last=train_line["loss_last"])
            # ---- the collective alone: pack + NCCL all-reduce of this step's gradients, CUDA events, max over ranks
            if world > 1:
                for _ in range(3):
                    treducer.replay()
                barrier()
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(10):
                    treducer.replay()                  # pack (one launch, 1 / N folded in) + NCCL all-reduce of the flat buffer
                b_.record()
                barrier()
                train_scalars["grad_allreduce_ms"] = round(max_ranks(a.elapsed_time(b_) / 10), 4)
                train_scalars["grad_allreduce_bytes"] = treducer.last_bytes
            # ---- training step end to end: batch AND targets from pinned host memory every step, loss read back
            if args.metric == "train_step":
                for b in host:
                    t_ = make_targets(b, tgen, "cpu")
                    tgt_host[id(b)] = tuple(x.pin_memory() for x in t_)
                loss_host = torch.empty((), dtype=torch.float32).pin_memory()

                def e2e_step(i):
                    hb = host[i % n_batches]
                    db = hb.to(dev, non_blocking=True)
                    dt = tuple(x.to(dev, non_blocking=True) for x in tgt_host[id(hb)])
                    loss = ts.step(*db.forward_args(), *dt, scene_stats=stats_of(resident[i % n_batches]))[0]
                    loss_host.copy_(loss.reshape(()), non_blocking=True)
                    torch.cuda.synchronize()
                for i in range(2):
                    e2e_step(i)
                barrier()
                t0 = time.perf_counter()
                for i in range(n_train):
                    e2e_step(i)
                barrier()
                e2e_s = max_ranks(time.perf_counter() - t0)
                h2d = host[0].nbytes() + sum(x.numel() * x.element_size() for x in tgt_host[id(host[0])])
                train_e2e_line = {"value": round(world * scenes * n_train / e2e_s, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}
        except Exception as exc:          # the forward numbers above stay valid; rank-local failures are reported, not fatal
            train_line = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            train_scalars["train_step_error"] = train_line["error"]

    # ================= per-kernel pass for the roofline (rank 0) =================
    roofline, kernels = None, {}
    if rank == 0:
        timer = ops.KernelTimer()
        ops.set_timer(timer)
        n_pass = min(args.steps, 10)
        if args.metric == "train_step" and not args.no_train:
            model.train()
            cot2 = None
            for i in range(n_pass):
                torch.cuda._sleep(30_000_000)
                model.zero_grad(set_to_none=True)
                outs = model(*resident[i % n_batches].forward_args(), istrain=True)
                if cot2 is None:
                    cot2 = [torch.randn_like(o) / o.numel() for o in outs[:7]]
                sum((o * c).sum() for o, c in zip(outs[:7], cot2)).backward()
            model.zero_grad(set_to_none=True)
            model.eval()
        else:
            for i in range(n_pass):
                flush.zero_()
                # park the GPU behind a ~15 ms spin so the host enqueues the whole step ahead of it: the events then
                # bracket back-to-back GPU execution, not host launch gaps
                torch.cuda._sleep(30_000_000)
                step_eager(resident[i % n_batches])
        torch.cuda.synchronize()
        ops.set_timer(None)
        summ = timer.summary()
        step_ms = sum(d["ms"] for d in summ.values())
        hbm_names = ("vlsat_gat_edge_fwd", "vlsat_gat_edge_tc_fwd", "vlsat_permute_rows", "vlsat_tf32_split", "vlsat_add_layernorm_fwd", "vlsat_relu_fwd",
                     "vlsat_edge_descriptor_fwd", "vlsat_row_l2norm_fwd", "vlsat_spatial_tail_fwd", "vlsat_build_csr", "vlsat_scene_ranges")
        for name, d in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
            sec = d["ms"] * 1e-3
            hbm_bound = name in hbm_names
            if hbm_bound:
                ach, peak, unit = d["bytes"] / sec / 1e9, peaks["hbm"], "GB/s"
            else:
                ach, peak, unit = d["flops"] / sec / 1e12, peaks["tensor"], "TFLOP/s"
            kernels[name] = dict(bound="hbm" if hbm_bound else "tensor", achieved=round(ach, 3), peak=peak, unit=unit,
                                 frac=round(ach / peak, 5), share_of_step=round(d["ms"] / step_ms, 4),
                                 us_per_launch=round(d["ms"] * 1e3 / d["launches"], 2),
                                 launches_per_step=d["launches"] / n_pass)
        dom = next(iter(kernels))
        roofline = dict(kernel=dom, **{k: kernels[dom][k] for k in ("bound", "achieved", "peak", "unit", "frac")},
                        traffic=load_traffic(dom), peak_source=peaks["source"],
                        note=("tensor-bound kernels compute in BF16x3 (three bf16 MMAs per product for fp32 parity): 0.33 is the ceiling of frac"
                              if args.dtype == "f32" else "single-pass bf16 MMAs: frac is against the full bf16 peak"))
        # the other kernels the spec names, as scalars next to the dominant one (the driver's records keep `roofline` whole)
        short = {"vlsat_gat_edge_tc_fwd": "gat", "vlsat_gat_edge_fwd": "gat", "vlsat_linear_fwd[128x128]": "linear_128x128",
                 "vlsat_linear_fwd[128x64]": "linear_128x64", "vlsat_flash_attn_bf16x3_fwd": "flash",
                 "vlsat_pointnet_tc_fwd": "pointnet", "vlsat_flash_attn_bf16x3_bwd": "flash_bwd", "vlsat_gemm_pairs": "gemm_bwd"}
        for name, tag in short.items():
            if name in kernels:
                k = kernels[name]
                roofline[f"{tag}_frac"] = k["frac"]
                roofline[f"{tag}_achieved"] = k["achieved"]
                roofline[f"{tag}_unit"] = k["unit"]
                roofline[f"{tag}_us_per_launch"] = k["us_per_launch"]
                roofline[f"{tag}_share_of_step"] = k["share_of_step"]
        # the projection entry point as a whole (both tile kernels + the FFMA engine): what round 1 reported as one kernel
        lin = [(n_, summ[n_]) for n_ in summ if n_.startswith("vlsat_linear_fwd")]
        if lin:
            ms_, fl_, la_ = sum(d["ms"] for _, d in lin), sum(d["flops"] for _, d in lin), sum(d["launches"] for _, d in lin)
            roofline["linear_all_frac"] = round(fl_ / (ms_ * 1e-3) / 1e12 / peaks["tensor"], 5)
            roofline["linear_all_share_of_step"] = round(ms_ / step_ms, 4)
            roofline["linear_all_launches_per_step"] = la_ / n_pass
        # the step as a whole: algorithmic tensor work of one step (every tensor-bound entry point) over the measured step time
        # of the graph replay - per-kernel fractions understate the machine's use when kernels of two streams overlap
        tens = sum(d["flops"] for n_, d in summ.items() if n_ not in hbm_names) / n_pass
        whole_ms = train_line["ms_per_step"] if (args.metric == "train_step" and train_line) else fwd_ms
        roofline["step_tensor_gflop"] = round(tens / 1e9, 1)
        roofline["step_tensor_frac"] = round(tens / (whole_ms * 1e-3) / 1e12 / peaks["tensor"], 5)
        if "gat_frac" in roofline:
            gname = "vlsat_gat_edge_tc_fwd" if "vlsat_gat_edge_tc_fwd" in kernels else "vlsat_gat_edge_fwd"
            roofline["gat_traffic"] = load_traffic(gname)
            roofline["gat_bound"] = "hbm (GAT scatter: algorithmic bytes per call, SURVEY.md 8d)"

    if world > 1:
        dist.barrier()
    clocks = clk.summary()
    clk.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = time_reference(args.workload, args.metric, steps=2, warmup=1, budget_s=30.0)
        cpu = dict(value=round(r["value"], 4), unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"])
    headline_train = args.metric == "train_step" and train_line is not None and "error" not in train_line
    if args.metric == "train_step" and not headline_train:
        raise SystemExit(f"train_step leg failed: {train_line}")
    config = {"workload": workload_text(args.workload, kw, scenes, args.metric, args.dtype),
              "global_scenes": world * scenes,
              "parallelism": f"scene-sharded x{world}, " + ("NCCL all-reduce (mean) of the gradients, none in the forward" if headline_train else "no data-path collective"),
              "l2": "flushed between timed steps (256 MB write)", "gemm_engine": ops.gemm_engine(),
              "launch": "eager C-ABI launches" if args.eager else "CUDA graph replay of the C-ABI launches",
              "e2e_loop": "one stream, sync per step" if (args.eager or os.environ.get("VLSAT_E2E", "") == "serial") else
                          "StreamedInference: H2D / replay / D2H of neighbouring steps overlapped on three streams, every step's inputs from pinned host memory and its logits read back",
              "tolerance": ("parity tests: rtol 1e-3 + atol 1e-5 on probabilities; object logits (cross zero) rtol 1e-3 + 1e-4 x max|ref| on the BF16x3 engine"
                            if args.dtype == "f32" else
                            "single-pass bf16 vs the fp32 reference (tests/test_bf16_mode_gpu.py): probabilities |err| <= 2.5e-2, object logits |err| <= 3e-2 x max|ref|"),
              "fwd_ms": round(fwd_ms, 4), "fwd_scenes_per_s": round(fwd_value, 2), "fwd_e2e_scenes_per_s": fwd_e2e_line["value"]}
    config.update(train_scalars)
    config.update(prep)
    config.update(yard)
    if yard.get("ref_gpu_kind"):
        config["ref_gpu_note"] = ("reference PyTorch forward on the same B200 (python edge / scene loops of SGFN_MMG/model.py:260-265 and network_MMG.py:183-205 "
                                  "included when kind = reference); ours skips the pair projector in eval (its output is not returned), the reference computes it")
    if headline_train:
        line = {"metric": METRICS["train_step"], "value": train_line["value"], "unit": UNIT, "n_gpus": world, "steps": n_train,
                "warmup": args.warmup, "ms_per_step": train_line["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": config, "e2e": train_e2e_line,
                "gpu_launches": train_launches}
    else:
        line = {"metric": METRICS["fwd"], "value": round(fwd_value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(fwd_ms, 4), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": config, "e2e": fwd_e2e_line,
                "gpu_launches": fwd_launches}
    line.update({"clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "kernels": kernels, "fwd_bwd": fwd_bwd, "train_step": train_line})
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def load_traffic(kernel: str):
    """DRAM bytes per launch from the committed ncu capture (profiles/traffic.json), else null."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(kernel)
    return None


def run_reference(args, device: str):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_reference(args.workload, args.metric, steps=args.steps, warmup=max(args.warmup, 1), budget_s=240.0, device=device)
    kw = workload_kwargs(args.workload)
    line = {
        "impl": "reference" if device == "cpu" else "reference-gpu", "metric": METRICS[args.metric], "value": round(r["value"], 4), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.workload, kw, kw["num_scenes"], args.metric),
                   "sample_scenes": r["scenes"], "device": "host CPU" if device == "cpu" else "cuda:0", "reference_kind": r["kind"],
                   "note": "the reference evaluates generate_object_pair_features + triplet_projector_2d in eval mode and discards the result "
                           "(SGFN_MMG/model.py:319-322); the B200 arm computes them only when istrain=True"},
        "cpu_baseline": {"value": round(r["value"], 4), "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": round(r["value"], 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--metric", default="fwd", choices=list(METRICS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"], help="f32 = BF16x3 (fp32 parity); bf16 = single-pass bf16 MMAs (BASELINE configs #3 / #4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training legs")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-on-this-GPU yardstick")
    ap.add_argument("--eager", action="store_true", help="launch kernel by kernel instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args, "cpu")
    elif args.impl == "reference-gpu":
        run_reference(args, "cuda:0")
    else:
        import __graft_entry__ as g
        g.build()
        run_b200(args)


if __name__ == "__main__":
    main()
