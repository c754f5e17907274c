"""Thin Python wrappers over the C ABI (include/vlsat_b200.h): argument checking, output allocation
with torch (device memory + current stream are PyTorch's job here, nothing else), then one C call.

No wrapper has a CPU or PyTorch compute path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import contextlib
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import Epilogue, LinearOpts

ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
AGGR = {"max": 0, "add": 1, "mean": 2}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (vlsat_b200 has no CPU path), got {t.device}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    return t


def _rows(t: torch.Tensor, name: str) -> Tuple[int, int]:
    """2-D tensor with unit inner stride -> (data_ptr, leading dimension)."""
    _f32(t, name)
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError(f"{name}: expected a 2-D tensor with contiguous rows, got shape {tuple(t.shape)} strides {t.stride()}")
    ld = t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])
    return t.data_ptr(), ld


def _i64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.int64 or not t.is_contiguous():
        raise TypeError(f"{name}: expected a contiguous CUDA int64 tensor, got {t.dtype} on {t.device}")
    return t


class KernelTimer:
    """Optional per-entry-point CUDA-event timing (bench.py's roofline leg). Off by default."""

    def __init__(self):
        self.events = []          # (name, start, end, flops, bytes)

    def summary(self):
        """name -> dict(ms, launches, flops, bytes): totals over everything recorded."""
        out = {}
        for name, a, b, fl, by in self.events:
            d = out.setdefault(name, dict(ms=0.0, launches=0, flops=0.0, bytes=0.0))
            d["ms"] += a.elapsed_time(b); d["launches"] += 1; d["flops"] += fl; d["bytes"] += by
        return out


_timer: Optional[KernelTimer] = None


def set_timer(t: Optional[KernelTimer]) -> None:
    global _timer
    _timer = t


def _call(name: str, *args, work=(0.0, 0.0), tag: str = ""):
    """One C-ABI call. ``work`` = (algorithmic FLOPs, algorithmic HBM bytes) of this launch, used only
    by the optional timer (definitions: DESIGN.md section 'Kernels and their rooflines'); ``tag`` distinguishes the kernel
    instantiations behind one entry point in the timer's table (e.g. the 128x128 and 128x64 tile kernels of the projections)."""
    fn = getattr(_lib.load(), name)
    if _timer is None:
        return fn(*args)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    st = fn(*args)
    b.record()
    _timer.events.append((name + tag, a, b, float(work[0]), float(work[1])))
    return st


# ------------------------------------------------------------------------------------------ zero arena
# The backward pass accumulates into ~150 small zero-initialised buffers per step (bias gradients, LayerNorm dgamma / dbeta,
# scatter-add targets, slab-free weight gradients): one fill launch each. Inside ``zero_arena`` they are carved out of ONE buffer
# that is zeroed by one fill at the start of the step. The buffer is allocated afresh every step (inside a CUDA-graph capture:
# once, as a node of the graph), so slices never alias across steps; gradients that autograd keeps as ``.grad`` keep their arena
# alive. Sized from the previous step's demand; anything that does not fit falls back to ``torch.zeros``.
ZERO_ARENA_MAX_ITEM = 2 << 20          # bytes: larger buffers are bandwidth-sized fills anyway


class _ZeroArena:
    def __init__(self, hint_bytes: int, device):
        self.demand = 0                 # bytes asked for during this step (256-byte granules): next step's size
        self.off = 0
        self.fallbacks = 0
        # allocated up front on the step's main stream: every later use, on any stream, is ordered behind this fill
        self.buf = torch.zeros((hint_bytes // 4,), device=device, dtype=torch.float32) if hint_bytes > 0 and device is not None else None


_arena: Optional[_ZeroArena] = None


@contextlib.contextmanager
def zero_arena(owner, device):
    """``owner`` carries the size hint from step to step (attribute ``_zero_arena_bytes``). VLSAT_ZERO_ARENA=0 disables."""
    global _arena
    if os.environ.get("VLSAT_ZERO_ARENA", "1") == "0" or torch.device(device).type != "cuda":
        yield None
        return
    prev = _arena
    a = _ZeroArena(int(getattr(owner, "_zero_arena_bytes", 0)), device)
    _arena = a
    try:
        yield a
    finally:
        _arena = prev
        owner._zero_arena_bytes = a.demand
        owner._zero_arena_fallbacks = a.fallbacks


def zeros(shape, device) -> torch.Tensor:
    """fp32 zeros for an accumulation target of the backward pass: a slice of the step's arena when one is active."""
    a = _arena
    numel = 1
    for d in shape:
        numel *= int(d)
    nbytes = numel * 4
    if a is None or nbytes == 0 or nbytes > ZERO_ARENA_MAX_ITEM:
        return torch.zeros(shape, device=device, dtype=torch.float32)
    granule = (nbytes + 255) // 256 * 256
    a.demand += granule
    if a.buf is None or a.buf.device != torch.device(device) or a.off + granule > a.buf.numel() * 4:
        a.fallbacks += 1
        return torch.zeros(shape, device=device, dtype=torch.float32)
    out = a.buf[a.off // 4: a.off // 4 + numel].view(shape)
    a.off += granule
    return out


# ------------------------------------------------------------------------------------------ two-stream regions
_side_streams = {}


def two_streams() -> bool:
    """Independent branches of the inference forward run on two streams (default; VLSAT_STREAMS=1 serialises them). Off while
    the per-kernel timer is active, whose events must bracket serial execution. One CTA per SM for every tensor-core kernel,
    so the second kernel's CTAs start as the first one's exit: the partial last round of its tiles fills for free."""
    return os.environ.get("VLSAT_STREAMS", "2") == "2" and _timer is None


def side_stream(device) -> torch.cuda.Stream:
    key = str(device)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def fork_join(side_fn, main_fn, device):
    """(side_fn(), main_fn()) with ``side_fn`` on THE side stream next to ``main_fn`` on the current one. Allocator rules that
    make this safe with PyTorch's caching allocator (also inside a CUDA-graph capture): the side branch is issued FIRST;
    every tensor it reads must stay referenced by the caller until this function returns (so the main branch cannot be
    handed a block the side branch still reads); there is ONE side stream, every use of which starts by waiting for an
    event recorded on the main stream (so a block the side pool hands out again is ordered after its last main-stream use)."""
    if not two_streams() or torch.device(device).type != "cuda":    # CPU tensors: serial, the kernels' own checks raise
        return side_fn(), main_fn()
    main, side = torch.cuda.current_stream(), side_stream(device)
    fork = torch.cuda.Event()
    fork.record(main)
    side.wait_event(fork)
    with torch.cuda.stream(side):
        a = side_fn()
        join = torch.cuda.Event()
        join.record(side)
    b = main_fn()
    main.wait_event(join)
    return a, b


def launch_count() -> int:
    return int(_lib.load().vlsat_launch_count())


def gemm_engine() -> str:
    return _lib.load().vlsat_gemm_engine().decode()


# ------------------------------------------------------------------------------------------ encoders
def pointnet(x, w1, b1, w2, b2, w3, b3, want_argmax: bool = False):
    """A1. x [n_obj, c_in, P] -> [n_obj, c_out]; weights are the squeezed Conv1d(k=1) weights."""
    _f32(x, "x")
    if x.dim() != 3 or not x.is_contiguous():
        raise ValueError("pointnet: x must be a contiguous [n_obj, c_in, n_pts] tensor")
    n_obj, c_in, n_pts = x.shape
    for n, t in (("w1", w1), ("b1", b1), ("w2", w2), ("b2", b2), ("w3", w3), ("b3", b3)):
        _f32(t, n)
        if not t.is_contiguous():
            raise ValueError(f"pointnet: {n} must be contiguous")
    c1, c2, c_out = w1.shape[0], w2.shape[0], w3.shape[0]
    if w1.shape[1] != c_in or w2.shape[1] != c1 or w3.shape[1] != c2:
        raise ValueError("pointnet: weight shapes do not chain")
    out = torch.empty((n_obj, c_out), device=x.device, dtype=torch.float32)
    arg = torch.empty((n_obj, c_out), device=x.device, dtype=torch.int32) if want_argmax else None
    # the tensor-core encoder is BF16x3 only: with the 3xTF32 engine forced ('tc') the exact FFMA kernel runs instead
    use_tc = _engine in (ENGINES["auto"], ENGINES["bf16x3"]) and c1 == 64 and c2 == 128 and c_in <= 16 and c_out % 128 == 0
    st = _call("vlsat_pointnet_tc_fwd" if use_tc else "vlsat_pointnet_fwd", x.data_ptr(), n_obj, c_in, n_pts, w1.data_ptr(), b1.data_ptr(), c1,
                                        w2.data_ptr(), b2.data_ptr(), c2, w3.data_ptr(), b3.data_ptr(), c_out,
                                        out.data_ptr(), arg.data_ptr() if want_argmax else None, _stream(),
               work=(2.0 * n_obj * n_pts * (c_in * c1 + c1 * c2 + c2 * c_out), 4.0 * (x.numel() + n_obj * c_out)))
    _lib.check(st, "vlsat_pointnet_tc_fwd" if use_tc else "vlsat_pointnet_fwd")
    return (out, arg) if want_argmax else out


def edge_descriptor(desc: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """A2. desc [N, 11], edge_index [2, E] int64 -> [E, 11]."""
    _f32(desc, "descriptor"); _i64(edge_index, "edge_index")
    if desc.dim() != 2 or desc.shape[1] != 11 or not desc.is_contiguous():
        raise ValueError("edge_descriptor: descriptor must be contiguous [N, 11]")
    if edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError("edge_descriptor: edge_index must be [2, E]")
    e = edge_index.shape[1]
    out = torch.empty((e, 11), device=desc.device, dtype=torch.float32)
    _lib.check(_call("vlsat_edge_descriptor_fwd", desc.data_ptr(), desc.shape[0], edge_index.data_ptr(), e,
                                                     out.data_ptr(), _stream()), "vlsat_edge_descriptor_fwd")
    return out


# ------------------------------------------------------------------------------------- dense projection
ENGINES = {"auto": 0, "simt": 1, "tc": 2, "tc1": 3, "bf16x3": 4}
FMT_TF32, FMT_BF16 = 0, 1            # (hi, lo) pair formats: VLSAT_SPLIT_* of include/vlsat_b200.h
_engine = ENGINES[os.environ.get("VLSAT_GEMM_ENGINE", "auto")]
_weight_splits = {}


def set_gemm_engine(name: str) -> None:
    """'auto' (tcgen05 BF16x3 where bf16 pairs are TMA-addressable, 3xTF32 next, FFMA otherwise), 'simt' (exact fp32
    FFMA everywhere), 'bf16x3' / 'tc' (force that tensor-core engine; ineligible shapes fall to the next one),
    'tc1' (single-pass TF32, not fp32-accurate)."""
    global _engine
    _engine = ENGINES[name]


PRECISIONS = {"fp32": 0, "bf16": 1}


def set_precision(name: str) -> None:
    """Arithmetic of the tensor-core kernels: 'fp32' (default; BF16x3, fp32 parity with the reference) or 'bf16' (one MMA
    per product on the hi halves of the operand pairs - BASELINE configs #3 / #4; tolerance in tests/test_bf16_mode_gpu.py)."""
    _lib.check(_lib.load().vlsat_set_precision(PRECISIONS[name]), "vlsat_set_precision")


def precision() -> str:
    return "bf16" if int(_lib.load().vlsat_get_precision()) == 1 else "fp32"


def _elem_bytes() -> float:
    return 2.0 if precision() == "bf16" else 4.0


def tensor_cores_enabled() -> bool:
    return _engine != ENGINES["simt"]


def pair_fmt(pair) -> int:
    return FMT_BF16 if pair[0].dtype == torch.bfloat16 else FMT_TF32


def default_fmt(k: int) -> int:
    """Pair format the dense-projection engine wants for an operand with ``k`` columns."""
    return FMT_BF16 if (_engine in (ENGINES["auto"], ENGINES["bf16x3"]) and k % 8 == 0) else FMT_TF32


def split_pair(x: torch.Tensor, fmt: Optional[int] = None):
    """(hi, lo) pair of an fp32 matrix in format ``fmt`` (default: what ``linear`` wants for this width), compact."""
    fmt = default_fmt(x.shape[1]) if fmt is None else fmt
    return bf16_split(x) if fmt == FMT_BF16 else tf32_split(x)


def _tc_eligible(x, ldx, w, ldw, n, k, fmt) -> bool:
    if fmt == FMT_BF16:
        return k % 8 == 0 and k >= 32 and n >= 8
    return (k % 4 == 0 and k >= 32 and ldx % 4 == 0 and ldw % 4 == 0 and x.data_ptr() % 16 == 0
            and w.data_ptr() % 16 == 0 and n >= 8)


def _cacheable_weight(w: torch.Tensor) -> bool:
    """A weight whose split may be cached by storage address: a parameter, a view of one, or a persistent derived tensor
    (no autograd history). Temporaries built inside a differentiable forward (e.g. the head-major permutations of
    train_path, a fresh copy per call) are NOT: caching them would pin one buffer per call for ever."""
    base = w if w._base is None else w._base
    return isinstance(base, torch.nn.Parameter) or (w.grad_fn is None and not w.requires_grad)


def weight_pair(w: torch.Tensor, fmt: int = FMT_BF16):
    """(hi, lo) pair of a weight matrix [N, K] (row stride allowed): cached per parameter version when ``w`` is
    parameter-backed, split per call otherwise."""
    _, ldw = _rows(w, "w")
    return _weight_split(w, ldw, fmt) if _cacheable_weight(w) else split_pair(w, fmt)


def act_pair(x: torch.Tensor):
    """bf16 (hi, lo) pair of an activation, remembered on the tensor object while its version stands: the projections that
    read the same tensor (q / k / v of one input, forward and backward of one layer) share one split pass."""
    ent = getattr(x, "_vlsat_pair", None)
    if ent is not None and ent[0] == x._version and ent[1][0].shape == x.shape:
        return ent[1]
    pair = bf16_split(x)
    try:
        x._vlsat_pair = (x._version, pair)
    except AttributeError:
        pass
    return pair


def view_rows(x: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    """``x.view(rows, cols)`` that keeps the remembered bf16 pair: [E, H * d] seen as (edge, head) rows [E * H, d] is the same
    memory for the compact pair tensors too, so the projection over the rows needs no split pass of its own."""
    y = x.view(rows, cols)
    ent = getattr(x, "_vlsat_pair", None)
    if ent is not None and ent[0] == x._version and ent[1][0].shape == x.shape and ent[1][0].is_contiguous() and ent[1][1].is_contiguous():
        try:
            y._vlsat_pair = (y._version, (ent[1][0].view(rows, cols), ent[1][1].view(rows, cols)))
        except AttributeError:
            pass
    return y


def _weight_split(w: torch.Tensor, ldw: int, fmt: int):
    """(hi, lo) copies of a weight (view) in pair format ``fmt``, cached until the parameter is written again."""
    key = (w.data_ptr(), tuple(w.shape), ldw, fmt)
    ent = _weight_splits.get(key)
    if ent is not None and ent[0] == w._version:
        return ent[1]
    n, k = w.shape
    if fmt == FMT_BF16:
        hl = ent[1] if ent is not None else torch.empty((2, n, k), device=w.device, dtype=torch.bfloat16)
        _lib.check(_call("vlsat_bf16_split", w.data_ptr(), ldw, n, k, hl[0].data_ptr(), hl[1].data_ptr(), k, _stream()),
                   "vlsat_bf16_split")
    else:
        hl = ent[1] if ent is not None else torch.empty((2, n, k), device=w.device, dtype=torch.float32)
        _lib.check(_call("vlsat_tf32_split", w.data_ptr(), ldw, n, k, hl[0].data_ptr(), hl[1].data_ptr(), _stream()),
                   "vlsat_tf32_split")
    _weight_splits[key] = (w._version, hl, w)        # holding w keeps its storage (and this key) unique
    return hl


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
           out: Optional[torch.Tensor] = None,
           gather: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]] = None,
           residual: Optional[torch.Tensor] = None, alpha: float = 1.0, beta: float = 1.0,
           scale_ptr: Optional[torch.Tensor] = None, bias_per_row: bool = False,
           x_is_weight: bool = False, x_split=None, w_split=None, emit_split=False, want_y: bool = True,
           cache_w: bool = True):
    """y = post(act(x w^T + bias + ga[ia] + gb[ib])), post(t) = (alpha t + beta residual) * exp(scale).

    x [M, K] and w [N, K] may be column-slice views (row stride = leading dimension); ``out`` may be a
    column slice of a wider buffer. ``x_is_weight=True`` swaps the roles for the split cache: x is
    a parameter (split cached), w an activation (split per call) - used to emit y^T = W x^T.
    ``x_split=(hi, lo)``: pair of x already available (compact [M, K]; bf16 or tf32 pairs, the engine follows the
    format) - skips the split pass. ``emit_split``: the epilogue also writes the (hi, lo) pair of y and the call returns
    ``(y, (hi, lo))`` - ``True`` = the format a following projection wants, or ``FMT_TF32`` / ``FMT_BF16`` given as
    ``'tf32'`` / ``'bf16'`` (``want_y=False``: only the pair is written, y is None). ``cache_w=False``: do not cache the
    split of ``w`` (it is an activation or a temporary). ``x`` itself may be such a ``(hi, lo)`` pair when the unsplit
    activation was never materialised (tensor-core engines only)."""
    x_pair_only = isinstance(x, tuple)
    if x_pair_only:
        x_split = x
        x = x_split[0]
        if not (tensor_cores_enabled() and x.shape[1] % 4 == 0 and x.shape[1] >= 32 and w.shape[0] >= 8):
            raise RuntimeError("linear: a pre-split activation can only feed the tensor-core engine")
        xp, ldx = x.data_ptr(), x.shape[1]
    else:
        xp, ldx = _rows(x, "x")
    wp, ldw = _rows(w, "w")
    m, k = x.shape
    n = w.shape[0]
    if w.shape[1] != k:
        raise ValueError(f"linear: x is [{m},{k}] but w is {tuple(w.shape)}")
    if not want_y and not emit_split:
        raise ValueError("linear: nothing to compute")
    yp, ldy = None, n
    if want_y:
        if out is None:
            out = torch.empty((m, n), device=x.device, dtype=torch.float32)
        elif tuple(out.shape) != (m, n):
            raise ValueError(f"linear: out has shape {tuple(out.shape)}, expected {(m, n)}")
        yp, ldy = _rows(out, "out")
    else:
        out = None
    epi = Epilogue()
    y_split = None
    if emit_split:
        efmt = default_fmt(n) if emit_split is True else {"tf32": FMT_TF32, "bf16": FMT_BF16}[emit_split]
        if n % (8 if efmt == FMT_BF16 else 4):
            raise ValueError("linear: emit_split needs N % 4 == 0 (tf32 pairs) / N % 8 == 0 (bf16 pairs)")
        y_split = torch.empty((2, m, n), device=x.device, dtype=torch.bfloat16 if efmt == FMT_BF16 else torch.float32)
        epi.split_hi, epi.split_lo, epi.ld_split, epi.split_fmt = y_split[0].data_ptr(), y_split[1].data_ptr(), n, efmt
    epi.alpha, epi.beta, epi.act = alpha, beta, act
    if bias is not None:
        _f32(bias, "bias")
        if bias.numel() != (m if bias_per_row else n) or not bias.is_contiguous():
            raise ValueError("linear: bias must be contiguous [N] (or [M] with bias_per_row)")
        epi.bias = bias.data_ptr()
        epi.bias_per_row = int(bias_per_row)
    if gather is not None:
        ga, ia, gb, ib = gather
        gap, ldga = _rows(ga, "gather_a")
        _i64(ia, "idx_a")
        if ga.shape[1] != n or ia.numel() != m:
            raise ValueError("linear: gather operands must be [*, N] with [M] indices")
        epi.gather_a, epi.idx_a, epi.ld_gather = gap, ia.data_ptr(), ldga
        if gb is not None:                       # second gather term is optional
            gbp, ldgb = _rows(gb, "gather_b")
            _i64(ib, "idx_b")
            if ldga != ldgb or gb.shape[1] != n or ib.numel() != m:
                raise ValueError("linear: gather operands must be [*, N] with equal row strides and [M] indices")
            epi.gather_b, epi.idx_b = gbp, ib.data_ptr()
    if residual is not None:
        rp, ldr = _rows(residual, "residual")
        if tuple(residual.shape) != (m, n):
            raise ValueError("linear: residual must be [M, N]")
        epi.residual, epi.ld_res = rp, ldr
    if scale_ptr is not None:
        _f32(scale_ptr, "scale_ptr")
        epi.scale_ptr = scale_ptr.data_ptr()
    opts = LinearOpts()
    opts.engine = ENGINES["simt"]
    ws = None
    # operand pair format: given pairs decide; otherwise the engine's preference for this K
    given = x_split if x_split is not None else w_split
    fmt = pair_fmt(given) if given is not None else default_fmt(k)
    if x_split is not None and w_split is not None and pair_fmt(x_split) != pair_fmt(w_split):
        raise ValueError("linear: x_split and w_split are in different pair formats")
    if _engine == ENGINES["tc1"] and given is None:
        fmt = FMT_TF32
    if fmt == FMT_BF16 and not _tc_eligible(x, ldx, w, ldw, n, k, FMT_BF16) and given is None:
        fmt = FMT_TF32
    if _engine != ENGINES["simt"] and m > 0 and _tc_eligible(x, ldx, w, ldw, n, k, fmt):
        opts.engine = ENGINES["bf16x3"] if fmt == FMT_BF16 else (ENGINES["tc1"] if _engine == ENGINES["tc1"] else ENGINES["tc"])
        if x_is_weight:
            hl = _weight_split(x, ldx, fmt)
            opts.x_hi, opts.x_lo = hl[0].data_ptr(), hl[1].data_ptr()
            if w_split is not None:
                opts.w_hi, opts.w_lo = w_split[0].data_ptr(), w_split[1].data_ptr()
            else:
                ws = torch.empty((2 * n * k,), device=x.device, dtype=torch.float32)
        else:
            # cache_w=False: w is an activation / a per-call temporary (backward GEMMs), never cache its split
            hl = w_split if w_split is not None else (_weight_split(w, ldw, fmt) if (cache_w and _cacheable_weight(w)) else split_pair(w, fmt))
            opts.w_hi, opts.w_lo = hl[0].data_ptr(), hl[1].data_ptr()
            if x_split is not None:
                opts.x_hi, opts.x_lo = x_split[0].data_ptr(), x_split[1].data_ptr()
            else:
                ws = torch.empty((2 * m * k,), device=x.device, dtype=torch.float32)
        if ws is not None:
            opts.workspace, opts.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    elif x_pair_only:
        raise RuntimeError("linear: a pre-split activation can only feed the tensor-core engine")
    # which kernel runs (csrc/gemm_tc.cu linear_tc): 128 x 64 tiles for narrow outputs and for problems whose 128 x 128 tiling
    # would leave more than half of the SMs idle, 128 x 128 tiles otherwise; the FFMA engine for shapes pairs cannot address
    if opts.engine == ENGINES["simt"]:
        tag = "[ffma]"
    else:
        tiles128 = ((m + 127) // 128) * ((n + 127) // 128)
        tag = "[128x64]" if (n <= 64 or 2 * tiles128 <= 148) else "[128x128]"
    st = _call("vlsat_linear_fwd", xp, ldx, wp, ldw, yp, ldy, m, n, k, C.byref(epi), C.byref(opts), _stream(),
               work=(2.0 * m * n * k, 4.0 * (m * k + n * k + m * n)), tag=tag)
    _lib.check(st, "vlsat_linear_fwd")
    if emit_split:
        return out, (y_split[0], y_split[1])
    return out


GEMM_NN, GEMM_TN = 2, 3


def gemm_pairs_ok(*pairs) -> bool:
    """Operands the stored-operand backward GEMMs can address: bf16 pairs with 16-byte aligned rows."""
    return tensor_cores_enabled() and all(p is not None and p[0].dtype == torch.bfloat16 and p[0].stride(0) % 8 == 0
                                          and p[0].data_ptr() % 16 == 0 and p[1].data_ptr() % 16 == 0 for p in pairs)


def gemm_nn(a_pair, b_pair, n: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y [M, n] = a [M, K] . b [K, n] on bf16 (hi, lo) pairs read as stored (dX = dZ W with W the forward's weight pair)."""
    m, k = a_pair[0].shape[0], b_pair[0].shape[0]
    if out is None:
        out = torch.empty((m, n), device=a_pair[0].device, dtype=torch.float32)
    yp, ldy = _rows(out, "out")
    st = _call("vlsat_gemm_pairs", GEMM_NN, a_pair[0].data_ptr(), a_pair[1].data_ptr(), a_pair[0].stride(0), b_pair[0].data_ptr(),
               b_pair[1].data_ptr(), b_pair[0].stride(0), yp, ldy, m, n, k, None, 0, _stream(), work=(2.0 * m * n * k, 4.0 * (m * k + n * k + m * n)))
    _lib.check(st, "vlsat_gemm_pairs(NN)")
    return out


def gemm_tn(a_pair, b_pair, m: int, n: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y [m, n] = a^T b with a stored [K, m], b stored [K, n] (dW = dZ^T X): the reduction runs over the stored rows and is
    split over CTAs for small outputs (deterministic slab sums)."""
    k = a_pair[0].shape[0]
    if out is None:
        out = torch.empty((m, n), device=a_pair[0].device, dtype=torch.float32)
    yp, ldy = _rows(out, "out")
    ws_bytes = int(_lib.load().vlsat_gemm_pairs_workspace_bytes(GEMM_TN, m, n, k))
    ws = torch.empty((ws_bytes // 4,), device=out.device, dtype=torch.float32) if ws_bytes else None
    st = _call("vlsat_gemm_pairs", GEMM_TN, a_pair[0].data_ptr(), a_pair[1].data_ptr(), a_pair[0].stride(0), b_pair[0].data_ptr(),
               b_pair[1].data_ptr(), b_pair[0].stride(0), yp, ldy, m, n, k, ws.data_ptr() if ws is not None else None, ws_bytes, _stream(),
               work=(2.0 * m * n * k, 4.0 * (m * k + n * k + m * n)))
    _lib.check(st, "vlsat_gemm_pairs(TN)")
    return out


def linear_chain(x: torch.Tensor, layers, x_split=None, emit_last: bool = False, **last_kw):
    """Row MLP ``x -> act(x w0^T + b0) -> ...``: ``layers`` = [(w, b, act), ...]. On the tensor-core engines an intermediate
    activation exists only as the (hi, lo) pair the next projection reads (written by the producing epilogue); the last
    layer takes ``last_kw`` (residual=..., out=..., ...) and, with ``emit_last``, also returns the pair of its output."""
    h, hp = x, x_split
    for li, (w, b, act) in enumerate(layers):
        last = li == len(layers) - 1
        n = w.shape[0]
        if last:
            return linear(h if h is not None else hp, w, b, act=act, x_split=hp if h is not None else None, emit_split=emit_last, **last_kw)
        pair_only = tensor_cores_enabled() and n % 8 == 0 and n >= 32 and layers[li + 1][0].shape[0] >= 8
        if pair_only:
            _, hp2 = linear(h if h is not None else hp, w, b, act=act, x_split=hp if h is not None else None, emit_split=True, want_y=False)
            h, hp = None, hp2
        else:
            h, hp = linear(h if h is not None else hp, w, b, act=act, x_split=hp if h is not None else None), None


def add_layernorm(x: torch.Tensor, res: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor,
                  eps: float = 1e-5, relu: bool = False, out: Optional[torch.Tensor] = None, emit_split: bool = False):
    """LayerNorm(x + res) (+ ReLU). ``emit_split=True`` also returns the bf16 (hi, lo) pair of the result (the x operand of
    a following projection): ``(y, (hi, lo))``."""
    xp, ldx = _rows(x, "x")
    m, d = x.shape
    rp, ldr = (None, 0)
    if res is not None:
        rp, ldr = _rows(res, "res")
        if tuple(res.shape) != (m, d):
            raise ValueError("add_layernorm: res shape mismatch")
    if out is None:
        out = torch.empty((m, d), device=x.device, dtype=torch.float32)
    yp, ldy = _rows(out, "out")
    _f32(gamma, "gamma"); _f32(beta, "beta")
    pair = None
    if emit_split and d % 128 == 0 and tensor_cores_enabled() and default_fmt(d) == FMT_BF16:
        pair = torch.empty((2, m, d), device=x.device, dtype=torch.bfloat16)
    st = _call("vlsat_add_layernorm_fwd", xp, ldx, rp, ldr, gamma.data_ptr(), beta.data_ptr(), yp, ldy, m, d,
                                             eps, int(relu), pair[0].data_ptr() if pair is not None else None,
                                             pair[1].data_ptr() if pair is not None else None, d, _stream())
    _lib.check(st, "vlsat_add_layernorm_fwd")
    if emit_split:
        return out, ((pair[0], pair[1]) if pair is not None else None)
    return out


def relu(x: torch.Tensor, out: Optional[torch.Tensor] = None, emit_split: bool = False):
    """y = relu(x); ``emit_split=True`` also returns the bf16 (hi, lo) pair of y (2-D x with cols % 8 == 0): ``(y, pair)``."""
    _f32(x, "x")
    if not x.is_contiguous():
        raise ValueError("relu: x must be contiguous")
    if out is None:
        out = torch.empty_like(x)
    if emit_split and x.dim() == 2 and x.shape[1] % 8 == 0 and tensor_cores_enabled() and default_fmt(x.shape[1]) == FMT_BF16:
        pair = torch.empty((2,) + tuple(x.shape), device=x.device, dtype=torch.bfloat16)
        _lib.check(_call("vlsat_relu_pair_fwd", x.data_ptr(), out.data_ptr(), pair[0].data_ptr(), pair[1].data_ptr(), x.numel(), _stream()),
                   "vlsat_relu_pair_fwd")
        return out, (pair[0], pair[1])
    _lib.check(_call("vlsat_relu_fwd", x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "vlsat_relu_fwd")
    return (out, None) if emit_split else out


def row_l2norm(x: torch.Tensor) -> torch.Tensor:
    _f32(x, "x")
    if x.dim() != 2 or not x.is_contiguous():
        raise ValueError("row_l2norm: x must be contiguous 2-D")
    out = torch.empty_like(x)
    _lib.check(_call("vlsat_row_l2norm_fwd", x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], _stream()),
               "vlsat_row_l2norm_fwd")
    return out


def spatial_tail(desc: torch.Tensor, out: torch.Tensor, col0: int) -> None:
    _f32(desc, "descriptor")
    op, ld = _rows(out, "out")
    if desc.shape[1] != 11 or not desc.is_contiguous() or out.shape[0] != desc.shape[0]:
        raise ValueError("spatial_tail: descriptor must be contiguous [N, 11] and out [N, >= col0 + 8]")
    _lib.check(_call("vlsat_spatial_tail_fwd", desc.data_ptr(), op, ld, col0, desc.shape[0], _stream()),
               "vlsat_spatial_tail_fwd")


# -------------------------------------------------------------------------------------------- attention
FC_PACK_HEAD = 1344   # floats before w2 in the packed self_attn_fc buffer (see csrc/node_attn.cu)


def pack_attn_fc(fc: torch.nn.Sequential) -> torch.Tensor:
    """Pack ``self_attn_fc`` (network_MMG.py:165-173) as
    w0[32,4] b0 g0 be0 | w1[32,32] b1 g1 be1 | w2[H,32] b2[H] in one fp32 vector."""
    parts = [fc[0].weight, fc[0].bias, fc[2].weight, fc[2].bias, fc[3].weight, fc[3].bias,
             fc[5].weight, fc[5].bias, fc[6].weight, fc[6].bias]
    return torch.cat([p.detach().reshape(-1).float() for p in parts]).contiguous()


def scene_ranges(batch_ids: torch.Tensor):
    """A6 bookkeeping: per-node [start, end) of its scene. Returns (seg_start, seg_end, err_flag)."""
    b = _i64(batch_ids.reshape(-1), "batch_ids")
    n = b.numel()
    seg = torch.empty((2, n), device=b.device, dtype=torch.int32)
    err = torch.zeros((1,), device=b.device, dtype=torch.int32)
    _lib.check(_call("vlsat_scene_ranges", b.data_ptr(), n, seg[0].data_ptr(), seg[1].data_ptr(), err.data_ptr(),
                                              _stream()), "vlsat_scene_ranges")
    return seg[0], seg[1], err


def node_bias_table(centres, seg_start, seg_end, fc_pack, n_heads: int) -> torch.Tensor:
    """Distance-bias table of the scene-resident node attention: [H, N, 64], entry (h, a, j) = bias of query a towards the
    j-th node of its scene (scenes of more than 64 nodes are left to the streaming kernel, which evaluates the MLP in place)."""
    cp, ldc = _rows(centres, "centres")
    n = centres.shape[0]
    cap = int(_lib.load().vlsat_node_bias_table_max_scene())
    if fc_pack.numel() != FC_PACK_HEAD + 33 * n_heads:
        raise ValueError("node_bias_table: packed self_attn_fc has the wrong size for this head count")
    tab = torch.empty((n_heads, n, cap), device=centres.device, dtype=torch.float32)
    _lib.check(_call("vlsat_node_bias_table", cp, ldc, seg_start.data_ptr(), seg_end.data_ptr(), _f32(fc_pack, "fc_pack").data_ptr(),
                     n_heads, tab.data_ptr(), n, _stream()), "vlsat_node_bias_table")
    return tab


def node_attn(q, k, v, centres, seg_start, seg_end, fc_pack, n_heads: int, bias_table: Optional[torch.Tensor] = None) -> torch.Tensor:
    """A6 + A7. With ``bias_table`` (``node_bias_table``, computed once per forward) scenes of up to 64 nodes run on the
    scene-resident kernel and only larger ones on the streaming kernel; without it everything streams."""
    qp, ldq = _rows(q, "q"); kp, ldk = _rows(k, "k"); vp_, ldv = _rows(v, "v")
    cp, ldc = _rows(centres, "centres")
    n, d = q.shape
    dk = d // n_heads
    if fc_pack.numel() != FC_PACK_HEAD + 33 * n_heads:
        raise ValueError("node_attn: packed self_attn_fc has the wrong size for this head count")
    out = torch.empty((n, d), device=q.device, dtype=torch.float32)
    skip = 0
    if bias_table is not None and dk in (32, 64) and ldv % 4 == 0 and (qp | kp | vp_) % 16 == 0:
        skip = bias_table.shape[2]
        st = _call("vlsat_node_attn_scene_fwd", qp, ldq, kp, ldk, vp_, ldv, bias_table.data_ptr(), seg_start.data_ptr(),
                   seg_end.data_ptr(), n_heads, dk, out.data_ptr(), d, n, _stream())
        _lib.check(st, "vlsat_node_attn_scene_fwd")
    st = _call("vlsat_node_attn_fwd", qp, ldq, kp, ldk, vp_, ldv, cp, ldc, seg_start.data_ptr(), seg_end.data_ptr(),
                                         _f32(fc_pack, "fc_pack").data_ptr(), n_heads, dk, out.data_ptr(), d, n, skip, _stream())
    _lib.check(st, "vlsat_node_attn_fwd")
    return out


def dense_attn(q, k, v, n_heads: int, weights=None, way: str = "mul", mask=None) -> torch.Tensor:
    """softmax(mask(q k^T / sqrt(dk) (*|+) weights)) v with the reference's dense arguments (attention.py:41-78):
    weights [H, nq, nk], mask [nq, nk] or [H, nq, nk] with 0 = masked (float, bool or integer)."""
    qp, ldq = _rows(q, "q"); kp, ldk = _rows(k, "k"); vp_, ldv = _rows(v, "v")
    nq, d = q.shape
    nk = k.shape[0]
    out = torch.empty((nq, d), device=q.device, dtype=torch.float32)
    wp, wstride, wcode = None, 0, 0
    if weights is not None:
        weights = _f32(weights, "attention_weights").contiguous()
        if tuple(weights.shape) != (n_heads, nq, nk):
            raise ValueError(f"attention_weights must be [H, nq, nk] = {(n_heads, nq, nk)}, got {tuple(weights.shape)}")
        wp, wstride, wcode = weights.data_ptr(), nq * nk, {"mul": 1, "add": 2}[way]
    mp, mstride = None, 0
    if mask is not None:
        mask = mask.to(torch.float32).contiguous()
        if tuple(mask.shape) == (n_heads, nq, nk) and n_heads > 1:
            mstride = nq * nk
        elif mask.numel() != nq * nk:
            raise ValueError(f"attention_mask must be [nq, nk] or [H, nq, nk], got {tuple(mask.shape)}")
        mp = mask.data_ptr()
    st = _call("vlsat_dense_attn_fwd", qp, ldq, kp, ldk, vp_, ldv, wp, wstride, wcode, mp, mstride, out.data_ptr(), d, nq, nk, n_heads,
               d // n_heads, _stream(), work=(4.0 * nq * nk * d, 4.0 * (2 * nq * d + 2 * nk * d + (n_heads + 1) * nq * nk)))
    _lib.check(st, "vlsat_dense_attn_fwd")
    return out


def flash_attn(q, k, v, n_heads: int, want_lse: bool = False):
    qp, ldq = _rows(q, "q"); kp, ldk = _rows(k, "k"); vp_, ldv = _rows(v, "v")
    nq, d = q.shape
    nk = k.shape[0]
    dk = d // n_heads
    out = torch.empty((nq, d), device=q.device, dtype=torch.float32)
    lse = torch.empty((n_heads, nq), device=q.device, dtype=torch.float32) if want_lse else None
    st = _call("vlsat_flash_attn_fwd", qp, ldq, kp, ldk, vp_, ldv, out.data_ptr(), d,
                                          lse.data_ptr() if want_lse else None, nq, nk, n_heads, dk, _stream(),
               work=(4.0 * nq * nk * d, 4.0 * (2 * nq * d + 2 * nk * d)))
    _lib.check(st, "vlsat_flash_attn_fwd")
    return (out, lse) if want_lse else out


def tf32_split(x: torch.Tensor):
    """(hi, lo) with hi = tf32-rounded x and lo = x - hi, both compact [rows, cols]; cols % 4 == 0."""
    xp, ldx = _rows(x, "x")
    r, c = x.shape
    hl = torch.empty((2, r, c), device=x.device, dtype=torch.float32)
    _lib.check(_call("vlsat_tf32_split", xp, ldx, r, c, hl[0].data_ptr(), hl[1].data_ptr(), _stream()), "vlsat_tf32_split")
    return hl[0], hl[1]


def flash_attn_tc(q, k, vt, nk: int, n_heads: int, want_lse: bool = False):
    """Tensor-core streaming attention. q [nq, D], k [nk, D], vt [D, >= nk] = transposed values with a
    row stride that is a multiple of 4 (column slices allowed); D = n_heads * 64. Each operand may also be
    given as its tf32 ``(hi, lo)`` split (as emitted by ``linear(..., emit_split=True)``)."""
    qh, ql = q if isinstance(q, tuple) else tf32_split(q)
    kh, kl = k if isinstance(k, tuple) else tf32_split(k)
    vh, vl = vt if isinstance(vt, tuple) else tf32_split(vt)
    vt = vh
    nq, d = qh.shape
    if d != n_heads * 64:
        raise ValueError("flash_attn_tc needs head size 64")
    out = torch.empty((nq, d), device=qh.device, dtype=torch.float32)
    lse = torch.empty((n_heads, nq), device=qh.device, dtype=torch.float32) if want_lse else None
    st = _call("vlsat_flash_attn_tc_fwd", qh.data_ptr(), ql.data_ptr(), d, kh.data_ptr(), kl.data_ptr(), d,
               vh.data_ptr(), vl.data_ptr(), vt.shape[1], out.data_ptr(), d, lse.data_ptr() if want_lse else None,
               nq, nk, n_heads, 64, _stream(), work=(4.0 * nq * nk * d, 4.0 * (2 * nq * d + 2 * nk * d)))
    _lib.check(st, "vlsat_flash_attn_tc_fwd")
    return (out, lse) if want_lse else out


ATTN_ENGINE = os.environ.get("VLSAT_ATTN_ENGINE", "bf16x3")      # 'bf16x3' (default) or 'tf32x3'


def bf16_split(x: torch.Tensor):
    """(hi, lo) bf16 pair of an fp32 matrix: hi = bf16(x), lo = bf16(x - hi); compact, row stride padded to 8."""
    xp, ldx = _rows(x, "x")
    r, c = x.shape
    ld = (c + 7) // 8 * 8
    hl = torch.empty((2, r, ld), device=x.device, dtype=torch.bfloat16)
    _lib.check(_call("vlsat_bf16_split", xp, ldx, r, c, hl[0].data_ptr(), hl[1].data_ptr(), ld, _stream()), "vlsat_bf16_split")
    return hl[0], hl[1]


def flash_attn_bf16(q, k, vt, nk: int, n_heads: int, want_lse: bool = False):
    """Tensor-core streaming attention, BF16x3 operands. q [nq, D], k [nk, D], vt [D, >= nk] fp32 (column-slice views
    allowed) or, each, the bf16 ``(hi, lo)`` pair a projection epilogue emitted; D = n_heads * 64."""
    qh, ql = q if isinstance(q, tuple) else bf16_split(q)
    kh, kl = k if isinstance(k, tuple) else bf16_split(k)
    vh, vl = vt if isinstance(vt, tuple) else bf16_split(vt[:, :nk])
    for t in (qh, kh, vh):
        if t.dtype != torch.bfloat16:
            raise TypeError("flash_attn_bf16: operand pairs must be bf16 pairs")
    nq, d = qh.shape
    if d != n_heads * 64:
        raise ValueError("flash_attn_bf16 needs head size 64")
    out = torch.empty((nq, d), device=qh.device, dtype=torch.float32)
    lse = torch.empty((n_heads, nq), device=qh.device, dtype=torch.float32) if want_lse else None
    ws_bytes = int(_lib.load().vlsat_flash_attn_bf16x3_workspace_bytes(nq, nk, n_heads))
    ws = torch.empty((ws_bytes // 4,), device=qh.device, dtype=torch.float32) if ws_bytes else None
    st = _call("vlsat_flash_attn_bf16x3_fwd", qh.data_ptr(), ql.data_ptr(), qh.stride(0), kh.data_ptr(), kl.data_ptr(), kh.stride(0),
               vh.data_ptr(), vl.data_ptr(), vh.stride(0), out.data_ptr(), d, lse.data_ptr() if want_lse else None,
               nq, nk, n_heads, 64, ws.data_ptr() if ws is not None else None, ws_bytes, _stream(), work=(4.0 * nq * nk * d, 4.0 * (2 * nq * d + 2 * nk * d)))
    _lib.check(st, "vlsat_flash_attn_bf16x3_fwd")
    return (out, lse) if want_lse else out


def bf16_split_t(x: torch.Tensor, want_t: bool = True):
    """bf16 (hi, lo) pair of an fp32 matrix [n, c] and, with ``want_t``, of its transpose [c, round8(n)] (tail columns zero):
    ((hi, lo), (hi_t, lo_t) or None). One pass over x (csrc/flash_attn_bwd.cu)."""
    xp, ldx = _rows(x, "x")
    n, c = x.shape
    ld = (c + 7) // 8 * 8
    hl = torch.empty((2, n, ld), device=x.device, dtype=torch.bfloat16)
    ldt = (n + 7) // 8 * 8
    hlt = torch.empty((2, c, ldt), device=x.device, dtype=torch.bfloat16) if want_t else None
    _lib.check(_call("vlsat_bf16_split_t", xp, ldx, n, c, hl[0].data_ptr(), hl[1].data_ptr(), ld,
                     hlt[0].data_ptr() if want_t else None, hlt[1].data_ptr() if want_t else None, ldt, _stream()), "vlsat_bf16_split_t")
    return (hl[0], hl[1]), ((hlt[0], hlt[1]) if want_t else None)


def flash_attn_bwd_stats(dout, out, lse, n_heads: int):
    """(lse2, delta), both [H, round64(nq)]: lse * log2(e) and rowsum_head(dout * out), padded with (1e30, 0)."""
    dp, lddo = _rows(dout, "dout"); op, ldo = _rows(out, "out")
    nq, d = dout.shape
    ld = (nq + 63) // 64 * 64
    st = torch.empty((2, n_heads, ld), device=dout.device, dtype=torch.float32)
    _lib.check(_call("vlsat_flash_attn_bwd_stats", dp, lddo, op, ldo, _f32(lse, "lse").data_ptr(), lse.stride(0), st[0].data_ptr(),
                     st[1].data_ptr(), ld, nq, n_heads, d // n_heads, _stream()), "vlsat_flash_attn_bwd_stats")
    return st[0], st[1]


def flash_attn_bf16_bwd(q, k, v, dout, out, lse, n_heads: int, prep=None):
    """Streaming tensor-core backward of A9 (csrc/flash_attn_bwd.cu): (dq, dk, dv) fp32 for q [nq, H*64], k / v [nk, H*64],
    the upstream gradient dout, the forward's output and log-sum-exp [H, nq]. Scores never reach HBM."""
    nq, d = q.shape
    nk = k.shape[0]
    if d != n_heads * 64:
        raise ValueError("flash_attn_bf16_bwd needs head size 64")
    dq = torch.empty((nq, d), device=q.device, dtype=torch.float32)
    dk = torch.empty((nk, d), device=q.device, dtype=torch.float32)
    dv = torch.empty((nk, d), device=q.device, dtype=torch.float32)
    if nq == 0 or nk == 0:
        return dq.zero_(), dk.zero_(), dv.zero_()
    if prep is not None:                     # (q, q^T, k, k^T, v) pairs the forward already made
        qp, qt, kp, kt, vp_ = prep
    else:
        qp, qt = bf16_split_t(q)
        kp, kt = bf16_split_t(k)
        vp_, _ = bf16_split_t(v, want_t=False)
    dop, dot = bf16_split_t(dout)
    lse2, delta = flash_attn_bwd_stats(dout, out, lse, n_heads)
    pair = lambda p: _lib.Bf16Pair(p[0].data_ptr(), p[1].data_ptr(), p[0].stride(0))
    o = _lib.FlashBwdOperands(pair(qp), pair(kp), pair(vp_), pair(dop), pair(qt), pair(kt), pair(dot),
                              lse2.data_ptr(), delta.data_ptr(), lse2.stride(0))
    ws_bytes = int(_lib.load().vlsat_flash_attn_bf16x3_bwd_workspace_bytes(nq, nk, n_heads))
    ws = torch.empty((ws_bytes // 4,), device=q.device, dtype=torch.float32) if ws_bytes else None
    # 7 products of 2 nq nk 64 flops per head (S and dP are computed in both launches)
    st = _call("vlsat_flash_attn_bf16x3_bwd", C.byref(o), dq.data_ptr(), d, dk.data_ptr(), d, dv.data_ptr(), d, nq, nk, n_heads, 64,
               ws.data_ptr() if ws is not None else None, ws_bytes, _stream(), work=(14.0 * nq * nk * d, 4.0 * (4 * nq * d + 4 * nk * d)))
    _lib.check(st, "vlsat_flash_attn_bf16x3_bwd")
    return dq, dk, dv


# -------------------------------------------------------------------------------------- graph attention
def build_csr(index_row: torch.Tensor, n_nodes: int):
    """Stable grouping of edges by ``index_row``: (row_ptr [N+1] int32, perm [E] int32)."""
    _i64(index_row, "index_row")
    e = index_row.numel()
    dev = index_row.device
    row_ptr = torch.empty((n_nodes + 1,), device=dev, dtype=torch.int32)
    perm = torch.empty((max(e, 1),), device=dev, dtype=torch.int32)
    ws = torch.empty((n_nodes + 1 + e,), device=dev, dtype=torch.int32)
    st = _call("vlsat_build_csr", index_row.data_ptr(), e, n_nodes, row_ptr.data_ptr(), perm.data_ptr(),
                                     ws.data_ptr(), ws.numel() * 4, _stream())
    _lib.check(st, "vlsat_build_csr")
    return row_ptr, perm[:e]


def gat_edge(q, v, k, edge_index, row_ptr, perm, c1, c1b, c2, c2b, n_heads: int, aggr: str = "max",
             use_edge: bool = True, want_prob: bool = False, want_argmax: bool = False,
             out: Optional[torch.Tensor] = None):
    """A8 core. q [N, H*d_n], v [N, H*d_o], k [E, H*d_e] (column-slice views allowed) -> xx [N, H*d_o]."""
    qp, ldq = _rows(q, "q"); vp_, ldv = _rows(v, "v")
    n = q.shape[0]
    _i64(edge_index, "edge_index")
    e = edge_index.shape[1]
    d_n, d_o = q.shape[1] // n_heads, v.shape[1] // n_heads
    kp, ldk, d_e = None, 0, 0
    if use_edge:
        kp, ldk = _rows(k, "k")
        d_e = k.shape[1] // n_heads
        if k.shape[0] != e:
            raise ValueError("gat_edge: k must have one row per edge")
    for nme, t in (("c1", c1), ("c1b", c1b), ("c2", c2), ("c2b", c2b)):
        _f32(t, nme)
        if not t.is_contiguous():
            raise ValueError(f"gat_edge: {nme} must be contiguous")
    hid = c1.shape[0]
    if c1.shape[1] != d_n + d_e or tuple(c2.shape) != (d_o, hid):
        raise ValueError(f"gat_edge: MLP weights {tuple(c1.shape)}, {tuple(c2.shape)} do not match d_n={d_n}, d_e={d_e}, d_o={d_o}")
    d_a = n_heads * d_o
    if out is None:
        out = torch.empty((n, d_a), device=q.device, dtype=torch.float32)
    xp, ldxx = _rows(out, "out")
    prob = torch.empty((e, d_o, n_heads), device=q.device, dtype=torch.float32) if want_prob else None
    arg = torch.empty((n, d_a), device=q.device, dtype=torch.int32) if want_argmax else None
    st = _call("vlsat_gat_edge_fwd", qp, ldq, vp_, ldv, kp, ldk, edge_index.data_ptr(), row_ptr.data_ptr(),
                                        perm.data_ptr(), c1.data_ptr(), c1b.data_ptr(), c2.data_ptr(), c2b.data_ptr(),
                                        n, e, n_heads, d_n, d_e, d_o, hid, AGGR[aggr], int(use_edge), xp, ldxx,
                                        prob.data_ptr() if want_prob else None, arg.data_ptr() if want_argmax else None,
                                        _stream(),
               # edge row + int64 index pair per edge; q, v read and xx written once per node
               work=(2.0 * e * n_heads * (hid * (d_n + d_e) + d_o * hid),
                     e * (n_heads * d_e * 4.0 + 16.0) + n * (n_heads * d_n + 2.0 * d_a) * 4.0))
    _lib.check(st, "vlsat_gat_edge_fwd")
    return out, prob, arg


def permute_rows(x: torch.Tensor, idx: torch.Tensor, gather: bool = True, emit_split: bool = False):
    """gather: out[i] = x[idx[i]];  scatter (gather=False): out[idx[i]] = x[i].  idx int32 permutation.
    ``emit_split=True``: also return the bf16 (hi, lo) pair of the result (or None when not applicable): ``(out, pair)``."""
    xp, ldx = _rows(x, "x")
    m, d = x.shape
    out = torch.empty((m, d), device=x.device, dtype=torch.float32)
    if idx.dtype != torch.int32 or not idx.is_cuda:
        raise TypeError("permute_rows: idx must be a CUDA int32 tensor")
    pair = None
    if emit_split and d % 8 == 0 and tensor_cores_enabled() and default_fmt(d) == FMT_BF16:
        pair = torch.empty((2, m, d), device=x.device, dtype=torch.bfloat16)
    _lib.check(_call("vlsat_permute_rows", xp, ldx, idx.data_ptr(), m, d, out.data_ptr(), d, int(gather),
                     pair[0].data_ptr() if pair is not None else None, pair[1].data_ptr() if pair is not None else None, _stream()),
               "vlsat_permute_rows")
    if emit_split:
        return out, ((pair[0], pair[1]) if pair is not None else None)
    return out


def permute_edges(edge_index: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    _i64(edge_index, "edge_index")
    e = edge_index.shape[1]
    out = torch.empty_like(edge_index)
    _lib.check(_call("vlsat_permute_edges", edge_index.data_ptr(), perm.data_ptr(), e, out.data_ptr(), _stream()),
               "vlsat_permute_edges")
    return out


def gat_tc_supported(n_heads: int, d_e: int, hid: int, d_o: int) -> bool:
    return (tensor_cores_enabled() and 128 % n_heads == 0 and d_e == 64 and hid % 32 == 0 and 32 <= hid <= 128
            and d_o in (32, 64))


_gat_workspaces = {}


def _gat_workspace(device, stream: int, words: int) -> torch.Tensor:
    """INT_MIN-filled scratch of the tensor-core edge kernel, kept per (device, stream, size): the kernel's finalize
    pass restores the fill, so only the first call pays for it (the C ABI itself owns nothing: it is told
    ``workspace_ready``)."""
    key = (str(device), stream, words)
    ws = _gat_workspaces.get(key)
    if ws is None:
        ws = _gat_workspaces[key] = torch.full((words,), -2 ** 31, device=device, dtype=torch.int32)
    return ws


def gat_edge_tc(k_hm, qc, v_hm, src, dst, c1k_split, c2_split, c2b, n_nodes: int, n_heads: int, out: torch.Tensor,
                want_prob: bool = False, d_n: int = 0):
    """Tensor-core A8 core (max aggregation). k_hm [E, H*d_e] head-major proj_edge output (CSR edge order) as fp32 or
    as the bf16 (hi, lo) pair its projection emitted, qc [N, H*hid] / v_hm [N, H*d_o] head-major node operands
    (column-slice views allowed), c1k / c2 bf16 pairs."""
    k_split = k_hm if isinstance(k_hm, tuple) else None
    if k_split is not None:
        k_hm = k_split[0]
    e = k_hm.shape[0]
    d_e = k_hm.shape[1] // n_heads
    hid = qc.shape[1] // n_heads
    d_o = v_hm.shape[1] // n_heads
    qp, ldq = _rows(qc, "qc"); vp_, ldv = _rows(v_hm, "v_hm")
    xp, ldxx = _rows(out, "out")
    kh = kl = None
    if e > 0:
        kh, kl = k_split if k_split is not None else bf16_split(k_hm)
        if kh.dtype != torch.bfloat16 or kh.stride(0) != n_heads * d_e:
            raise TypeError("gat_edge_tc: k must be a compact bf16 (hi, lo) pair")
    if c1k_split[0].dtype != torch.bfloat16 or c2_split[0].dtype != torch.bfloat16:
        raise TypeError("gat_edge_tc: c1k / c2 must be bf16 (hi, lo) pairs")
    prob = torch.empty((e, d_o, n_heads), device=out.device, dtype=torch.float32) if want_prob else None
    ws = _gat_workspace(out.device, _stream(), n_nodes * n_heads * d_o)
    st = _call("vlsat_gat_edge_tc_fwd", kh.data_ptr() if e else None, kl.data_ptr() if e else None, qp, ldq, vp_, ldv,
               src.data_ptr() if e else None, dst.data_ptr() if e else None,
               c1k_split[0].data_ptr(), c1k_split[1].data_ptr(), c2_split[0].data_ptr(), c2_split[1].data_ptr(),
               c2b.data_ptr(), n_nodes, e, n_heads, d_e, hid, d_o, xp, ldxx, prob.data_ptr() if want_prob else None,
               ws.data_ptr(), ws.numel() * 4, 1, _stream(),
               # same algorithmic work as vlsat_gat_edge_fwd, independent of the node-side folding
               # s = 4 bytes per element in the fp32-parity mode, 2 in the single-pass bf16 mode
               work=(2.0 * e * n_heads * (hid * ((d_n or d_e) + d_e) + d_o * hid),
                     e * (n_heads * d_e * _elem_bytes() + 16.0) + n_nodes * (n_heads * (d_n or d_e) + 2.0 * n_heads * d_o) * _elem_bytes()))
    _lib.check(st, "vlsat_gat_edge_tc_fwd")
    return out, prob


# ------------------------------------------------------------------ backward / training-mode primitives
def _round4(n: int) -> int:
    return (n + 3) // 4 * 4


def transpose(x: torch.Tensor, mult: int = 4) -> torch.Tensor:
    """x [R, C] (row stride allowed) -> [C, round_mult(R)] with the tail columns zero (GEMM operand for a reduction
    over R). For a 3-D contiguous x [B, R, C] returns [B, C, round4(R)]."""
    _f32(x, "x")
    if x.dim() == 2:
        xp, ldx = _rows(x, "x")
        r, c = x.shape
        out = torch.empty((c, (r + mult - 1) // mult * mult), device=x.device, dtype=torch.float32)
        _lib.check(_call("vlsat_transpose", xp, ldx, 0, out.data_ptr(), out.shape[1], 0, 1, r, c, _stream()), "vlsat_transpose")
        return out
    if x.dim() != 3 or not x.is_contiguous():
        raise ValueError("transpose: expected a 2-D matrix or a contiguous 3-D batch")
    b, r, c = x.shape
    out = torch.empty((b, c, _round4(r)), device=x.device, dtype=torch.float32)
    _lib.check(_call("vlsat_transpose", x.data_ptr(), c, r * c, out.data_ptr(), out.shape[2], c * out.shape[2], b, r, c, _stream()),
               "vlsat_transpose")
    return out


def act_bwd(dy: torch.Tensor, y: Optional[torch.Tensor], act: int, want_dz: bool = True, want_dbias: bool = True,
            scale: float = 1.0, scale_ptr: Optional[torch.Tensor] = None, emit_pair: bool = False, return_pair: bool = False):
    """(dz, dbias): dz = dy * act'(y) * scale * exp(scale_ptr), dbias = column sums of dz. ``emit_pair``: the same pass also
    writes the bf16 (hi, lo) pair of dz - the operand of the backward GEMMs that read it next - and remembers it on the
    returned tensor (on ``dy`` itself when the activation is the identity and no dz is written): ``act_pair`` finds it."""
    dp, lddy = _rows(dy, "dy")
    m, n = dy.shape
    yp, ldy = (None, 0) if y is None else _rows(y, "y")
    dz = torch.empty((m, n), device=dy.device, dtype=torch.float32) if want_dz else None
    db = zeros((n,), dy.device) if want_dbias else None
    vec = (n % 4 == 0 and lddy % 4 == 0 and (y is None or ldy % 4 == 0) and dp % 16 == 0 and (yp is None or yp % 16 == 0) and m > 0)
    if vec:
        pair = None
        if emit_pair and n % 8 == 0 and tensor_cores_enabled() and default_fmt(n) == FMT_BF16:
            pair = torch.empty((2, m, n), device=dy.device, dtype=torch.bfloat16)
        if dz is None and db is None and pair is None:
            return (dz, db, None) if return_pair else (dz, db)
        _lib.check(_call("vlsat_act_bwd_pair", dp, lddy, yp, ldy, act, scale, scale_ptr.data_ptr() if scale_ptr is not None else None,
                         dz.data_ptr() if want_dz else None, n, db.data_ptr() if want_dbias else None,
                         pair[0].data_ptr() if pair is not None else None, pair[1].data_ptr() if pair is not None else None, n, m, n,
                         _stream()), "vlsat_act_bwd_pair")
        if pair is not None:
            owner = dz if dz is not None else (dy if (act == ACT_NONE and scale == 1.0 and scale_ptr is None) else None)
            if owner is not None:
                try:
                    owner._vlsat_pair = (owner._version, (pair[0], pair[1]))
                except AttributeError:
                    pass
        if return_pair:
            return dz, db, ((pair[0], pair[1]) if pair is not None else None)
        return dz, db
    if dz is None and db is None:
        return (dz, db, None) if return_pair else (dz, db)
    _lib.check(_call("vlsat_act_bwd", dp, lddy, yp, ldy, act, scale, scale_ptr.data_ptr() if scale_ptr is not None else None,
                     dz.data_ptr() if want_dz else None, n, db.data_ptr() if want_dbias else None, m, n, _stream()), "vlsat_act_bwd")
    return (dz, db, None) if return_pair else (dz, db)


def scatter_add_rows(x: torch.Tensor, idx: torch.Tensor, out: torch.Tensor, rows_per_idx: int = 1) -> torch.Tensor:
    """out[idx[i // R] * R + i % R, :] += x[i, :] (out is accumulated into)."""
    xp, ldx = _rows(x, "x"); op, ldo = _rows(out, "out")
    _i64(idx, "idx")
    if x.shape[1] != out.shape[1] or idx.numel() * rows_per_idx != x.shape[0]:
        raise ValueError("scatter_add_rows: shape mismatch")
    _lib.check(_call("vlsat_scatter_add_rows", xp, ldx, idx.data_ptr(), rows_per_idx, x.shape[0], x.shape[1], op, ldo, _stream()),
               "vlsat_scatter_add_rows")
    return out


def gather_rows(x: torch.Tensor, idx: torch.Tensor, rows_per_idx: int = 1) -> torch.Tensor:
    xp, ldx = _rows(x, "x")
    _i64(idx, "idx")
    rows = idx.numel() * rows_per_idx
    out = torch.empty((rows, x.shape[1]), device=x.device, dtype=torch.float32)
    _lib.check(_call("vlsat_gather_rows", xp, ldx, idx.data_ptr(), rows_per_idx, rows, x.shape[1], out.data_ptr(), x.shape[1], _stream()),
               "vlsat_gather_rows")
    return out


def add_layernorm_bwd(dy, x, res, gamma, beta, eps: float, relu: bool):
    dp, lddy = _rows(dy, "dy"); xp, ldx = _rows(x, "x")
    rp, ldr = (None, 0) if res is None else _rows(res, "res")
    m, d = x.shape
    dx = torch.empty((m, d), device=x.device, dtype=torch.float32)
    dg = zeros((d,), x.device)
    db = zeros((d,), x.device)
    _lib.check(_call("vlsat_add_layernorm_bwd", dp, lddy, xp, ldx, rp, ldr, gamma.data_ptr(), beta.data_ptr(), dx.data_ptr(), d,
                     dg.data_ptr(), db.data_ptr(), m, d, eps, int(relu), _stream()), "vlsat_add_layernorm_bwd")
    return dx, dg, db


def gat_softmax_aggr(t, v_hm, dst, row_ptr, n_nodes: int, n_heads: int, aggr: str):
    """(xx [N, H*d_o] interleaved, prob [E*H, d_o], argmax [N, H*d_o] int32 or None)."""
    _f32(t, "t")
    vp_, ldv = _rows(v_hm, "v_hm")
    d_o = v_hm.shape[1] // n_heads
    e = t.shape[0] // n_heads
    if not t.is_contiguous() or t.shape[1] != d_o:
        raise ValueError("gat_softmax_aggr: t must be contiguous [E*H, d_o]")
    xx = torch.empty((n_nodes, n_heads * d_o), device=t.device, dtype=torch.float32)
    prob = torch.empty_like(t)
    arg = torch.empty((n_nodes, n_heads * d_o), device=t.device, dtype=torch.int32) if aggr == "max" else None
    _lib.check(_call("vlsat_gat_softmax_aggr_fwd", t.data_ptr() if e else None, vp_, ldv, dst.data_ptr() if e else None,
                     row_ptr.data_ptr(), n_nodes, e, n_heads, d_o, AGGR[aggr], xx.data_ptr(), xx.shape[1], prob.data_ptr() if e else None,
                     arg.data_ptr() if arg is not None else None, _stream()), "vlsat_gat_softmax_aggr_fwd")
    return xx, prob, arg


def gat_softmax_aggr_bwd(dxx, prob, v_hm, dst, row_ptr, arg, n_nodes: int, n_heads: int, aggr: str):
    """(dt [E*H, d_o], dv_hm [N, H*d_o])."""
    dp, lddxx = _rows(dxx, "dxx"); vp_, ldv = _rows(v_hm, "v_hm")
    d_o = v_hm.shape[1] // n_heads
    e = prob.shape[0] // n_heads
    dt = torch.empty_like(prob)
    dv = zeros((n_nodes, n_heads * d_o), dxx.device)
    _lib.check(_call("vlsat_gat_softmax_aggr_bwd", dp, lddxx, prob.data_ptr() if e else None, vp_, ldv, dst.data_ptr() if e else None,
                     row_ptr.data_ptr(), arg.data_ptr() if arg is not None else None, n_nodes, e, n_heads, d_o, AGGR[aggr],
                     dt.data_ptr() if e else None, dv.data_ptr(), dv.shape[1], _stream()), "vlsat_gat_softmax_aggr_bwd")
    return dt, dv


def attn_prob_bwd(s, dp, lse, delta, scale: float):
    """In place on dp: dS. Returns (dS [nq, nk], dS^T [nk, round4(nq)], P^T [nk, round4(nq)])."""
    nq, nk = s.shape
    if not (s.is_contiguous() and dp.is_contiguous()):
        raise ValueError("attn_prob_bwd: s and dp must be contiguous")
    ldt = _round4(nq)
    ds_t = torch.empty((nk, ldt), device=s.device, dtype=torch.float32)
    p_t = torch.empty((nk, ldt), device=s.device, dtype=torch.float32)
    _lib.check(_call("vlsat_attn_prob_bwd", s.data_ptr(), dp.data_ptr(), nk, lse.data_ptr(), delta.data_ptr(), scale, dp.data_ptr(),
                     ds_t.data_ptr(), p_t.data_ptr(), ldt, nq, nk, _stream()), "vlsat_attn_prob_bwd")
    return dp, ds_t, p_t


def attn_prob_bwd_pairs(s, dp, lse, delta, scale: float):
    """bf16 (hi, lo) pairs of dS [nq, nk], dS^T [nk, round8(nq)] and P^T [nk, round8(nq)] (no fp32 outputs); nk % 8 == 0."""
    nq, nk = s.shape
    if not (s.is_contiguous() and dp.is_contiguous()) or nk % 8:
        raise ValueError("attn_prob_bwd_pairs: s and dp must be contiguous with nk % 8 == 0")
    ldt = (nq + 7) // 8 * 8
    ds = torch.empty((2, nq, nk), device=s.device, dtype=torch.bfloat16)
    ds_t = torch.empty((2, nk, ldt), device=s.device, dtype=torch.bfloat16)
    p_t = torch.empty((2, nk, ldt), device=s.device, dtype=torch.bfloat16)
    _lib.check(_call("vlsat_attn_prob_bwd_pairs", s.data_ptr(), dp.data_ptr(), nk, lse.data_ptr(), delta.data_ptr(), scale,
                     ds[0].data_ptr(), ds[1].data_ptr(), nk, ds_t[0].data_ptr(), ds_t[1].data_ptr(), p_t[0].data_ptr(), p_t[1].data_ptr(),
                     ldt, nq, nk, _stream()), "vlsat_attn_prob_bwd_pairs")
    return (ds[0], ds[1]), (ds_t[0], ds_t[1]), (p_t[0], p_t[1])


def rowdot_heads(a, b, n_heads: int) -> torch.Tensor:
    ap, lda = _rows(a, "a"); bp, ldb = _rows(b, "b")
    m, d = a.shape
    out = torch.empty((n_heads, m), device=a.device, dtype=torch.float32)
    _lib.check(_call("vlsat_rowdot_heads", ap, lda, bp, ldb, out.data_ptr(), m, n_heads, d // n_heads, _stream()), "vlsat_rowdot_heads")
    return out


def pair_features(centres, seg_start, seg_end, pair_off, n_pairs: int) -> torch.Tensor:
    cp, ldc = _rows(centres, "centres")
    out = torch.empty((n_pairs, 4), device=centres.device, dtype=torch.float32)
    _lib.check(_call("vlsat_pair_features", cp, ldc, seg_start.data_ptr(), seg_end.data_ptr(), _i64(pair_off, "pair_off").data_ptr(),
                     centres.shape[0], out.data_ptr(), _stream()), "vlsat_pair_features")
    return out


def node_attn_bias(q, k, v, bias, pair_off, seg_start, seg_end, n_heads: int, max_scene: int) -> torch.Tensor:
    qp, ldq = _rows(q, "q"); kp, ldk = _rows(k, "k"); vp_, ldv = _rows(v, "v")
    n, d = q.shape
    out = torch.empty((n, d), device=q.device, dtype=torch.float32)
    _lib.check(_call("vlsat_node_attn_bias_fwd", qp, ldq, kp, ldk, vp_, ldv, _f32(bias, "bias").data_ptr(), pair_off.data_ptr(),
                     seg_start.data_ptr(), seg_end.data_ptr(), n_heads, d // n_heads, max_scene, out.data_ptr(), d, n, _stream()),
               "vlsat_node_attn_bias_fwd")
    return out


def node_attn_bias_bwd(q, k, v, bias, pair_off, seg_start, seg_end, dout, n_heads: int, max_scene: int):
    qp, ldq = _rows(q, "q"); kp, ldk = _rows(k, "k"); vp_, ldv = _rows(v, "v"); dp, lddo = _rows(dout, "dout")
    n, d = q.shape
    dq = torch.empty((n, d), device=q.device, dtype=torch.float32)
    dk = zeros((k.shape[0], d), q.device)
    dv = zeros((v.shape[0], d), q.device)
    dbias = torch.empty_like(bias)
    _lib.check(_call("vlsat_node_attn_bias_bwd", qp, ldq, kp, ldk, vp_, ldv, bias.data_ptr(), pair_off.data_ptr(), seg_start.data_ptr(),
                     seg_end.data_ptr(), dp, lddo, n_heads, d // n_heads, max_scene, dq.data_ptr(), d, dk.data_ptr(), d, dv.data_ptr(), d,
                     dbias.data_ptr(), n, _stream()), "vlsat_node_attn_bias_bwd")
    return dq, dk, dv, dbias


def pointnet_pool_bwd(dz3, arg, h2, w3, n_pts: int):
    """(dW3 [c_out, c2], dh2 [n_obj*n_pts, c2])."""
    n_obj, c_out = dz3.shape
    c2 = w3.shape[1]
    dw3 = zeros((c_out, c2), dz3.device)
    dh2 = zeros((n_obj * n_pts, c2), dz3.device)
    if not (dz3.is_contiguous() and h2.is_contiguous() and w3.is_contiguous() and arg.is_contiguous()):
        raise ValueError("pointnet_pool_bwd: operands must be contiguous")
    _lib.check(_call("vlsat_pointnet_pool_bwd", dz3.data_ptr(), arg.data_ptr(), h2.data_ptr(), w3.data_ptr(), n_obj, n_pts, c_out, c2,
                     dw3.data_ptr(), dh2.data_ptr(), _stream()), "vlsat_pointnet_pool_bwd")
    return dw3, dh2


def dropout(x: torch.Tensor, p: float, seed: int, offset: int, device_step: Optional[torch.Tensor] = None,
            emit_pair: bool = False) -> torch.Tensor:
    """Counter-based dropout (csrc/backward.cu). ``emit_pair``: the result feeds a projection - the same pass also writes its
    bf16 (hi, lo) pair and remembers it on the returned tensor (``act_pair`` finds it); ignored where pairs do not apply."""
    xp, ldx = _rows(x, "x")
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    if device_step is not None and (device_step.dtype != torch.int64 or not device_step.is_cuda):
        raise TypeError("dropout: device_step must be a CUDA int64 scalar tensor")
    step_p = device_step.data_ptr() if device_step is not None else None
    r, c = x.shape
    if (emit_pair and r > 0 and c % 8 == 0 and c >= 32 and ldx % 4 == 0 and xp % 16 == 0 and tensor_cores_enabled() and default_fmt(c) == FMT_BF16):
        pair = torch.empty((2, r, c), device=x.device, dtype=torch.bfloat16)
        _lib.check(_call("vlsat_dropout_pair", xp, ldx, out.data_ptr(), c, r, c, p, seed, offset, step_p, pair[0].data_ptr(), pair[1].data_ptr(),
                         c, _stream()), "vlsat_dropout_pair")
        out._vlsat_pair = (out._version, (pair[0], pair[1]))
        return out
    _lib.check(_call("vlsat_dropout", xp, ldx, out.data_ptr(), c, r, c, p, seed, offset, step_p, _stream()), "vlsat_dropout")
    return out


def batchnorm(x, gamma, beta, mean, rstd, running_mean, running_var, momentum: float, eps: float, batch_stats: bool, relu: bool):
    xp, ldx = _rows(x, "x")
    m, n = x.shape
    y = torch.empty((m, n), device=x.device, dtype=torch.float32)
    _lib.check(_call("vlsat_batchnorm_fwd", xp, ldx, gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                     running_mean.data_ptr() if running_mean is not None else None,
                     running_var.data_ptr() if running_var is not None else None, momentum, eps, int(batch_stats), int(relu),
                     y.data_ptr(), n, m, n, _stream()), "vlsat_batchnorm_fwd")
    return y


def batchnorm_bwd(dy, x, mean, rstd, gamma, beta, relu: bool, batch_stats: bool):
    dp, lddy = _rows(dy, "dy"); xp, ldx = _rows(x, "x")
    m, n = x.shape
    dx = torch.empty((m, n), device=x.device, dtype=torch.float32)
    dg = torch.empty((n,), device=x.device, dtype=torch.float32)
    db = torch.empty((n,), device=x.device, dtype=torch.float32)
    _lib.check(_call("vlsat_batchnorm_bwd", dp, lddy, xp, ldx, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                     int(relu), int(batch_stats), dx.data_ptr(), n, dg.data_ptr(), db.data_ptr(), m, n, _stream()), "vlsat_batchnorm_bwd")
    return dx, dg, db


def row_l2norm_bwd(dy, x):
    if not (dy.is_contiguous() and x.is_contiguous()):
        raise ValueError("row_l2norm_bwd: operands must be contiguous")
    dx = torch.empty_like(x)
    _lib.check(_call("vlsat_row_l2norm_bwd", dy.data_ptr(), x.data_ptr(), dx.data_ptr(), x.shape[0], x.shape[1], _stream()),
               "vlsat_row_l2norm_bwd")
    return dx


def dot_accum(a, b, out) -> None:
    _lib.check(_call("vlsat_dot_accum", a.data_ptr(), b.data_ptr(), a.numel(), out.data_ptr(), _stream()), "vlsat_dot_accum")


def wgrad_small(dz: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """dW [N, K] = dz^T x for N, K <= 128 and a tall row count (exact FP32, slab-parallel with atomic accumulation)."""
    dp, lddz = _rows(dz, "dz"); xp, ldx = _rows(x, "x")
    m, n = dz.shape
    k = x.shape[1]
    dw = zeros((n, k), dz.device)
    _lib.check(_call("vlsat_wgrad_small", dp, lddz, xp, ldx, m, n, k, dw.data_ptr(), k, _stream(), work=(2.0 * m * n * k, 4.0 * m * (n + k))),
               "vlsat_wgrad_small")
    return dw
