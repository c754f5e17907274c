"""Differentiable operators of the path: ``torch.autograd.Function`` wrappers whose forward AND backward are
vlsat_b200 kernels (C ABI). PyTorch's autograd engine only chains them (and sums gradients of tensors used twice);
no torch compute kernel evaluates any part of the model.

Backward of every dense projection = two more projections on the same GEMM engine:
    dZ = dY * act'(Y)            (vlsat_act_bwd, also emits dbias = column sums)
    dX = dZ W                    (vlsat_linear_fwd on W^T)
    dW = dZ^T X                  (vlsat_linear_fwd on dZ^T, X^T; vlsat_transpose zero-pads the reduction length)
Row gathers in the forward epilogue become atomic row scatter-adds (vlsat_scatter_add_rows).

Reference: autograd of network_MMG.py / network_PointNet.py / attention.py / SGFN_MMG/model.py:288-335 (the reference
has no hand-written backward).
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
from torch.autograd import Function

from . import ops

ACT_NONE, ACT_RELU, ACT_SIGMOID = ops.ACT_NONE, ops.ACT_RELU, ops.ACT_SIGMOID


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Incoming gradient with contiguous rows (autograd may hand over expanded or strided tensors)."""
    if t is None:
        return None
    if t.dim() == 2 and (t.shape[1] == 1 or t.stride(1) == 1) and t.stride(0) >= t.shape[1]:
        return t
    return t.contiguous()


def _pairs_engine(*dims) -> bool:
    """The stored-operand backward GEMMs (vlsat_gemm_pairs, BF16x3) serve this projection: tensor-core engine with bf16
    pairs and every stored row length a multiple of 8."""
    # widths below 32 stay on the exact FFMA path: they are tiny, and some of their gradients vanish in exact arithmetic
    # (everything upstream of the distance bias sums to zero over a softmax row), where BF16x3 noise is all there is
    return ops.tensor_cores_enabled() and all(d % 8 == 0 and d >= 32 for d in dims) and ops.default_fmt(8) == ops.FMT_BF16


def weight_grad(dz: torch.Tensor, x: torch.Tensor, dz_pair=None, x_pair=None) -> torch.Tensor:
    """dW [N, K] = dz^T x. Tensor-core engine: both operands are read as stored (MN-major descriptors, the reduction over
    the rows split across CTAs - csrc/gemm_tc.cu); otherwise a slab-parallel FP32 kernel for small [N, K], or the GEMM
    engine on transposed copies."""
    n, k = dz.shape[1], x.shape[1]
    if dz.shape[0] == 0:
        return torch.zeros((n, k), device=dz.device, dtype=torch.float32)
    if _pairs_engine(n, k):
        return ops.gemm_tn(dz_pair if dz_pair is not None else ops.act_pair(dz), x_pair if x_pair is not None else ops.act_pair(x), n, k)
    if n <= 128 and k <= 128:
        return ops.wgrad_small(dz, x)
    return ops.linear(ops.transpose(dz), ops.transpose(x), cache_w=False)


# ------------------------------------------------------------------------------------------- projections
class _Linear(Function):
    @staticmethod
    def forward(ctx, x, w, bias, act, ga, ia, gb, ib, residual, scale, emit_pair=False):
        if (residual is not None or scale is not None) and act != ACT_NONE:
            raise ValueError("linear: residual / logit scale cannot be combined with an activation")
        gather = (ga, ia, gb, ib) if ga is not None else None
        n, k = w.shape
        # tensor-core engine: the bf16 pairs of x and w feed the forward GEMM and, kept for the backward, dW = dZ^T X and
        # dX = dZ W without a transposed copy or a second split pass
        pairs = _pairs_engine(n, k) and x.shape[0] > 0 and k >= 32
        xp = ops.act_pair(x) if pairs else None
        wp = ops.weight_pair(w) if pairs else None
        # emit_pair: the consumer of y is another projection - its bf16 pair leaves through this GEMM's epilogue
        emit = bool(emit_pair) and pairs and n >= 32
        y = ops.linear(x, w, bias, act=act, gather=gather, residual=residual, scale_ptr=scale, x_split=xp, w_split=wp,
                       emit_split="bf16" if emit else False)
        if emit:
            y, ypair = y
            y._vlsat_pair = (y._version, ypair)
        ctx.act = act
        ctx.has = (bias is not None, ga is not None, gb is not None, residual is not None, scale is not None)
        ctx.shapes = (ga.shape if ga is not None else None, gb.shape if gb is not None else None)
        ctx.pairs = (xp, wp)
        keep_y = act != ACT_NONE or scale is not None
        ctx.save_for_backward(x, w, y if keep_y else None, ia, ib, scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y, ia, ib, scale = ctx.saved_tensors
        has_b, has_ga, has_gb, has_res, has_scale = ctx.has
        need = ctx.needs_input_grad
        dy = _c(dy)
        n = dy.shape[1]
        d_scale = None
        if has_scale and need[9]:
            d_scale = ops.zeros(tuple(scale.shape), scale.device)
            ops.dot_accum(dy, y, d_scale)                       # y = e^s z  =>  dy/ds = y
        d_res = dy if (has_res and need[8]) else None
        plain = ctx.act == ACT_NONE and not has_scale
        want_db = has_b and need[2]
        xp, wp = ctx.pairs
        pairs = xp is not None and dy.shape[0] > 0              # dz feeds the stored-operand GEMMs: its pair comes out of this pass
        dzp = None
        if plain:
            dz = dy
            if want_db or pairs:
                _, db, dzp = ops.act_bwd(dy, None, ACT_NONE, want_dz=False, want_dbias=want_db, emit_pair=pairs, return_pair=True)
            else:
                db = None
        else:
            # the fp32 dz is only read by the scatter-adds of gathered operands; the two GEMMs read its bf16 pair
            need_dz = not pairs or (has_ga and need[4]) or (has_gb and need[6])
            dz, db, dzp = ops.act_bwd(dy, y, ctx.act, want_dz=need_dz, want_dbias=want_db, scale_ptr=scale if has_scale else None,
                                      emit_pair=pairs, return_pair=True)
            if dzp is None and dz is None:                      # pair emission declined (shape): fall back to the fp32 dz
                dz, db2 = ops.act_bwd(dy, y, ctx.act, want_dz=True, want_dbias=False, scale_ptr=scale if has_scale else None)
        dx = dw = d_ga = d_gb = None
        if xp is not None and dy.shape[0] > 0:
            if dzp is None:
                dzp = ops.act_pair(dz)
            if need[0] and need[1]:
                # the two gradients are independent: dW on the side stream next to dX (ops.fork_join; dzp, xp, wp stay
                # referenced by this frame until the join). Most of these GEMMs are one tile-latency long, so two at a
                # time nearly halves their share of the step.
                dw, dx = ops.fork_join(lambda: ops.gemm_tn(dzp, xp, n, w.shape[1]), lambda: ops.gemm_nn(dzp, wp, w.shape[1]), dy.device)
            elif need[0]:
                dx = ops.gemm_nn(dzp, wp, w.shape[1])           # dZ [M, N] . W [N, K], W as the forward stores it
            elif need[1]:
                dw = ops.gemm_tn(dzp, xp, n, w.shape[1])        # dZ^T X, both as stored
        else:
            if need[0]:
                wt = ops.transpose(w)                           # [K, round4(N)]
                dx = ops.linear(dz, wt[:, :n], cache_w=False)   # (an empty batch yields an empty dx)
            if need[1]:
                dw = weight_grad(dz, x)                         # [N, K], reduction over the rows
        if has_ga and need[4]:
            d_ga = ops.scatter_add_rows(dz, ia, ops.zeros(tuple(ctx.shapes[0]), dz.device))
        if has_gb and need[6]:
            d_gb = ops.scatter_add_rows(dz, ib, ops.zeros(tuple(ctx.shapes[1]), dz.device))
        return dx, dw, db, None, d_ga, None, d_gb, None, d_res, d_scale, None


def linear(x, w, bias=None, act=ACT_NONE, gather=None, residual=None, scale=None, emit_pair=False):
    """y = act(x w^T + bias + ga[ia] + gb[ib]) (+ residual) (* exp(scale)); all tensor arguments differentiable.
    ``emit_pair``: y feeds another projection directly - write its bf16 (hi, lo) pair from this GEMM's epilogue."""
    ga, ia, gb, ib = gather if gather is not None else (None, None, None, None)
    return _Linear.apply(x, w, bias, act, ga, ia, gb, ib, residual, scale, emit_pair)


class _Relu(Function):
    @staticmethod
    def forward(ctx, x, emit_pair=True):
        if emit_pair:
            y, pair = ops.relu(x.contiguous(), emit_split=True)           # the next layer's projections read the pair
        else:
            y, pair = ops.relu(x.contiguous()), None
        if pair is not None:
            y._vlsat_pair = (y._version, pair)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _c(dy)
        return ops.act_bwd(dy.view(-1, y.shape[-1]), y.view(-1, y.shape[-1]), ACT_RELU, want_dbias=False)[0].view(y.shape), None


def relu(x, emit_pair: bool = True):
    """``emit_pair=False``: something other than a projection reads the result (a dropout pass that emits the pair itself)."""
    return _Relu.apply(x, emit_pair)


class _AddLayerNorm(Function):
    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps, relu_out):
        y = ops.add_layernorm(x, res, gamma, beta, eps=eps, relu=relu_out, emit_split=True)
        if isinstance(y, tuple):                                # (y, pair or None): q / k / v and node projections read the pair
            y, pair = y
            if pair is not None:
                y._vlsat_pair = (y._version, pair)
        ctx.save_for_backward(x, res, gamma, beta)
        ctx.eps, ctx.relu = eps, relu_out
        return y

    @staticmethod
    def backward(ctx, dy):
        x, res, gamma, beta = ctx.saved_tensors
        dx, dg, db = ops.add_layernorm_bwd(_c(dy), x, res, gamma, beta, ctx.eps, ctx.relu)
        return dx, (dx if res is not None else None), dg, db, None, None


def add_layernorm(x, res, gamma, beta, eps=1e-5, relu_out=False):
    return _AddLayerNorm.apply(x, res, gamma, beta, eps, relu_out)


# -------------------------------------------------------------------------------------------- dropout
class DropoutState:
    """Counter-based dropout stream: every call consumes ``numel`` counters of (seed, offset). ``device_step`` (set by
    ``GraphedTrainStep``) is an int64 counter in device memory mixed into the seed at run time, so that replays of a
    captured training step draw fresh masks although seed / offset are baked into the graph."""
    seed = 0
    offset = 0
    device_step = None

    @classmethod
    def manual_seed(cls, seed: int) -> None:
        cls.seed, cls.offset = int(seed) & 0x7FFFFFFFFFFFFFFF, 0

    @classmethod
    def take(cls, numel: int):
        s, o = cls.seed, cls.offset
        cls.offset += int(numel)
        return s, o


class _Dropout(Function):
    @staticmethod
    def forward(ctx, x, p, emit_pair=False):
        ctx.p = p
        ctx.key = DropoutState.take(x.numel())
        ctx.step = DropoutState.device_step
        return ops.dropout(x, p, *ctx.key, device_step=ctx.step, emit_pair=bool(emit_pair) and x.dim() == 2)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout(_c(dy), ctx.p, *ctx.key, device_step=ctx.step), None, None


def dropout(x, p: float, training: bool, emit_pair: bool = False):
    """``emit_pair``: x feeds a projection next - its bf16 pair leaves with the dropout pass."""
    if not training or p <= 0.0:
        return x
    return _Dropout.apply(x, float(p), emit_pair)


# ------------------------------------------------------------------------------------------ weight column blocks
class _SplitCols(Function):
    """Column blocks of a weight [N, K] as views (``w[:, a:b]``), with ONE backward: the blocks' gradients concatenated.
    Plain slicing gives every block its own SliceBackward - a zero fill of the whole [N, K], a copy into the slice and an add
    per extra block (61 launches per training step for the split nn_edge / attention-MLP / pair-projector weights, 9 now)."""

    @staticmethod
    def forward(ctx, w, *sizes):
        ctx.sizes, ctx.rows = sizes, w.shape[0]
        return tuple(w.split(list(sizes), dim=1))

    @staticmethod
    def backward(ctx, *grads):
        ref = next(g for g in grads if g is not None)
        parts = [g if g is not None else torch.zeros((ctx.rows, n), device=ref.device, dtype=ref.dtype) for g, n in zip(grads, ctx.sizes)]
        return (torch.cat(parts, 1),) + (None,) * len(ctx.sizes)


def split_cols(w, sizes):
    """``w[:, :s0], w[:, s0:s0+s1], ...`` (views of ``w``; sizes must add up to its width) with a single fused backward."""
    sizes = tuple(int(n) for n in sizes)
    if sum(sizes) != w.shape[1] or any(n <= 0 for n in sizes):
        raise ValueError(f"split_cols: sizes {sizes} do not tile the {w.shape[1]} columns")
    if not (torch.is_grad_enabled() and w.requires_grad):
        return tuple(w.split(list(sizes), dim=1))
    return _SplitCols.apply(w, *sizes)


# ------------------------------------------------------------------------------------------ batch norm
class _BatchNorm(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, batch_stats, relu_out):
        n = x.shape[1]
        if batch_stats:
            mean = torch.empty((n,), device=x.device, dtype=torch.float32)
            rstd = torch.empty((n,), device=x.device, dtype=torch.float32)
        else:
            mean = running_mean
            rstd = torch.rsqrt(running_var + eps)                  # [n] parameter-sized bookkeeping
        y = ops.batchnorm(x, gamma, beta, mean, rstd, running_mean if batch_stats else None,
                          running_var if batch_stats else None, momentum, eps, batch_stats, relu_out)
        ctx.save_for_backward(x, gamma, beta, mean, rstd)
        ctx.flags = (batch_stats, relu_out)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, mean, rstd = ctx.saved_tensors
        batch_stats, relu_out = ctx.flags
        dx, dg, db = ops.batchnorm_bwd(_c(dy), x, mean, rstd, gamma, beta, relu_out, batch_stats)
        return dx, dg, db, None, None, None, None, None, None


def batchnorm(x, bn: torch.nn.BatchNorm1d, training: bool, relu_out: bool = False):
    """nn.BatchNorm1d (+ ReLU): batch statistics and running-stat update in training mode, running stats in eval."""
    if training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    momentum = 0.1 if bn.momentum is None else bn.momentum
    return _BatchNorm.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, momentum, bn.eps, bool(training), relu_out)


# --------------------------------------------------------------------------------------- small row ops
class _RowL2Norm(Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.row_l2norm(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.row_l2norm_bwd(dy.contiguous(), x)


def row_l2norm(x):
    return _RowL2Norm.apply(x)


class _PermuteRows(Function):
    @staticmethod
    def forward(ctx, x, perm, gather, emit_pair=False):
        ctx.perm, ctx.gather = perm, gather
        if emit_pair:
            y, pair = ops.permute_rows(x, perm, gather=gather, emit_split=True)
            if pair is not None:
                y._vlsat_pair = (y._version, pair)
            return y
        return ops.permute_rows(x, perm, gather=gather)

    @staticmethod
    def backward(ctx, dy):
        return ops.permute_rows(_c(dy), ctx.perm, gather=not ctx.gather), None, None, None


def permute_rows(x, perm, gather: bool, emit_pair: bool = False):
    """``emit_pair``: the permuted rows feed projections - their bf16 pair leaves with the permutation pass."""
    if x.shape[0] == 0:
        return x
    return _PermuteRows.apply(x, perm, gather, emit_pair)


class _GatherRows(Function):
    @staticmethod
    def forward(ctx, x, idx):
        ctx.idx, ctx.shape = idx, x.shape
        return ops.gather_rows(x, idx)

    @staticmethod
    def backward(ctx, dy):
        out = ops.zeros(tuple(ctx.shape), dy.device)
        return ops.scatter_add_rows(_c(dy), ctx.idx, out), None


def gather_rows(x, idx):
    """x[idx] for int64 row indices (Gen_Index, network_util.py:50-62); backward = atomic row scatter-add."""
    return _GatherRows.apply(x, idx)


# ---------------------------------------------------------------------------------------- PointNet (A1)
class _PointNet(Function):
    """Fused forward (max-pool with arg-max); backward recomputes the per-point activations h1, h2 with two
    projections, routes dOut through the arg-max (vlsat_pointnet_pool_bwd) and finishes with projection backwards."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3):
        out, arg = ops.pointnet(x, w1, b1, w2, b2, w3, b3, want_argmax=True)
        ctx.save_for_backward(x, w1, b1, w2, b2, w3, out, arg)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w1, b1, w2, b2, w3, out, arg = ctx.saved_tensors
        n_obj, c_in, n_pts = x.shape
        xt = ops.transpose(x)                                       # [n_obj, c_in, P] -> [n_obj, P, round4(c_in)]
        ldx = xt.shape[2]
        xr = xt.view(n_obj * n_pts, ldx)[:, :c_in]                  # rows = points
        pairs = _pairs_engine(w2.shape[0], w2.shape[1])
        h1p = dz2p = None
        if pairs:                                                   # h1's pair feeds the second projection and dW2 = dZ2^T H1
            h1, h1p = ops.linear(xr, w1, b1, act=ACT_RELU, emit_split="bf16")
            h2 = ops.linear(h1, w2, b2, act=ACT_RELU, x_split=h1p)
        else:
            h1 = ops.linear(xr, w1, b1, act=ACT_RELU)
            h2 = ops.linear(h1, w2, b2, act=ACT_RELU)
        dz3, db3 = ops.act_bwd(_c(dout), out, ACT_RELU)            # max of ReLU outputs: zero where the pooled value is 0
        dw3, dh2 = ops.pointnet_pool_bwd(dz3, arg, h2, w3, n_pts)
        if pairs:
            dz2, db2, dz2p = ops.act_bwd(dh2, h2, ACT_RELU, emit_pair=True, return_pair=True)
        else:
            dz2, db2 = ops.act_bwd(dh2, h2, ACT_RELU)
        dw2 = weight_grad(dz2, h1, dz_pair=dz2p, x_pair=h1p)
        if pairs:
            dh1 = ops.gemm_nn(dz2p if dz2p is not None else ops.act_pair(dz2), ops.weight_pair(w2), w2.shape[1])
        else:
            dh1 = ops.linear(dz2, ops.transpose(w2)[:, :w2.shape[0]], cache_w=False)
        dz1, db1 = ops.act_bwd(dh1, h1, ACT_RELU)
        dw1 = weight_grad(dz1, xr)
        return None, dw1, db1, dw2, db2, dw3, db3


def pointnet(x, w1, b1, w2, b2, w3, b3):
    return _PointNet.apply(x, w1, b1, w2, b2, w3, b3)


# -------------------------------------------------------------------- graph attention softmax + aggregate
class _GatSoftmaxAggr(Function):
    @staticmethod
    def forward(ctx, t, v_hm, g, n_heads, aggr):
        xx, prob, arg = ops.gat_softmax_aggr(t, v_hm, g.dst, g.row_ptr, g.num_nodes, n_heads, aggr)
        ctx.save_for_backward(prob, v_hm, arg)
        ctx.g, ctx.n_heads, ctx.aggr = g, n_heads, aggr
        ctx.mark_non_differentiable(prob)
        return xx, prob

    @staticmethod
    def backward(ctx, dxx, _dprob):
        prob, v_hm, arg = ctx.saved_tensors
        g = ctx.g
        dt, dv = ops.gat_softmax_aggr_bwd(_c(dxx), prob, v_hm, g.dst, g.row_ptr, arg, g.num_nodes, ctx.n_heads, ctx.aggr)
        return dt, dv, None, None, None


def gat_softmax_aggr(t, v_hm, g, n_heads: int, aggr: str):
    """t [E*H, d_o] (CSR edge order), v_hm [N, H*d_o] head-major -> (xx [N, H*d_o] interleaved c*H+h, prob [E*H, d_o])."""
    return _GatSoftmaxAggr.apply(t.contiguous(), v_hm, g, n_heads, aggr)


# ------------------------------------------------------------------------------ node attention (A6 + A7)
class _NodeAttn(Function):
    @staticmethod
    def forward(ctx, q, k, v, bias, sctx, n_heads):
        out = ops.node_attn_bias(q, k, v, bias, sctx.pair_off, sctx.seg_start, sctx.seg_end, n_heads, sctx.max_scene)
        ctx.save_for_backward(q, k, v, bias)
        ctx.sctx, ctx.n_heads = sctx, n_heads
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, bias = ctx.saved_tensors
        s = ctx.sctx
        dq, dk, dv, dbias = ops.node_attn_bias_bwd(q, k, v, bias, s.pair_off, s.seg_start, s.seg_end, _c(dout), ctx.n_heads,
                                                   s.max_scene)
        return dq, dk, dv, dbias, None, None


def node_attn(q, k, v, bias, sctx, n_heads: int):
    return _NodeAttn.apply(q, k, v, bias.contiguous(), sctx, n_heads)


# ------------------------------------------------------------------------ edge cross-attention (A9)
FLASH_BWD_QUERY_BLOCK = 8192     # queries per score block in the materialising backward ([block, nk] fp32 buffers)
FLASH_BWD_STREAMING = os.environ.get("VLSAT_FLASH_BWD", "stream") != "blocks"    # round-1 block path kept for the other engines


class _FlashAttn(Function):
    """Forward: streaming tensor-core kernel (scores never materialised) + log-sum-exp. Backward: per head and per
    block of queries, S = Q K^T and dP = dO V^T as projections, one fused score-stage kernel (vlsat_attn_prob_bwd)
    and three more projections (dV += P^T dO, dK += dS^T Q, dQ = dS K)."""

    @staticmethod
    def forward(ctx, q, k, v, n_heads):
        nk = k.shape[0]
        d = q.shape[1]
        ctx.prep = None
        if d == n_heads * 64 and ops.tensor_cores_enabled():
            if FLASH_BWD_STREAMING and q.shape[0] > 0 and nk > 0:
                # one pass per operand gives the pair the forward reads and the transposed pair the backward reads
                # (v: the forward takes V^T, the backward V)
                qp, qt = ops.bf16_split_t(q)
                kp, kt = ops.bf16_split_t(k)
                vp, vt = ops.bf16_split_t(v)
                out, lse = ops.flash_attn_bf16(qp, kp, vt, nk, n_heads, want_lse=True)
                ctx.prep = (qp, qt, kp, kt, vp)
            else:
                vt = ops.transpose(v)                               # [D, round4(nk)]
                out, lse = ops.flash_attn_bf16(q, k, vt, nk, n_heads, want_lse=True)
        else:
            out, lse = ops.flash_attn(q, k, v, n_heads, want_lse=True)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.n_heads = n_heads
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        H = ctx.n_heads
        nq, d = q.shape
        nk = k.shape[0]
        dk_ = d // H
        scale = 1.0 / math.sqrt(dk_)
        dout = _c(dout)
        if ctx.prep is not None:
            # streaming tcgen05 backward: S, P, dP, dS live in tensor memory only (csrc/flash_attn_bwd.cu)
            dq, dk, dv = ops.flash_attn_bf16_bwd(q, k, v, dout, out, lse, H, prep=ctx.prep)
            return dq, dk, dv, None
        delta = ops.rowdot_heads(dout, out, H)                      # [H, nq]
        dq = torch.empty_like(q)
        dk = torch.empty((nk, d), device=q.device, dtype=torch.float32)
        dv = torch.empty((nk, d), device=q.device, dtype=torch.float32)
        pairs = ops.tensor_cores_enabled() and ops.default_fmt(8) == ops.FMT_BF16 and nk % 8 == 0 and nk >= 32
        for h in range(H):
            cs = slice(h * dk_, (h + 1) * dk_)
            kh, vh = k[:, cs], v[:, cs]
            kh_t = ops.transpose(kh)                                # [dk, round4(nk)]
            for i0 in range(0, nq, FLASH_BWD_QUERY_BLOCK):
                i1 = min(nq, i0 + FLASH_BWD_QUERY_BLOCK)
                qb, dob = q[i0:i1, cs], dout[i0:i1, cs]
                s = ops.linear(qb, kh, cache_w=False)               # [nb, nk] raw scores
                dp = ops.linear(dob, vh, cache_w=False)             # [nb, nk]
                first = i0 == 0
                if pairs and i1 - i0 >= 32:
                    # the score stage writes dS, dS^T, P^T as the bf16 pairs the next three products read: the fp32
                    # [block, nk] matrices are neither stored nor re-read for a split pass
                    ds, ds_t, p_t = ops.attn_prob_bwd_pairs(s, dp, lse[h, i0:i1], delta[h, i0:i1], scale)
                    ops.linear(p_t, ops.transpose(dob, 8), out=dv[:, cs], residual=None if first else dv[:, cs], cache_w=False)
                    ops.linear(ds_t, ops.transpose(qb, 8), out=dk[:, cs], residual=None if first else dk[:, cs], cache_w=False)
                    ops.linear(ds, kh_t[:, :nk], out=dq[i0:i1, cs], cache_w=False)
                    continue
                ds, ds_t, p_t = ops.attn_prob_bwd(s, dp, lse[h, i0:i1], delta[h, i0:i1], scale)
                ops.linear(p_t, ops.transpose(dob), out=dv[:, cs], residual=None if first else dv[:, cs], cache_w=False)
                ops.linear(ds_t, ops.transpose(qb), out=dk[:, cs], residual=None if first else dk[:, cs], cache_w=False)
                ops.linear(ds, kh_t[:, :nk], out=dq[i0:i1, cs], cache_w=False)
        return dq, dk, dv, None


def flash_attn(q, k, v, n_heads: int):
    return _FlashAttn.apply(q, k, v, n_heads)
