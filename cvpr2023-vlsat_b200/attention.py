"""Host-side mirror of ``src/model/transformer/attention.py`` (MultiHeadAttention /
ScaledDotProductAttention): same constructor arguments, same parameter names (``attention.fc_q`` ...
``layer_norm``), same initialisation; the forward runs on the vlsat_b200 CUDA kernels.

The reference evaluates dense [1,H,n,n] score tensors with a block-diagonal mask and an additive bias
(attention.py:60-75). Here the two uses on the hot path get their own entry points:
  * ``attend_scenes`` - attention between the nodes of each scene with the distance bias computed
    inside the kernel (network_MMG.py:217-218);
  * ``attend_all``    - unmasked attention over all keys with a streaming softmax (network_MMG.py:231).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops
from ._cache import DerivedCache, require_inference


_last_err_flag = {}          # device -> the scene-range error flag of the most recent context (read by validate_inputs)


def validate_inputs(device=None) -> None:
    """Raise if the ``batch_ids`` of the most recent forward on ``device`` were not non-decreasing (DataLoader.py:172 always
    writes them sorted; the segment kernels binary-search them, so unsorted ids would give silently wrong attention masks).
    One small D2H read, i.e. a host sync: the eager inference path calls it after the last launch of a forward, the
    training path piggy-backs on the sync it already has, CUDA-graph replays check the previous replay's flag
    (``GraphedForward``)."""
    for dev, flag in list(_last_err_flag.items()):
        if device is not None and torch.device(device) != dev:
            continue
        if flag is not None and int(flag.item()) != 0:
            _last_err_flag[dev] = None
            raise RuntimeError("vlsat_b200: batch_ids must be non-decreasing scene ids (one contiguous block of nodes per scene, as "
                               "collate_fn_mmg builds them, src/dataset/DataLoader.py:153-176)")


class SceneContext:
    """Per-batch bookkeeping for node attention: scene range of every node + packed bias MLP."""

    def __init__(self, batch_ids: torch.Tensor, centres: torch.Tensor, fc_pack: torch.Tensor, num_heads: int):
        self.seg_start, self.seg_end, self.err_flag = ops.scene_ranges(batch_ids)
        _last_err_flag[batch_ids.device] = self.err_flag
        self.centres = centres.detach().contiguous()
        self.fc_pack = fc_pack
        self.num_heads = num_heads
        # the bias depends on the centres only: one table serves every attention call of this forward
        self.bias_table = ops.node_bias_table(self.centres, self.seg_start, self.seg_end, fc_pack, num_heads)


class ScaledDotProductAttention(nn.Module):
    """Parameter container with the reference's names/init (attention.py:11-39)."""

    def __init__(self, d_model, d_k, d_v, h):
        super().__init__()
        self.fc_q = nn.Linear(d_model, h * d_k)
        self.fc_k = nn.Linear(d_model, h * d_k)
        self.fc_v = nn.Linear(d_model, h * d_v)
        self.fc_o = nn.Linear(h * d_v, d_model)
        self.d_model, self.d_k, self.d_v, self.h = d_model, d_k, d_v, h
        self.init_weights()

    def init_weights(self):
        for fc in (self.fc_q, self.fc_k, self.fc_v, self.fc_o):
            nn.init.xavier_uniform_(fc.weight)
            nn.init.constant_(fc.bias, 0)


class MultiHeadAttention(nn.Module):
    def __init__(self, d_model, d_k, d_v, h, dropout=.1, identity_map_reordering=False, can_be_stateful=False,
                 attention_module=None, attention_module_kwargs=None):
        super().__init__()
        if identity_map_reordering or can_be_stateful or attention_module is not None:
            raise NotImplementedError("only the plain post-LN MultiHeadAttention used on the VL-SAT hot path is built")
        if d_k != d_v or h * d_k != d_model:
            raise NotImplementedError("vlsat_b200 attention kernels need d_k == d_v == d_model / h")
        self.identity_map_reordering = False
        self.attention = ScaledDotProductAttention(d_model=d_model, d_k=d_k, d_v=d_v, h=h)
        self.dropout = nn.Dropout(p=dropout)
        self.layer_norm = nn.LayerNorm(d_model)
        self.can_be_stateful = False
        self._cache = DerivedCache()

    # ---- derived weights -------------------------------------------------------------------------
    def _w(self, which: str):
        a = self.attention
        if which == "qkv":
            srcs = (a.fc_q.weight, a.fc_k.weight, a.fc_v.weight, a.fc_q.bias, a.fc_k.bias, a.fc_v.bias)
            return self._cache.get("qkv", srcs, lambda: (torch.cat([a.fc_q.weight, a.fc_k.weight, a.fc_v.weight], 0).contiguous(),
                                                          torch.cat([a.fc_q.bias, a.fc_k.bias, a.fc_v.bias], 0).contiguous()))
        if which == "kv":
            srcs = (a.fc_k.weight, a.fc_v.weight, a.fc_k.bias, a.fc_v.bias)
            return self._cache.get("kv", srcs, lambda: (torch.cat([a.fc_k.weight, a.fc_v.weight], 0).contiguous(),
                                                         torch.cat([a.fc_k.bias, a.fc_v.bias], 0).contiguous()))
        raise KeyError(which)

    def _project(self, q_in: torch.Tensor, kv_in: torch.Tensor, same: bool, q_split=None, kv_split=None):
        a = self.attention
        d = a.h * a.d_k
        if same:
            w, b = self._w("qkv")
            qkv = ops.linear(q_in, w, b, x_split=q_split)
            return qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
        q = ops.linear(q_in, a.fc_q.weight.detach(), a.fc_q.bias.detach(), x_split=q_split)
        w, b = self._w("kv")
        kv = ops.linear(kv_in, w, b, x_split=kv_split)
        return q, kv[:, :d], kv[:, d:]

    def _finish(self, q_in: torch.Tensor, att: torch.Tensor, relu: bool, out: Optional[torch.Tensor], emit_split: bool = False):
        a = self.attention
        # fc_o with the residual folded into the GEMM epilogue, then LayerNorm (attention.py:76,122); with emit_split the
        # LayerNorm also writes the (hi, lo) pair the next projection reads: returns (y, pair or None)
        pre = ops.linear(att, a.fc_o.weight.detach(), a.fc_o.bias.detach(), residual=q_in, alpha=1.0, beta=1.0)
        return ops.add_layernorm(pre, None, self.layer_norm.weight.detach(), self.layer_norm.bias.detach(),
                                 eps=self.layer_norm.eps, relu=relu, out=out, emit_split=emit_split)

    # ---- hot-path entry points -------------------------------------------------------------------
    def attend_scenes(self, q_in, kv_in, ctx: SceneContext, relu: bool = False, out=None, q_split=None, kv_split=None,
                      emit_split: bool = False):
        """LN(q_in + fc_o(softmax_scene(QK^T/sqrt(dk) + bias) V)); q_in, kv_in are [N, d_model]. ``q_split`` / ``kv_split``:
        (hi, lo) pairs of the inputs when a producer emitted them; ``emit_split``: return (y, pair of y)."""
        require_inference(self, "MultiHeadAttention")
        if ctx.num_heads != self.attention.h:
            raise ValueError(f"distance bias has {ctx.num_heads} heads but attention has {self.attention.h} "
                             "(network_GNN.py:211 hard-codes 8 heads for this reason)")
        q, k, v = self._project(q_in, kv_in, q_in is kv_in, q_split, kv_split)
        att = ops.node_attn(q, k, v, ctx.centres, ctx.seg_start, ctx.seg_end, ctx.fc_pack, self.attention.h, ctx.bias_table)
        return self._finish(q_in, att, relu, out, emit_split)

    def attend_all(self, q_in, kv_in, relu: bool = False, out=None, q_split=None, kv_split=None, emit_split: bool = False):
        """LN(q_in + fc_o(softmax(QK^T/sqrt(dk)) V)) over all keys, no mask/bias; 2-D inputs.
        ``q_split`` / ``kv_split``: (hi, lo) pairs of the inputs if the producer already emitted them."""
        from . import train_path as T
        if T.differentiable(self):
            y = T.mha_all(self, q_in, kv_in, relu_out=relu)
            if out is not None:
                raise ValueError("attend_all: `out=` is an inference-path option")
            return (y, None) if emit_split else y
        a = self.attention
        if a.d_k == 64 and ops.tensor_cores_enabled() and kv_in.shape[1] % 4 == 0 and kv_in.shape[1] >= 32:
            # tensor-core path: Q, K row-major; the value projection is emitted transposed (V^T = W_v x^T).
            # The projections write the tf32 splits the attention kernel consumes directly from their epilogues.
            nk = kv_in.shape[0]
            if ops.ATTN_ENGINE == "bf16x3":
                # attention operands as bf16 (hi, lo) pairs, written by the projection epilogues: the attention kernel is
                # L2-bandwidth bound and bf16 pairs halve its traffic (csrc/flash_attn_bf16.cu)
                if kv_split is None:
                    kv_split = ops.split_pair(kv_in)
                def project_q():
                    return ops.linear(q_in, a.fc_q.weight.detach(), a.fc_q.bias.detach(), x_split=q_split, emit_split="bf16", want_y=False)[1]

                def project_kv():
                    _, k_ = ops.linear(kv_in, a.fc_k.weight.detach(), a.fc_k.bias.detach(), x_split=kv_split, emit_split="bf16", want_y=False)
                    if nk % 8 == 0:
                        _, vt_ = ops.linear(a.fc_v.weight.detach(), kv_in, a.fc_v.bias.detach(), bias_per_row=True, x_is_weight=True,
                                            w_split=kv_split, emit_split="bf16", want_y=False)
                    else:
                        vt_ = torch.empty((a.h * a.d_v, (nk + 3) // 4 * 4), device=q_in.device, dtype=torch.float32)
                        ops.linear(a.fc_v.weight.detach(), kv_in, a.fc_v.bias.detach(), out=vt_[:, :nk], bias_per_row=True,
                                   x_is_weight=True, w_split=kv_split)
                    return k_, vt_
                # the query projection (2D edges) next to the key / value projections (3D edges): independent, two streams
                q, (k, vt) = ops.fork_join(project_q, project_kv, q_in.device)
                att = ops.flash_attn_bf16(q, k, vt, nk, a.h)
                return self._finish(q_in, att, relu, out, emit_split)
            if kv_split is None or ops.pair_fmt(kv_split) != ops.FMT_TF32:
                kv_split = ops.tf32_split(kv_in)
            if q_split is not None and ops.pair_fmt(q_split) != ops.FMT_TF32:
                q_split = None
            _, q = ops.linear(q_in, a.fc_q.weight.detach(), a.fc_q.bias.detach(), x_split=q_split, emit_split="tf32", want_y=False)
            _, k = ops.linear(kv_in, a.fc_k.weight.detach(), a.fc_k.bias.detach(), x_split=kv_split, emit_split="tf32", want_y=False)
            if nk % 4 == 0:
                _, vt = ops.linear(a.fc_v.weight.detach(), kv_in, a.fc_v.bias.detach(), bias_per_row=True, x_is_weight=True,
                                   w_split=kv_split, emit_split="tf32", want_y=False)
            else:
                vt = torch.zeros((a.h * a.d_v, (nk + 3) // 4 * 4), device=q_in.device, dtype=torch.float32)
                ops.linear(a.fc_v.weight.detach(), kv_in, a.fc_v.bias.detach(), out=vt[:, :nk], bias_per_row=True,
                           x_is_weight=True, w_split=kv_split)
            att = ops.flash_attn_tc(q, k, vt, nk, a.h)
        else:
            q, k, v = self._project(q_in, kv_in, q_in is kv_in)
            att = ops.flash_attn(q, k, v, a.h)
        return self._finish(q_in, att, relu, out, emit_split)

    # ---- reference signature -----------------------------------------------------------------------
    def forward(self, queries, keys, values, attention_mask=None, attention_weights=None, way='mul', use_knn=False,
                output_attn=False):
        """Same signature as attention.py:105, [b_s, n, d] inputs. Without mask / weights the streaming kernel runs
        (``attend_all``). With the dense ``attention_mask`` [b_s, 1 | h, nq, nk] (0 = masked) and ``attention_weights``
        [b_s, h, nq, nk] that the reference's own ``MMG.forward`` builds (network_MMG.py:183-205) and passes with
        ``way='add'`` (:217-218), the exact-fp32 dense kernel runs (csrc/dense_attn.cu; inference only) - so this module can
        replace the reference's MultiHeadAttention under the reference's MMG. The fast path of this package never builds
        those tensors (``MMG.forward`` here calls ``attend_scenes``)."""
        if use_knn or output_attn:
            raise NotImplementedError("use_knn / output_attn (attention.py:62-63,124-125) are not used on the VL-SAT path and not built")
        if queries.dim() != 3 or keys is not values and keys.data_ptr() != values.data_ptr():
            raise NotImplementedError("expected [b_s, n, d] inputs with keys is values (the only use on the path)")
        if attention_mask is None and attention_weights is None:
            outs = [self.attend_all(queries[b], keys[b]) for b in range(queries.shape[0])]
            return torch.stack(outs, 0)
        require_inference(self, "MultiHeadAttention with dense attention_mask / attention_weights")
        if way not in ("mul", "add"):
            raise NotImplementedError(way)                       # attention.py:72
        a = self.attention
        outs = []
        for b in range(queries.shape[0]):
            q_in, kv_in = queries[b].contiguous(), keys[b].contiguous()
            q, k, v = self._project(q_in, kv_in, queries is keys or queries.data_ptr() == keys.data_ptr())
            w = attention_weights[b] if attention_weights is not None else None
            m = attention_mask[b] if attention_mask is not None else None
            if m is not None and m.dim() == 3 and m.shape[0] == 1:
                m = m[0]
            att = ops.dense_attn(q, k, v, a.h, weights=w, way=way, mask=m)
            outs.append(self._finish(q_in, att, False, None))
        return torch.stack(outs, 0)
