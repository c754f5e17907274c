"""Host-side mirror of the graph-attention layer: ``GraphEdgeAttenNetwork`` / ``MultiHeadedEdgeAttention``
(src/model/model_utils/network_MMG.py:12-112, twin in network_GNN.py:44-156) plus ``Gen_Index`` /
``Aggre_Index`` / ``MLP`` / ``build_mlp`` (network_util.py:13-73). Same constructor arguments and
parameter names; forward on the vlsat_b200 kernels.

What changes relative to the reference's execution (not its maths):
  * ``Linear`` is row-wise, so ``proj_query(x)[src]``, ``proj_value(x)[dst]`` and the two node blocks of
    ``nn_edge.0`` (``W[:, :D_n] x[src]``, ``W[:, D_n+D_e:] x[dst]``) are computed once per NODE by a
    single fused projection and gathered per edge inside the consumers - the [E, D] gathered copies
    and the [E, 2 D_n + D_e] concatenation never exist;
  * edges are grouped by source node once per batch (CSR) and the per-(edge, head) MLP, softmax,
    value product and max/add/mean aggregation run in one kernel without atomics.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops
from ._cache import DerivedCache, require_inference


# -------------------------------------------------------------------------------- network_util mirrors
def MLP(channels: list, do_bn=False, on_last=False, drop_out=None):
    """Conv1d(k=1) stack with the reference's module indices (network_util.py:13-28)."""
    if do_bn:
        raise NotImplementedError("use_bn/WITH_BN=True is not supported by vlsat_b200 (mmgnet.json: WITH_BN=false)")
    n = len(channels)
    layers = []
    offset = 0 if on_last else 1
    for i in range(1, n):
        layers.append(nn.Conv1d(channels[i - 1], channels[i], kernel_size=1, bias=True))
        if i < (n - offset):
            layers.append(nn.ReLU())
            if drop_out is not None:
                layers.append(nn.Dropout(drop_out))
    return nn.Sequential(*layers)


def build_mlp(dim_list, activation='relu', do_bn=False, dropout=0, on_last=False):
    """Linear stack with the reference's module indices (network_util.py:31-47)."""
    if do_bn:
        raise NotImplementedError("use_bn/WITH_BN=True is not supported by vlsat_b200 (mmgnet.json: WITH_BN=false)")
    if activation != 'relu':
        raise NotImplementedError("only relu MLPs are on the VL-SAT hot path")
    layers = []
    for i in range(len(dim_list) - 1):
        layers.append(nn.Linear(dim_list[i], dim_list[i + 1]))
        final_layer = (i == len(dim_list) - 2)
        if not final_layer or on_last:
            layers.append(nn.ReLU())
        if dropout > 0:
            layers.append(nn.Dropout(p=dropout))
    return nn.Sequential(*layers)


def _flow_rows(flow: str):
    if flow == "target_to_source":
        return 0, 1
    if flow == "source_to_target":
        return 1, 0
    raise ValueError(f"unknown flow {flow!r}")


class GraphContext:
    """Edge bookkeeping shared by every GAT call of a forward pass: oriented edge_index (row 0 = the
    node that receives the aggregate = ``edge_index[i]``, row 1 = ``edge_index[j]``) and its CSR."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, flow: str = "target_to_source"):
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise TypeError("edge_index must be an int64 [2, E] tensor (PyG MessagePassing contract)")
        i, _ = _flow_rows(flow)
        ei = (edge_index if i == 0 else edge_index.flip(0)).contiguous()
        self.num_nodes = num_nodes
        self.num_edges = ei.shape[1]
        # perm[i] = original id of the i-th edge in source-sorted (CSR) order (stable). Every per-edge tensor
        # is processed in that order (``to_sorted`` / ``to_original``): all per-edge maths is row-wise and the
        # edge cross-attention is permutation-equivariant, so this is exact - and it is the identity for the
        # reference's own data, whose edge lists are always grouped by subject (dataset_3dssg.py:263-266).
        self.row_ptr, self.perm = ops.build_csr(ei[0], num_nodes)
        self.edge_index = ops.permute_edges(ei, self.perm) if self.num_edges else ei
        self.src = self.edge_index[0]
        self.dst = self.edge_index[1]
        self.identity = torch.arange(self.num_edges, device=ei.device, dtype=torch.int32)
        self._head_rows = {}

    def head_rows(self, num_heads: int) -> torch.Tensor:
        """int64 [E*H]: row (e, h) of a head-major edge tensor -> row src(e)*H + h of a head-major node tensor."""
        r = self._head_rows.get(num_heads)
        if r is None:
            h = torch.arange(num_heads, device=self.src.device, dtype=torch.int64)
            r = (self.src.view(-1, 1) * num_heads + h.view(1, -1)).reshape(-1).contiguous()
            self._head_rows[num_heads] = r
        return r

    def to_sorted(self, edge_rows: torch.Tensor, emit_split: bool = False):
        if not self.num_edges:
            return (edge_rows, None) if emit_split else edge_rows
        return ops.permute_rows(edge_rows, self.perm, gather=True, emit_split=emit_split)

    def to_original(self, edge_rows: torch.Tensor, emit_split: bool = False):
        if not self.num_edges:
            return (edge_rows, None) if emit_split else edge_rows
        return ops.permute_rows(edge_rows, self.perm, gather=False, emit_split=emit_split)


class Gen_Index(nn.Module):
    """``x_i, x_j = x[edge_index[i]], x[edge_index[j]]`` (network_util.py:50-62). Kept for API parity; the
    fused layer never materialises these."""

    def __init__(self, flow="target_to_source"):
        super().__init__()
        self.flow = flow

    def forward(self, x, edges_indices):
        i, j = _flow_rows(self.flow)
        return x.index_select(0, edges_indices[i]), x.index_select(0, edges_indices[j])


class Aggre_Index(nn.Module):
    """Scatter-aggregate onto ``edge_index[i]`` (network_util.py:64-73) via the CSR segmented kernel."""

    def __init__(self, aggr='add', node_dim=-2, flow="source_to_target"):
        super().__init__()
        if aggr not in ops.AGGR:
            raise NotImplementedError(aggr)
        self.aggr, self.flow = aggr, flow

    def forward(self, x, edge_index, dim_size):
        raise NotImplementedError(
            "stand-alone Aggre_Index is fused into vlsat_gat_edge_fwd; use GraphEdgeAttenNetwork.forward")


# ------------------------------------------------------------------------------------------ the layer
class MultiHeadedEdgeAttention(nn.Module):
    def __init__(self, num_heads: int, dim_node: int, dim_edge: int, dim_atten: int, use_bn=False,
                 attention='fat', use_edge: bool = True, **kwargs):
        super().__init__()
        assert dim_node % num_heads == 0
        assert dim_edge % num_heads == 0
        assert dim_atten % num_heads == 0
        self.name = 'MultiHeadedEdgeAttention'
        self.dim_node, self.dim_edge, self.dim_atten = dim_node, dim_edge, dim_atten
        self.d_n = d_n = dim_node // num_heads
        self.d_e = d_e = dim_edge // num_heads
        self.d_o = d_o = dim_atten // num_heads
        self.num_heads = num_heads
        self.use_edge = use_edge
        self.nn_edge = build_mlp([dim_node * 2 + dim_edge, (dim_node + dim_edge), dim_edge], do_bn=use_bn, on_last=False)
        self.mask_obj = 0.5
        drop = kwargs.get('DROP_OUT_ATTEN', None)
        self.attention = attention
        if attention != 'fat':
            raise NotImplementedError(f"attention={attention!r}: only 'fat' exists in the reference (network_MMG.py:70)")
        if use_edge:
            self.nn = MLP([d_n + d_e, d_n + d_e, d_o], do_bn=use_bn, drop_out=drop)
        else:
            self.nn = MLP([d_n, d_n * 2, d_o], do_bn=use_bn, drop_out=drop)
        self.proj_edge = build_mlp([dim_edge, dim_edge])
        self.proj_query = build_mlp([dim_node, dim_node])
        self.proj_value = build_mlp([dim_node, dim_atten])
        self._cache = DerivedCache()
        self.last_edge_split = None      # tf32 split of the last new edge feature (tensor-core path), else None

    # ---- derived weights -------------------------------------------------------------------------
    def _convs(self):
        return [m for m in self.nn if isinstance(m, nn.Conv1d)]

    def attn_mlp_weights(self):
        c1, c2 = self._convs()
        srcs = (c1.weight, c1.bias, c2.weight, c2.bias)
        return self._cache.get("mlp", srcs, lambda: (c1.weight.squeeze(-1).contiguous(), c1.bias.contiguous(),
                                                      c2.weight.squeeze(-1).contiguous(), c2.bias.contiguous()))

    def node_projection_weights(self):
        """[proj_query; proj_value; nn_edge.0[:, :D_n]; nn_edge.0[:, D_n+D_e:]] as one [*, D_n] matrix and
        the matching bias (zeros for the nn_edge blocks: its bias is added on the edge side)."""
        wq, wv, w1 = self.proj_query[0], self.proj_value[0], self.nn_edge[0]
        dn, de = self.dim_node, self.dim_edge

        def build():
            w = torch.cat([wq.weight, wv.weight, w1.weight[:, :dn], w1.weight[:, dn + de:]], 0).contiguous()
            b = torch.cat([wq.bias, wv.bias, torch.zeros(2 * w1.weight.shape[0], device=w.device, dtype=w.dtype)], 0).contiguous()
            return w, b
        return self._cache.get("node", (wq.weight, wq.bias, wv.weight, wv.bias, w1.weight), build)

    def tc_weights(self):
        """Head-major derived weights of the tensor-core edge kernel (csrc/gat_tc.cu):
        node projection  [W_qc ; W_v' ; W1[:, :D_n] ; W1[:, D_n+D_e:]]  with
          W_qc[h*hid + j, :] = sum_c C1[j, c] W_q[c*H + h, :]   (proj_query folded into the first MLP layer),
          W_v'[h*d_o + c, :] = W_v[c*H + h, :];
        edge projection  W_pe'[h*d_e + c, :] = W_pe[c*H + h, :];  C1k = C1[:, d_n:] and C2 as bf16 (hi, lo) pairs."""
        wq, wv, w1, pe = self.proj_query[0], self.proj_value[0], self.nn_edge[0], self.proj_edge[0]
        c1, c2 = self._convs()
        H, dn, de, do = self.num_heads, self.d_n, self.d_e, self.d_o
        Dn, De = self.dim_node, self.dim_edge

        def build():
            C1 = c1.weight.squeeze(-1)
            hid = C1.shape[0]
            w_qc = torch.einsum("jc,chi->hji", C1[:, :dn], wq.weight.view(dn, H, Dn)).reshape(H * hid, Dn)
            b_qc = (torch.einsum("jc,ch->hj", C1[:, :dn], wq.bias.view(dn, H)) + c1.bias.view(1, hid)).reshape(H * hid)
            w_v = wv.weight.view(do, H, Dn).permute(1, 0, 2).reshape(H * do, Dn)
            b_v = wv.bias.view(do, H).t().reshape(H * do)
            hid1 = w1.weight.shape[0]
            w_node = torch.cat([w_qc, w_v, w1.weight[:, :Dn], w1.weight[:, Dn + De:]], 0).contiguous()
            b_node = torch.cat([b_qc, b_v, torch.zeros(2 * hid1, device=w_node.device, dtype=w_node.dtype)], 0).contiguous()
            w_pe = pe.weight.view(de, H, De).permute(1, 0, 2).reshape(H * de, De).contiguous()
            b_pe = pe.bias.view(de, H).t().reshape(H * de).contiguous()
            c1k = C1[:, dn:].contiguous()
            c2w = c2.weight.squeeze(-1).contiguous()
            return dict(w_node=w_node, b_node=b_node, w_pe=w_pe, b_pe=b_pe, c1k=ops.bf16_split(c1k), c2=ops.bf16_split(c2w),
                        c2b=c2.bias.detach().clone().contiguous(), hid=hid)
        srcs = (wq.weight, wq.bias, wv.weight, wv.bias, w1.weight, pe.weight, pe.bias, c1.weight, c1.bias, c2.weight, c2.bias)
        return self._cache.get("tc", srcs, build)

    def tc_eligible(self) -> bool:
        hid = self._convs()[0].weight.shape[0]
        return (self.use_edge and g_aggr(self) == "max" and ops.gat_tc_supported(self.num_heads, self.d_e, hid, self.d_o)
                and self.dim_node % 4 == 0 and self.dim_edge % 4 == 0)

    def fused_tc(self, x: torch.Tensor, edge: torch.Tensor, g: GraphContext, xx_out: torch.Tensor, want_prob: bool = False,
                 edge_split=None, x_split=None):
        """Tensor-core version of ``fused`` (same contract; ``edge`` and ``g`` in CSR edge order). Intermediate
        activations that only feed another projection are never stored unsplit: the producing epilogue emits
        the tf32 (hi, lo) pair the consumer reads. Also returns the split of the new edge feature."""
        require_inference(self, "MultiHeadedEdgeAttention")
        w = self.tc_weights()
        H, hid, da, de = self.num_heads, w["hid"], self.dim_atten, self.dim_edge
        dn = self.dim_node
        hid1 = self.nn_edge[0].weight.shape[0]
        node = ops.linear(x, w["w_node"], w["b_node"], x_split=x_split)   # [N, H*hid + D_a + 2*hid1]
        qc, v_hm = node[:, :H * hid], node[:, H * hid:H * hid + da]
        a_src, b_dst = node[:, H * hid + da:H * hid + da + hid1], node[:, H * hid + da + hid1:]
        w1, w2 = self.nn_edge[0], self.nn_edge[2]
        if g.num_edges == 0:
            new_edge = torch.empty((0, de), device=x.device, dtype=torch.float32)
            ops.gat_edge_tc(torch.empty((0, H * self.d_e), device=x.device), qc, v_hm, g.src, g.dst, w["c1k"], w["c2"], w["c2b"],
                            g.num_nodes, H, xx_out, d_n=self.d_n)
            self.last_edge_split = None
            return new_edge, (torch.empty((0, self.d_o, H), device=x.device) if want_prob else None)
        if edge_split is None:
            edge_split = ops.split_pair(edge)
        _, h = ops.linear(edge, w1.weight.detach()[:, dn:dn + de], w1.bias.detach(), act=ops.ACT_RELU,
                          gather=(a_src, g.src, b_dst, g.dst), x_split=edge_split, emit_split=True, want_y=False)
        new_edge, self.last_edge_split = ops.linear(h, w2.weight.detach(), w2.bias.detach(), emit_split=True)
        _, k_hm = ops.linear(edge, w["w_pe"], w["b_pe"], x_split=edge_split, emit_split="bf16", want_y=False)   # rows (e, h)
        _, prob = ops.gat_edge_tc(k_hm, qc, v_hm, g.src, g.dst, w["c1k"], w["c2"], w["c2b"], g.num_nodes, H, xx_out,
                                  want_prob=want_prob, d_n=self.d_n)
        return new_edge, prob

    # ---- fused path ------------------------------------------------------------------------------
    def fused(self, x: torch.Tensor, edge: torch.Tensor, g: GraphContext, xx_out: torch.Tensor,
              want_prob: bool = False, edge_split=None, x_split=None):
        """x [N, D_n] (may be a column slice), edge [E, D_e] in CSR edge order -> writes the aggregate into
        ``xx_out`` [N, D_a]; returns (new edge feature [E, D_e], prob or None), both in CSR edge order."""
        require_inference(self, "MultiHeadedEdgeAttention")
        self.last_edge_split = None
        if self.tc_eligible():
            return self.fused_tc(x, edge, g, xx_out, want_prob, edge_split=edge_split, x_split=x_split)
        dn, de, da = self.dim_node, self.dim_edge, self.dim_atten
        hid1 = self.nn_edge[0].weight.shape[0]
        w_node, b_node = self.node_projection_weights()
        node = ops.linear(x, w_node, b_node)                       # [N, D_n + D_a + 2*hid1]
        q, v = node[:, :dn], node[:, dn:dn + da]
        a_src, b_dst = node[:, dn + da:dn + da + hid1], node[:, dn + da + hid1:]
        w1 = self.nn_edge[0]
        h = ops.linear(edge, w1.weight.detach()[:, dn:dn + de], w1.bias.detach(), act=ops.ACT_RELU,
                       gather=(a_src, g.src, b_dst, g.dst))        # relu(W1 . cat[x_i, e, x_j] + b1)
        w2 = self.nn_edge[2]
        new_edge = ops.linear(h, w2.weight.detach(), w2.bias.detach())
        k = None
        if self.use_edge:
            pe = self.proj_edge[0]
            k = ops.linear(edge, pe.weight.detach(), pe.bias.detach())
        c1, c1b, c2, c2b = self.attn_mlp_weights()
        _, prob, _ = ops.gat_edge(q, v, k, g.edge_index, g.row_ptr, g.identity, c1, c1b, c2, c2b, self.num_heads,
                                  aggr=g_aggr(self), use_edge=self.use_edge, want_prob=want_prob, out=xx_out)
        return new_edge, prob

    # ---- reference signature (already-gathered per-edge inputs) -------------------------------------
    def forward(self, query, edge, value, weight=None, istrain=False):
        """(x_i, edge, x_j) -> (prob * value [E, D_a], new edge feature, prob [E, d_o, H]); every edge is
        treated as its own one-edge node so the same fused kernel serves this entry point."""
        e_cnt = query.shape[0]
        ar = torch.arange(e_cnt, device=query.device, dtype=torch.int64)
        from . import train_path as T
        if T.differentiable(self):
            g = GraphContext(torch.stack([ar, ar], 0), e_cnt)
            x, new_edge, prob = T.edge_attention(self, query.contiguous(), edge.contiguous(), g, x_value=value.contiguous(), aggr="add")
            return x, new_edge, prob.view(e_cnt, self.num_heads, self.d_o).permute(0, 2, 1).contiguous()
        w1, w2 = self.nn_edge[0], self.nn_edge[2]
        h = ops.linear(torch.cat([query, edge, value], 1), w1.weight.detach(), w1.bias.detach(), act=ops.ACT_RELU)
        new_edge = ops.linear(h, w2.weight.detach(), w2.bias.detach())
        q = ops.linear(query.contiguous(), self.proj_query[0].weight.detach(), self.proj_query[0].bias.detach())
        v = ops.linear(value.contiguous(), self.proj_value[0].weight.detach(), self.proj_value[0].bias.detach())
        k = ops.linear(edge.contiguous(), self.proj_edge[0].weight.detach(), self.proj_edge[0].bias.detach()) if self.use_edge else None
        g = GraphContext(torch.stack([ar, ar], 0), e_cnt)
        c1, c1b, c2, c2b = self.attn_mlp_weights()
        x, prob, _ = ops.gat_edge(q, v, k, g.edge_index, g.row_ptr, g.identity, c1, c1b, c2, c2b, self.num_heads,
                                  aggr="add", use_edge=self.use_edge, want_prob=True)
        return x, new_edge, prob


def g_aggr(m) -> str:
    return getattr(m, "_aggr", "max")


class GraphEdgeAttenNetwork(nn.Module):
    """network_MMG.py:12-41. ``return_prob=True`` gives the network_GNN.py:149-156 twin's 3-tuple."""

    def __init__(self, num_heads, dim_node, dim_edge, dim_atten, aggr='max', use_bn=False,
                 flow='target_to_source', attention='fat', use_edge: bool = True, return_prob: bool = False, **kwargs):
        super().__init__()
        self.name = 'edgeatten'
        self.dim_node, self.dim_edge, self.dim_atten = dim_node, dim_edge, dim_atten
        if attention != 'fat':
            raise NotImplementedError(f"attention={attention!r} (network_MMG.py:22-26: 'distance' can never be constructed "
                                      "because MultiHeadedEdgeAttention asserts 'fat')")
        if aggr not in ops.AGGR:
            raise NotImplementedError(f"aggr={aggr!r}")
        self.aggr, self.flow, self.return_prob = aggr, flow, return_prob
        self.index_get = Gen_Index(flow=flow)
        self.index_aggr = Aggre_Index(aggr=aggr, flow=flow)
        self.edgeatten = MultiHeadedEdgeAttention(dim_node=dim_node, dim_edge=dim_edge, dim_atten=dim_atten,
                                                  num_heads=num_heads, use_bn=use_bn, attention=attention,
                                                  use_edge=use_edge, **kwargs)
        self.edgeatten._aggr = aggr
        self.prop = build_mlp([dim_node + dim_atten, dim_node + dim_atten, dim_node], do_bn=use_bn, on_last=False)

    def forward_fused(self, cat_buf: torch.Tensor, edge_feature: torch.Tensor, g: GraphContext,
                      relu_nodes: bool = False, want_prob: bool = False, x_split=None, edge_split=None, emit_split: bool = False):
        """``cat_buf`` [N, D_n + D_a] already holds x in its first D_n columns; the aggregate is written
        into the remaining columns, so ``prop(cat[x, xx])`` reads one buffer (network_MMG.py:40). ``x_split`` /
        ``edge_split``: (hi, lo) pairs of x / the edge feature when their producers emitted them; ``emit_split``: also
        return the pair of the new node feature (4-tuple)."""
        dn = self.dim_node
        new_edge, prob = self.edgeatten.fused(cat_buf[:, :dn], edge_feature, g, cat_buf[:, dn:], want_prob,
                                              edge_split=edge_split, x_split=x_split)
        p0, p2 = self.prop[0], self.prop[2]
        res = ops.linear_chain(cat_buf, [(p0.weight.detach(), p0.bias.detach(), ops.ACT_RELU),
                                         (p2.weight.detach(), p2.bias.detach(), ops.ACT_RELU if relu_nodes else ops.ACT_NONE)],
                               emit_last=emit_split)
        if emit_split:
            return res[0], new_edge, prob, res[1]
        return res, new_edge, prob

    def forward(self, x, edge_feature, edge_index, weight=None, istrain=False):
        assert x.ndim == 2
        assert edge_feature.ndim == 2
        g = GraphContext(edge_index, x.shape[0], self.flow)
        from . import train_path as T
        if T.differentiable(self):
            from . import autograd as A
            out, new_edge, prob = T.gat_layer(self, x.contiguous(), A.permute_rows(edge_feature.contiguous(), g.perm, True), g)
            new_edge = A.permute_rows(new_edge, g.perm, False)
            if self.return_prob:
                e, H, do = g.num_edges, self.edgeatten.num_heads, self.edgeatten.d_o
                pr = prob.view(e, H, do).permute(0, 2, 1).contiguous()
                if e:
                    pr = ops.permute_rows(pr.view(e, -1), g.perm, gather=False).view(e, do, H)
                return out, new_edge, pr
            return out, new_edge
        cat_buf = torch.empty((x.shape[0], self.dim_node + self.dim_atten), device=x.device, dtype=torch.float32)
        cat_buf[:, :self.dim_node].copy_(x)
        out, new_edge, prob = self.forward_fused(cat_buf, g.to_sorted(edge_feature.contiguous()), g, want_prob=self.return_prob)
        new_edge = g.to_original(new_edge)
        if self.return_prob:
            e = prob.shape[0]
            prob = g.to_original(prob.view(e, -1)).view(prob.shape) if e else prob
            return out, new_edge, prob
        return out, new_edge
