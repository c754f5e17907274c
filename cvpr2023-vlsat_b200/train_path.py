"""Differentiable (training-capable) execution of the hot path on the vlsat_b200 kernels.

Used whenever autograd is recording or a module is in training mode (``differentiable(module)``); the fused
inference path in the module files is used otherwise. Same maths as the inference path, expressed with the
operators of ``autograd.py`` so that every forward and backward step is a vlsat_b200 kernel:

  * dense projections -> ``A.linear`` (backward = two more projections);
  * graph-attention layer -> head-major projections, the per-(edge, head) MLP as two projections over rows (e, h)
    (row gather of the node term in the epilogue), then ``A.gat_softmax_aggr`` (softmax, value product, CSR
    aggregation with saved arg-max) - network_MMG.py:34-41,84-112;
  * node attention -> the distance-bias MLP evaluated ONCE per batch on the compact list of same-scene pairs
    (projections + LayerNorm), then ``A.node_attn`` per call; autograd sums the bias gradient of the 2L calls
    (network_MMG.py:181-205,217-218);
  * edge cross-attention -> ``A.flash_attn`` (network_MMG.py:231);
  * dropout (8 sites), BatchNorm1d batch statistics and running-stat updates in training mode.

Parameter re-layouts (head-major row permutations, Conv1d weight squeezes, column slices) are torch views of the
parameters, so their gradients flow back through autograd's view bookkeeping.
"""
from __future__ import annotations

import torch

from . import autograd as A
from . import ops
from .gat import GraphContext

RELU, NONE, SIGMOID = ops.ACT_RELU, ops.ACT_NONE, ops.ACT_SIGMOID


def differentiable(module: torch.nn.Module) -> bool:
    """True when the call must go through the autograd-capable path."""
    return module.training or torch.is_grad_enabled()


# ------------------------------------------------------------------------------------------- encoders
def pointnet_feat(m, x: torch.Tensor) -> torch.Tensor:
    """PointNetfeat.forward (network_PointNet.py:121-176), no dropout inside."""
    w1, w2, w3 = m.conv1.weight.squeeze(-1), m.conv2.weight.squeeze(-1), m.conv3.weight.squeeze(-1)
    b1, b2, b3 = m.conv1.bias, m.conv2.bias, m.conv3.bias
    if x.shape[2] == 1:
        h = A.linear(x.reshape(x.shape[0], x.shape[1]).contiguous(), w1, b1, RELU)
        h = A.linear(h, w2, b2, RELU, emit_pair=True)
        return A.linear(h, w3, b3, RELU)
    return A.pointnet(x.contiguous(), w1, b1, w2, b2, w3, b3)


def rel_classifier(m, x: torch.Tensor) -> torch.Tensor:
    """PointNetRelClsMulti.forward (network_PointNet.py:328-341): fc1 ReLU fc2 Dropout ReLU fc3 sigmoid. Dropout and ReLU
    commute (both are non-negative scalings), so the ReLU sits in the projection epilogue."""
    h = A.linear(x, m.fc1.weight, m.fc1.bias, RELU, emit_pair=True)
    h = A.linear(h, m.fc2.weight, m.fc2.bias, RELU)
    if m.use_drop_out:
        h = A.dropout(h, m.dropout.p, m.training)
    return A.linear(h, m.fc3.weight, m.fc3.bias, SIGMOID)


# ------------------------------------------------------------------------------------------ attention
_scene_hint = None      # (n_pairs, max_scene) supplied by graph.GraphedTrainStep: no host sync inside a graph capture


def scene_stats(batch_ids: torch.Tensor):
    """(number of same-scene ordered pairs, largest scene) of a batch - one small host sync."""
    if batch_ids.numel() == 0:
        return 0, 1
    counts = torch.bincount(batch_ids.reshape(-1))
    tot, mx = torch.stack([(counts * counts).sum(), counts.max()]).tolist()
    return int(tot), max(int(mx), 1)


class SceneContextTrain:
    """Scene ranges + the compact same-scene pair list (one host sync per batch for its size)."""

    def __init__(self, batch_ids: torch.Tensor, centres: torch.Tensor):
        self.seg_start, self.seg_end, self.err_flag = ops.scene_ranges(batch_ids)
        sizes = (self.seg_end - self.seg_start).to(torch.int64)          # index bookkeeping
        ends = torch.cumsum(sizes, 0)
        self.pair_off = (ends - sizes).contiguous()
        n = sizes.numel()
        if _scene_hint is not None:
            tot, mx = _scene_hint
        elif n:
            tot, mx, bad = torch.stack([ends[-1], sizes.max(), self.err_flag[0].to(torch.int64)]).tolist()      # the one host sync
            if bad:
                raise RuntimeError("vlsat_b200: batch_ids must be non-decreasing scene ids (src/dataset/DataLoader.py:153-176)")
        else:
            tot, mx = 0, 1
        self.n_pairs, self.max_scene = int(tot), max(int(mx), 1)
        self.centres = centres.detach().contiguous()
        self.pair_feats = ops.pair_features(self.centres, self.seg_start, self.seg_end, self.pair_off, self.n_pairs)


def distance_bias(fc: torch.nn.Sequential, sctx: SceneContextTrain) -> torch.Tensor:
    """self_attn_fc on every same-scene pair (network_MMG.py:165-173,189-203): [pairs, H]."""
    h = A.linear(sctx.pair_feats, fc[0].weight, fc[0].bias, RELU)
    h = A.add_layernorm(h, None, fc[2].weight, fc[2].bias, fc[2].eps)
    h = A.linear(h, fc[3].weight, fc[3].bias, RELU)
    h = A.add_layernorm(h, None, fc[5].weight, fc[5].bias, fc[5].eps)
    return A.linear(h, fc[6].weight, fc[6].bias)


def _mha_finish(mha, q_in, att, relu_out: bool = False):
    a = mha.attention
    o = A.linear(att, a.fc_o.weight, a.fc_o.bias)
    o = A.dropout(o, mha.dropout.p, mha.training)                        # attention.py:121
    ln = mha.layer_norm
    return A.add_layernorm(o, q_in, ln.weight, ln.bias, ln.eps, relu_out)


def mha_scenes(mha, q_in, kv_in, bias, sctx: SceneContextTrain):
    a = mha.attention
    q = A.linear(q_in, a.fc_q.weight, a.fc_q.bias)
    k = A.linear(kv_in, a.fc_k.weight, a.fc_k.bias)
    v = A.linear(kv_in, a.fc_v.weight, a.fc_v.bias)
    return _mha_finish(mha, q_in, A.node_attn(q, k, v, bias, sctx, a.h))


def mha_all(mha, q_in, kv_in, relu_out: bool = False):
    a = mha.attention
    q = A.linear(q_in, a.fc_q.weight, a.fc_q.bias)
    k = A.linear(kv_in, a.fc_k.weight, a.fc_k.bias)
    v = A.linear(kv_in, a.fc_v.weight, a.fc_v.bias)
    return _mha_finish(mha, q_in, A.flash_attn(q, k, v, a.h), relu_out)


# ------------------------------------------------------------------------------------ graph attention
def _head_major(lin: torch.nn.Linear, d: int, H: int):
    """Rows of a projection permuted so that the interleaved layout f = c*H + h (``.view(E, d, H)``,
    network_MMG.py:97-98) becomes contiguous per head: row h*d + c <- row c*H + h. Views of the parameters."""
    w = lin.weight.view(d, H, lin.weight.shape[1]).permute(1, 0, 2).reshape(H * d, lin.weight.shape[1])
    b = lin.bias.view(d, H).t().reshape(H * d)
    return w, b


def edge_attention(m, x, edge, g: GraphContext, x_value=None, aggr=None):
    """MultiHeadedEdgeAttention on CSR-ordered edges: (xx [N, D_a] aggregated, new edge feature [E, D_e], prob).
    ``x_value``: separate node set for the destination side (the reference-signature entry point passes already
    gathered per-edge query / value rows and an identity graph)."""
    xv = x if x_value is None else x_value
    H, dn, de, do = m.num_heads, m.d_n, m.d_e, m.d_o
    Dn, De = m.dim_node, m.dim_edge
    E, N = g.num_edges, g.num_nodes
    tr = m.training
    # nn_edge on cat[x_i, e, x_j]: project per node, gather per edge
    w1, w2 = m.nn_edge[0], m.nn_edge[2]
    w1_src, w1_edge, w1_dst = A.split_cols(w1.weight, (Dn, De, w1.weight.shape[1] - Dn - De))    # one backward for the three blocks
    a_src = A.linear(x, w1_src)
    b_dst = A.linear(xv, w1_dst)
    h1 = A.linear(edge, w1_edge, w1.bias, RELU, gather=(a_src, g.src, b_dst, g.dst), emit_pair=True)
    new_edge = A.linear(h1, w2.weight, w2.bias, emit_pair=True)           # read next by the edge cross-attention's q / k / v projections
    # attention MLP over rows (e, h)
    convs = m._convs()
    c1, c2 = convs[0].weight.squeeze(-1), convs[1].weight.squeeze(-1)
    wq, bq = _head_major(m.proj_query[0], dn, H)
    wv, bv = _head_major(m.proj_value[0], do, H)
    q_hm = A.linear(x, wq, bq, emit_pair=True)                            # [N, H*d_n]
    v_hm = A.linear(xv, wv, bv)                                           # [N, H*d_o]
    # (node, head) / (edge, head) rows are views of the head-major projections; ops.view_rows carries the emitted pair along
    c1_q, c1_k = A.split_cols(c1, (dn, c1.shape[1] - dn)) if c1.shape[1] > dn else (c1, None)
    qc = A.linear(ops.view_rows(q_hm, N * H, dn), c1_q, convs[0].bias)         # [N*H, hid]: C1[:, :d_n] q + c1
    rows_q = g.head_rows(H)                                               # row (e, h) -> row src(e)*H + h
    drop = next((mod.p for mod in m.nn if isinstance(mod, torch.nn.Dropout)), 0.0)
    dropping = tr and drop > 0.0                                          # then the pair of `hidden` leaves with the dropout pass
    if m.use_edge:
        wk, bk = _head_major(m.proj_edge[0], de, H)
        k_hm = A.linear(edge, wk, bk, emit_pair=True)                     # [E, H*d_e]
        hidden = A.linear(ops.view_rows(k_hm, E * H, de), c1_k, None, RELU, gather=(qc, rows_q, None, None), emit_pair=not dropping)
    else:
        hidden = A.relu(A.gather_rows(qc, rows_q))                        # MLP [d_n, 2 d_n, d_o] on the query alone
    hidden = A.dropout(hidden, drop, tr, emit_pair=True)
    t = A.linear(hidden, c2, convs[1].bias)                               # [E*H, d_o]
    xx, prob = A.gat_softmax_aggr(t, v_hm, g, H, aggr or getattr(m, "_aggr", "max"))
    return xx, new_edge, prob


def gat_layer(layer, x, edge, g: GraphContext, relu_nodes: bool = False):
    """GraphEdgeAttenNetwork.forward on CSR-ordered edges (network_MMG.py:34-41)."""
    xx, new_edge, prob = edge_attention(layer.edgeatten, x, edge, g)
    p0, p2 = layer.prop[0], layer.prop[2]
    hid = A.linear(torch.cat([x, xx], 1), p0.weight, p0.bias, RELU, emit_pair=True)
    out = A.linear(hid, p2.weight, p2.bias, RELU if relu_nodes else NONE)
    return out, new_edge, prob


def mmg_forward(m, o3, o2, e3, e2, edge_index, batch_ids, obj_center):
    """MMG.forward (network_MMG.py:178-250)."""
    n = o3.shape[0]
    sctx = SceneContextTrain(batch_ids, obj_center)
    bias = distance_bias(m.self_attn_fc, sctx)
    g = GraphContext(edge_index, n, m.flow)
    e3 = A.permute_rows(e3.contiguous(), g.perm, True, emit_pair=True)    # both feed the layer's edge projections
    e2 = A.permute_rows(e2.contiguous(), g.perm, True, emit_pair=True)
    p = m.drop_out.p
    for i in range(m.depth):
        act = (i < m.depth - 1) or m.depth == 1
        o3 = mha_scenes(m.self_attn[i], o3, o3, bias, sctx)
        o2 = mha_scenes(m.cross_attn[i], o2, o3, bias, sctx)
        o3, e3, _ = gat_layer(m.gcn_3ds[i], o3, e3, g, relu_nodes=act)
        o2, e2, _ = gat_layer(m.gcn_2ds[i], o2, e2, g, relu_nodes=act)
        e2 = mha_all(m.cross_attn_rel[i], e2, e3)
        if act:
            dropping = m.training and p > 0.0
            e3, e2 = A.relu(e3, not dropping), A.relu(e2, not dropping)   # the node streams got theirs in the epilogue
            # the next layer's projections read these: their bf16 pairs leave with the dropout pass
            o3, o2 = A.dropout(o3, p, m.training, emit_pair=True), A.dropout(o2, p, m.training, emit_pair=True)
            e3, e2 = A.dropout(e3, p, m.training, emit_pair=True), A.dropout(e2, p, m.training, emit_pair=True)
    return o3, o2, A.permute_rows(e3, g.perm, False, emit_pair=True), A.permute_rows(e2, g.perm, False, emit_pair=True)


def gnn_layers_forward(m, node, edge, edge_index, obj_center, batch_ids):
    """GraphEdgeAttenNetworkLayers.forward (network_GNN.py:227-284)."""
    n = node.shape[0]
    sctx = SceneContextTrain(batch_ids, obj_center)
    bias = distance_bias(m.self_attn_fc, sctx)
    g = GraphContext(edge_index, n, m.flow)
    edge = A.permute_rows(edge.contiguous(), g.perm, True, emit_pair=True)
    p = m.drop_out.p if m.drop_out is not None else 0.0
    probs = []
    for i in range(m.num_layers):
        act = (i < m.num_layers - 1) or m.num_layers == 1
        node = mha_scenes(m.self_attn[i], node, node, bias, sctx)
        node, edge, prob = gat_layer(m.gconvs[i], node, edge, g, relu_nodes=act)
        if act:
            edge = A.relu(edge, not (m.training and p > 0.0))
            node, edge = A.dropout(node, p, m.training, emit_pair=True), A.dropout(edge, p, m.training, emit_pair=True)
        H, do = m.gconvs[i].edgeatten.num_heads, m.gconvs[i].edgeatten.d_o
        E = g.num_edges
        pr = prob.view(E, H, do).permute(0, 2, 1).contiguous()           # [E, d_o, H] as the reference returns it
        if E:
            pr = ops.permute_rows(pr.view(E, -1), g.perm, gather=False).view(E, do, H)
        probs.append(pr.cpu().detach() if m.probs_on_host else pr)
    return node, A.permute_rows(edge, g.perm, False), probs


# --------------------------------------------------------------------------------------- the full model
def mmgnet_forward(m, obj_points, obj_2d_feats, edge_indices, descriptor, batch_ids, istrain, use_spatial=True):
    """Mmgnet.forward (SGFN_MMG/model.py:288-335)."""
    n = obj_points.shape[0]
    tr = m.training
    obj_feature = pointnet_feat(m.obj_encoder, obj_points)                       # [N, 768]
    mimic3d = obj_feature[..., :512].clone() if istrain else None
    lin, bn = m.mlp_3d[0], m.mlp_3d[1]
    h = A.linear(obj_feature, lin.weight, lin.bias)
    h = A.batchnorm(h, bn, tr, relu_out=True)
    h = A.dropout(h, m.mlp_3d[3].p, tr)
    if use_spatial:
        tail = torch.empty((n, 8), device=h.device, dtype=torch.float32)
        ops.spatial_tail(descriptor.contiguous(), tail, 0)
        node3d = torch.cat([h, tail], 1)
    else:
        node3d = h
    edge_feature = ops.edge_descriptor(descriptor.contiguous(), edge_indices.contiguous()).unsqueeze(-1)   # no grad (model.py:302)
    rel2 = pointnet_feat(m.rel_encoder_2d, edge_feature)
    rel3 = pointnet_feat(m.rel_encoder_3d, edge_feature)
    with torch.no_grad():                                                       # frozen adapter (model.py:309-310)
        obj_2d = m.clip_adapter(obj_2d_feats.contiguous())
    mimic2d = obj_2d.clone() if istrain else None
    centre = descriptor[:, :3].contiguous()
    g3, g2, ge3, ge2 = mmg_forward(m.mmg, node3d, obj_2d, rel3, rel2, edge_indices, batch_ids, centre)

    dis = None
    if istrain:
        p0, p3 = m.triplet_projector_2d[0], m.triplet_projector_2d[3]
        src, dst = edge_indices[0].contiguous(), edge_indices[1].contiguous()
        p0_src, p0_dst, p0_edge = A.split_cols(p0.weight, (512, 512, p0.weight.shape[1] - 1024))
        a = A.linear(g2, p0_src)
        b = A.linear(g2, p0_dst)
        hh = A.linear(ge2, p0_edge, p0.bias, RELU, gather=(a, src, b, dst))                  # Linear Dropout ReLU: commute
        hh = A.dropout(hh, m.triplet_projector_2d[1].p, tr, emit_pair=True)
        dis = A.linear(hh, p3.weight, p3.bias)
    rel_cls_3d = rel_classifier(m.rel_predictor_3d, ge3)
    rel_cls_2d = rel_classifier(m.rel_predictor_2d, ge2)
    scale = m.obj_logit_scale.reshape(1)
    logits3 = A.linear(A.row_l2norm(g3), m.obj_predictor_3d.weight, m.obj_predictor_3d.bias, scale=scale)
    logits2 = A.linear(A.row_l2norm(g2), m.obj_predictor_2d.weight, m.obj_predictor_2d.bias, scale=scale)
    if istrain:
        return logits3, logits2, rel_cls_3d, rel_cls_2d, mimic3d, mimic2d, dis, m.obj_logit_scale.exp()
    return logits3, logits2, rel_cls_3d, rel_cls_2d
