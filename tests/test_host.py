"""Host logic that needs no GPU: state_dict compatibility, scene generator, sharding, error behaviour."""
import json
import os

import pytest
import torch

import cases
import vlsat_b200 as V
from oracle import vlsat_oracle as O
from vlsat_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_schema_equals_reference():
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_schema.json")))
    mine = {k: list(v.shape) for k, v in V.Mmgnet(cases.model_config({}), 160, 26).state_dict().items()}
    assert mine == ref                      # 197 entries, same names, same shapes, same order
    assert list(mine) == list(ref)


def test_parameter_count_and_frozen_adapter():
    m = V.Mmgnet(cases.model_config({}), 160, 26)
    assert sum(p.numel() for p in m.parameters()) == 27424934
    assert sum(p.numel() for p in m.parameters() if not p.requires_grad) == 262913


def test_gnn_layers_schema():
    net = V.GraphEdgeAttenNetworkLayers(**cases.GNN_CASE)
    sd = net.state_dict()
    assert tuple(sd["gconvs.0.edgeatten.nn_edge.0.weight"].shape) == (768, 1280)
    assert tuple(sd["gconvs.1.edgeatten.nn.3.weight"].shape) == (32, 96, 1)
    assert tuple(sd["self_attn_fc.6.weight"].shape) == (8, 32)


def test_unsupported_configs_fail_loudly():
    with pytest.raises(NotImplementedError):
        V.Mmgnet(cases.model_config(dict(WITH_BN=True)), 160, 26)
    with pytest.raises(NotImplementedError):
        V.Mmgnet(cases.model_config(dict(ATTENTION="distance")), 160, 26)
    with pytest.raises(NotImplementedError):
        V.GraphEdgeAttenNetwork(8, 512, 512, 256, aggr="min", DROP_OUT_ATTEN=0.5)
    with pytest.raises(RuntimeError):
        V.Mmgnet({"MODEL": {"N_LAYERS": 2}}, 160, 26)            # missing key -> RuntimeError like config.py:51
    with pytest.raises(AssertionError):
        V.MultiHeadedEdgeAttention(num_heads=3, dim_node=512, dim_edge=512, dim_atten=256)


def test_no_cpu_fallback():
    m = V.Mmgnet(cases.model_config({}), 160, 26).eval()
    b = synth.make_config_batch("cfg1")
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):      # fused inference path
        m(*b.forward_args())
    with pytest.raises(RuntimeError, match="CUDA"):                       # differentiable path (grad mode on)
        m(*b.forward_args())
    m.train()
    with pytest.raises(RuntimeError, match="CUDA"):                       # training mode
        m(*b.forward_args(), istrain=True)


def test_synthetic_batch_layout():
    b = synth.make_config_batch("cfg2", seed=3)
    assert tuple(b.obj_points.shape) == (640, 3, 256) and tuple(b.edge_indices.shape) == (2, 9600)
    assert b.edge_indices.dtype == torch.int64 and b.batch_ids.dtype == torch.int64
    assert tuple(b.batch_ids.shape) == (640, 1) and (b.batch_ids[1:] >= b.batch_ids[:-1]).all()
    assert b.obj_points.mean(2).abs().max() < 1e-4                      # zero-mean per object
    assert (b.descriptor[:, 6:] > 0).all()                               # logs are defined
    src, dst = b.edge_indices
    assert (src != dst).all() and (b.batch_ids[src] == b.batch_ids[dst]).all()
    assert (src[1:] >= src[:-1]).all()                                   # product order: sorted by subject
    full = synth.make_config_batch("cfg3", num_scenes=1)
    assert full.edge_indices.shape[1] == 40 * 39
    b2 = synth.make_config_batch("cfg2", seed=3)
    assert all(torch.equal(x, y) for x, y in zip(b.tensors(), b2.tensors()))


def test_scene_sharding_partitions_the_batch():
    b = synth.make_real_shaped_batch(7, seed=1, points_per_object=16)
    shards = [synth.shard_scenes(b, r, 2) for r in range(2)]
    assert sum(s.num_scenes for s in shards) == 7
    assert sum(s.obj_points.shape[0] for s in shards) == b.obj_points.shape[0]
    assert sum(s.edge_indices.shape[1] for s in shards) == b.edge_indices.shape[1]
    for s in shards:
        assert s.batch_ids.min() == 0 and s.batch_ids.max() == s.num_scenes - 1
        assert s.edge_indices.max() < s.obj_points.shape[0]
        assert (s.batch_ids.view(-1)[s.edge_indices[0]] == s.batch_ids.view(-1)[s.edge_indices[1]]).all()
    # forward needs no communication: the oracle on a shard equals the oracle on the full batch restricted
    # to that shard's scenes for every scene-local quantity (node attention + GAT are scene-local)
    layer = V.GraphEdgeAttenNetwork(4, 64, 32, 32, DROP_OUT_ATTEN=0.5)
    sd = cases.seeded_state(layer, 1)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(b.obj_points.shape[0], 64, generator=g)
    e = torch.randn(b.edge_indices.shape[1], 32, generator=g)
    full_x, full_e, _ = O.gat_layer(sd, "", x, e, b.edge_indices, 4)
    bids = b.batch_ids.view(-1)
    for r, s in enumerate(shards):
        keep_n = torch.isin(bids, torch.arange(r, 7, 2))
        keep_e = keep_n[b.edge_indices[0]]
        sx, se, _ = O.gat_layer(sd, "", x[keep_n], e[keep_e], s.edge_indices, 4)
        assert torch.allclose(sx, full_x[keep_n], atol=1e-6) and torch.allclose(se, full_e[keep_e], atol=1e-6)


def test_seeded_state_is_order_independent():
    a = synth.seeded_tensor("mmg.gcn_3ds.0.prop.0.weight", (768, 768), 0)
    b = synth.seeded_tensor("mmg.gcn_3ds.0.prop.0.weight", (768, 768), 0)
    c = synth.seeded_tensor("mmg.gcn_3ds.1.prop.0.weight", (768, 768), 0)
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_derived_cache_invalidation():
    from vlsat_b200._cache import DerivedCache
    p = torch.nn.Parameter(torch.ones(3))
    c = DerivedCache()
    calls = []
    f = lambda: calls.append(1) or (p.detach() * 2)
    c.get("k", (p,), f); c.get("k", (p,), f)
    assert len(calls) == 1
    with torch.no_grad():
        p.add_(1)
    assert torch.equal(c.get("k", (p,), f), torch.full((3,), 4.0)) and len(calls) == 2


def test_adopt_parameters_shares_storage():
    a = V.GraphEdgeAttenNetwork(4, 64, 32, 32, DROP_OUT_ATTEN=0.5)
    b = V.GraphEdgeAttenNetwork(4, 64, 32, 32, DROP_OUT_ATTEN=0.5)
    V.adopt_parameters(b, a)
    for (n1, p1), (n2, p2) in zip(a.named_parameters(), b.named_parameters()):
        assert n1 == n2 and p1 is p2


def test_loader_hook_reproduces_the_reference_sampling_stream():
    """N2 loader hook (host side): the indices drawn by data_prep.sample_object_indices select exactly the rows the
    reference's loop selects (dataset_3dssg.py:285-290: points[np.where(instances == id)[0]][choice]) under the same seed."""
    import numpy as np
    from vlsat_b200 import data_prep
    rs = np.random.RandomState(3)
    m = 5000
    points = rs.randn(m, 6).astype(np.float32)
    instances = rs.randint(1, 9, size=m)
    nodes = [3, 1, 7, 5]
    np.random.seed(11)
    want = []
    for instance_id in nodes:                                   # the reference's loop, verbatim semantics
        obj_pointset = points[np.where(instances == instance_id)[0]]
        choice = np.random.choice(len(obj_pointset), 64, replace=True)
        want.append(obj_pointset[choice, :])
    np.random.seed(11)
    idx = data_prep.sample_object_indices(instances, nodes, 64)
    assert idx.shape == (4, 64) and idx.dtype == np.int64
    for i in range(4):
        assert np.array_equal(points[idx[i]], want[i])
    with pytest.raises(ValueError):
        data_prep.sample_object_indices(instances, [99], 64)


def test_split_cols_is_slicing_with_one_backward():
    """autograd.split_cols: the blocks are views of the weight (no copy), the gradient equals that of plain slicing, unused
    blocks get zeros; without autograd it is a plain split."""
    from vlsat_b200 import autograd as A
    w = torch.randn(6, 10, requires_grad=True)
    x0, x1, x2 = torch.randn(4, 3), torch.randn(4, 5), torch.randn(4, 2)
    a, b, c = A.split_cols(w, (3, 5, 2))
    assert a.data_ptr() == w.data_ptr() and b.data_ptr() == w[:, 3:].data_ptr() and tuple(c.shape) == (6, 2)
    ((x0 @ a.t()).sum() * 2 + (x2 @ c.t()).pow(2).sum()).backward()          # block b unused
    got = w.grad.clone()
    w.grad = None
    ((x0 @ w[:, :3].t()).sum() * 2 + (x2 @ w[:, 8:].t()).pow(2).sum()).backward()
    assert torch.allclose(got, w.grad) and torch.count_nonzero(got[:, 3:8]) == 0
    with torch.no_grad():
        p, q = A.split_cols(w, (4, 6))
        assert p.grad_fn is None and p.data_ptr() == w.data_ptr() and tuple(q.shape) == (6, 6)
    with pytest.raises(ValueError):
        A.split_cols(w, (4, 4))


def test_view_rows_and_zero_arena_host_logic():
    """Host-side bookkeeping that needs no GPU: the (rows, head) view carries a remembered pair along as views of the same
    memory; the zero arena is inert off the GPU (plain torch.zeros, no state left behind)."""
    from vlsat_b200 import ops
    x = torch.randn(6, 16)
    hi, lo = torch.randn(6, 16).bfloat16(), torch.randn(6, 16).bfloat16()
    x._vlsat_pair = (x._version, (hi, lo))
    y = ops.view_rows(x, 12, 8)
    assert y.data_ptr() == x.data_ptr() and tuple(y.shape) == (12, 8)
    assert y._vlsat_pair[1][0].data_ptr() == hi.data_ptr() and tuple(y._vlsat_pair[1][1].shape) == (12, 8)
    x.add_(1.0)                                                  # a stale pair (version moved on) must not travel
    assert getattr(ops.view_rows(x, 12, 8), "_vlsat_pair", None) is None

    class Owner:
        pass
    owner = Owner()
    with ops.zero_arena(owner, "cpu") as arena:
        assert arena is None and ops._arena is None
        z = ops.zeros((3, 4), "cpu")
    assert z.shape == (3, 4) and torch.count_nonzero(z) == 0 and not hasattr(owner, "_zero_arena_bytes")
