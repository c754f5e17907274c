"""Pin the oracle (oracle/vlsat_oracle.py) against (a) the committed golden vectors that the UNMODIFIED
reference produced (oracle/make_golden.py) and (b) the live reference when /root/reference exists."""
import pytest
import torch

import cases
import vlsat_b200 as V
from conftest import assert_close
from oracle import ref_shims, vlsat_oracle as O
from vlsat_b200 import synth


def _mmgnet_state(overrides):
    net = V.Mmgnet(cases.model_config(overrides), 160, 26)
    return cases.seeded_state(net, cases.MMGNET_WEIGHT_SEED)


@pytest.mark.parametrize("name", list(cases.MMGNET_CASES))
def test_oracle_matches_reference_golden_full_model(name, golden):
    over, make = cases.MMGNET_CASES[name]
    sd = _mmgnet_state(over)
    b = make()
    cfg = cases.model_config(over)["MODEL"]
    with torch.no_grad():
        outs, inter = O.mmgnet_forward(sd, *b.forward_args(), istrain=True, depth=cfg["N_LAYERS"],
                                       num_heads=cfg["NUM_HEADS"], aggr=cfg["GCN_AGGR"], return_intermediates=True)
    g = golden(name)
    for i, (a, e) in enumerate(zip(outs[:7], g["train"][:7])):
        assert_close(a, e, f"{name} train output {i}", rtol=1e-4, atol=2e-5)
    for i, (a, e) in enumerate(zip(outs[:4], g["eval"])):
        assert_close(a, e, f"{name} eval output {i}", rtol=1e-4, atol=2e-5)
    if g["inter"]:
        assert_close(inter["obj_encoder"], g["inter"]["obj_encoder"][0], "obj_encoder", rtol=1e-4)
        assert_close(inter["rel_feature_3d"], g["inter"]["rel_encoder_3d"][0], "rel_encoder_3d", rtol=1e-4)
        assert_close(inter["obj_feature_2d"], g["inter"]["clip_adapter"][0], "clip_adapter", rtol=1e-4)
        for k, j in (("gcn_obj_3d", 0), ("gcn_obj_2d", 1), ("gcn_edge_3d", 2), ("gcn_edge_2d", 3)):
            assert_close(inter[k], g["inter"]["mmg"][j], k, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("name", list(cases.GAT_CASES))
def test_oracle_matches_reference_golden_gat_layer(name, golden):
    kw = dict(cases.GAT_CASES[name][0])
    seed = cases.GAT_CASES[name][4]
    layer = V.GraphEdgeAttenNetwork(**kw)
    sd = cases.seeded_state(layer, seed)
    x, ef, ei = cases.gat_inputs(name)
    with torch.no_grad():
        xo, eo, prob = O.gat_layer(sd, "", x, ef, ei, kw["num_heads"], kw["aggr"], kw.get("flow", "target_to_source"),
                                   kw.get("use_edge", True))
    g = golden("gat_layers")[name]
    assert_close(xo, g["x"], name + " x", rtol=1e-4)
    assert_close(eo, g["e"], name + " e", rtol=1e-4)
    assert_close(prob, g["prob"], name + " prob", rtol=1e-4)


def test_oracle_matches_reference_golden_gnn_layers(golden):
    net = V.GraphEdgeAttenNetworkLayers(**cases.GNN_CASE)
    sd = cases.seeded_state(net, 23)
    node, edge, ei, centres, bids = cases.gnn_inputs()
    with torch.no_grad():
        n, e, probs = O.gnn_layers_forward(sd, "", node, edge, ei, centres, bids, 2, 8)
    g = golden("gnn_layers")
    assert_close(n, g["node"], "node", rtol=1e-4, atol=2e-5)
    assert_close(e, g["edge"], "edge", rtol=1e-4, atol=2e-5)
    for a, b in zip(probs, g["probs"]):
        assert_close(a, b, "prob", rtol=1e-4)


@pytest.mark.parametrize("name", list(cases.POINTNET_CASES))
def test_oracle_matches_reference_golden_pointnet(name, golden):
    kw, n, p, seed = cases.POINTNET_CASES[name]
    enc = V.PointNetfeat(global_feat=True, batch_norm=False, input_transform=False, feature_transform=False, **kw)
    sd = cases.seeded_state(enc, seed)
    with torch.no_grad():
        out = O.pointnet_feat(sd, "", cases.pointnet_inputs(name))
    assert_close(out, golden("pointnet")[name], name, rtol=1e-4)


@pytest.mark.parametrize("name", list(cases.MHA_CASES))
def test_oracle_matches_reference_golden_mha(name, golden):
    d, h, nq, nk, seed = cases.MHA_CASES[name]
    att = V.MultiHeadAttention(d_model=d, d_k=d // h, d_v=d // h, h=h)
    sd = cases.seeded_state(att, seed)
    q, kv = cases.mha_inputs(name)
    with torch.no_grad():
        out = O.mha(sd, "", q, kv, kv, h)
    assert_close(out, golden("mha")[name], name, rtol=1e-4, atol=2e-5)


def test_edge_descriptor_and_index_demo(golden):
    g = golden("edge_descriptor")
    b = cases.MMGNET_CASES["mmgnet_ragged"][1]()
    assert_close(O.edge_descriptor(b.descriptor, b.edge_indices), g["ragged"], "edge descriptor", rtol=1e-5)
    # hand-derived known answers of the reference's in-file demo (network_util.py:75-99):
    x = torch.zeros(3, 5); x[1] = 1; x[2] = 2
    ei = torch.LongTensor([[0, 1, 2], [2, 1, 0]])
    xi, xj = O.gather_pairs(x, ei, "source_to_target")
    assert xi[:, 0].tolist() == [2, 1, 0] and xj[:, 0].tolist() == [0, 1, 2]
    tmp = -torch.arange(5.).view(5, 1).repeat(1, 2)
    ei2 = torch.LongTensor([[0, 1, 2, 1, 0], [2, 1, 1, 1, 1]])
    xx = O.aggregate(tmp, ei2, 3, "max", "source_to_target")
    assert xx[:, 0].tolist() == [0.0, -1.0, 0.0]          # node 0 empty -> 0, node 1 max(-1..-4), node 2 <- edge 0
    for flow in ("source_to_target", "target_to_source"):
        d = g["demo"][flow]
        a, c = O.gather_pairs(x, ei, flow)
        assert torch.equal(a, d["x_i"]) and torch.equal(c, d["x_j"])
        assert torch.equal(O.aggregate(tmp, ei2, 3, "max", flow), d["xx"])


def test_csr_restatement_is_a_stable_sort():
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, 17, (200,), generator=g)
    row_ptr, perm = O.build_csr(idx, 20)
    assert row_ptr[0] == 0 and row_ptr[-1] == 200
    assert torch.equal(idx[perm], torch.sort(idx)[0])
    for n in range(20):
        seg = perm[row_ptr[n]:row_ptr[n + 1]]
        assert torch.equal(seg, torch.sort(seg)[0])
        assert (idx[seg] == n).all()


@pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference_on_fresh_inputs():
    net, _ = ref_shims.build_reference_mmgnet(seed=0)
    synth.load_seeded(net, 7)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    b = synth.make_batch(2, [6, 4], 48, 10, seed=77, shuffle_edges=True)
    with torch.no_grad():
        ref = net(*b.forward_args(), istrain=True)
        mine = O.mmgnet_forward(sd, *b.forward_args(), istrain=True)
    for i, (a, e) in enumerate(zip(mine[:7], ref[:7])):
        assert_close(a, e, f"live output {i}", rtol=1e-4, atol=2e-5)
