"""TEST INFRASTRUCTURE: generate tests/golden/*.pt by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):  python oracle/make_golden.py
Weights and inputs are regenerated from seeds by tests/golden/cases.py, so only OUTPUTS are stored.
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import ref_shims  # noqa: E402
import cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name, obj):
    torch.save(obj, os.path.join(OUT, name + ".pt"))
    n = sum(v.numel() for v in _flat(obj))
    print(f"{name}: {n} values")


def _flat(o):
    if isinstance(o, torch.Tensor):
        yield o
    elif isinstance(o, dict):
        for v in o.values():
            yield from _flat(v)
    elif isinstance(o, (list, tuple)):
        for v in o:
            yield from _flat(v)


def main():
    ref_shims.install()
    from src.model.model_utils.network_MMG import GraphEdgeAttenNetwork
    from src.model.model_utils.network_GNN import GraphEdgeAttenNetworkLayers
    from src.model.model_utils.network_PointNet import PointNetfeat
    from src.model.transformer.attention import MultiHeadAttention
    from src.utils import op_utils
    from src.model.model_utils.network_util import Gen_Index, Aggre_Index

    # ---- full model ---------------------------------------------------------------------------------
    schema_written = False
    for name, (over, make) in cases.MMGNET_CASES.items():
        net, _ = ref_shims.build_reference_mmgnet(seed=0, overrides=over)
        net.load_state_dict(cases.seeded_state(net, cases.MMGNET_WEIGHT_SEED))
        net.eval()
        if not schema_written and not over:
            with open(os.path.join(OUT, "state_dict_schema.json"), "w") as f:
                json.dump({k: list(v.shape) for k, v in net.state_dict().items()}, f, indent=0)
            schema_written = True
        b = make()
        inter = {}
        hooks = []
        for mod_name in ("obj_encoder", "rel_encoder_3d", "rel_encoder_2d", "mlp_3d", "clip_adapter", "mmg",
                         "mmg.self_attn.0", "mmg.cross_attn.0", "mmg.gcn_3ds.0", "mmg.gcn_2ds.0", "mmg.cross_attn_rel.0"):
            mod = net.get_submodule(mod_name)
            hooks.append(mod.register_forward_hook(
                lambda m, i, o, key=mod_name: inter.__setitem__(key, [t.detach().clone() for t in (o if isinstance(o, tuple) else (o,))])))
        with torch.no_grad():
            ev = net(*b.forward_args(), istrain=False)
            for h in hooks:
                h.remove()
            tr = net(*b.forward_args(), istrain=True)
        keep_inter = name in ("mmgnet_cfg1", "mmgnet_ragged")
        save(name, dict(eval=[t.clone() for t in ev], train=[t.clone() for t in tr],
                        inter=inter if keep_inter else {}))

    # ---- graph attention layer (MMG flavour) --------------------------------------------------------------
    out = {}
    for name, (kw, n, e, iso, seed) in cases.GAT_CASES.items():
        layer = GraphEdgeAttenNetwork(**kw).eval()
        layer.load_state_dict(cases.seeded_state(layer, seed))
        x, ef, ei = cases.gat_inputs(name)
        with torch.no_grad():
            xo, eo = layer(x, ef, ei)
            xi, xj = layer.index_get(x, ei)
            msg, _, prob = layer.edgeatten(xi, ef, xj)
        out[name] = dict(x=xo, e=eo, prob=prob, msg=msg)
    save("gat_layers", out)

    # ---- SGFN twin -------------------------------------------------------------------------------------
    net = GraphEdgeAttenNetworkLayers(**cases.GNN_CASE).eval()
    net.load_state_dict(cases.seeded_state(net, 23))
    with torch.no_grad():
        node, edge, probs = net(*cases.gnn_inputs())
    save("gnn_layers", dict(node=node, edge=edge, probs=probs))

    # ---- PointNet -------------------------------------------------------------------------------------
    out = {}
    for name, (kw, n, p, seed) in cases.POINTNET_CASES.items():
        enc = PointNetfeat(global_feat=True, batch_norm=False, input_transform=False, feature_transform=False, **kw).eval()
        enc.load_state_dict(cases.seeded_state(enc, seed))
        with torch.no_grad():
            out[name] = enc(cases.pointnet_inputs(name))
    save("pointnet", out)

    # ---- attention (unmasked, the cross_attn_rel use) -------------------------------------------------
    out = {}
    for name, (d, h, nq, nk, seed) in cases.MHA_CASES.items():
        att = MultiHeadAttention(d_model=d, d_k=d // h, d_v=d // h, h=h).eval()
        att.load_state_dict(cases.seeded_state(att, seed))
        q, kv = cases.mha_inputs(name)
        with torch.no_grad():
            out[name] = att(q.unsqueeze(0), kv.unsqueeze(0), kv.unsqueeze(0)).squeeze(0)
    save("mha", out)

    # ---- edge descriptor + the in-file index demos (network_util.py:75-99) -----------------------------
    b = cases.MMGNET_CASES["mmgnet_ragged"][1]()
    with torch.no_grad():
        ed = op_utils.Gen_edge_descriptor(flow="target_to_source")(b.descriptor, b.edge_indices)
    x = torch.zeros(3, 5)
    x[1, :] = 1
    x[2, :] = 2
    demo = {}
    for flow in ("source_to_target", "target_to_source"):
        xi, xj = Gen_Index(flow=flow)(x, torch.LongTensor([[0, 1, 2], [2, 1, 0]]))
        tmp = torch.zeros(5, 2)
        for i in range(5):
            tmp[i] = -i
        xx = Aggre_Index(flow=flow, aggr="max")(tmp, torch.LongTensor([[0, 1, 2, 1, 0], [2, 1, 1, 1, 1]]), dim_size=3)
        demo[flow] = dict(x_i=xi, x_j=xj, xx=xx)
    save("edge_descriptor", dict(ragged=ed, demo=demo))


if __name__ == "__main__":
    main()
