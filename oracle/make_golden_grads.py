"""TEST INFRASTRUCTURE: gradient fixtures from the UNMODIFIED reference modules -> tests/golden/grads.pt.

Run in the build container only (needs /root/reference):  python oracle/make_golden_grads.py
For each case the reference runs forward + ``loss.backward()`` with loss = sum_i <out_i, R_i> (seeded cotangents,
tests/golden/cases.py) in two modes:
  * ``eval``  - module.eval(): dropout off, BatchNorm running statistics;
  * ``train`` - module.train() with every nn.Dropout probability forced to 0 (masks cannot be matched across RNGs):
                BatchNorm1d of mlp_3d uses batch statistics and updates its running buffers.
Stored per parameter: the whole gradient if small, else its first entries + sum + L2 norm (cases.grad_summary).
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import ref_shims  # noqa: E402
import cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "grads.pt")


def grads_of(module, outs, seed):
    module.zero_grad(set_to_none=True)
    cases.scalar_loss(outs, seed).backward()
    return {k: cases.grad_summary(p.grad) for k, p in module.named_parameters() if p.grad is not None}


def main():
    ref_shims.install()
    from src.model.model_utils.network_MMG import GraphEdgeAttenNetwork
    from src.model.model_utils.network_PointNet import PointNetfeat
    from src.model.transformer.attention import MultiHeadAttention

    out = {}
    for name in cases.GRAD_MMGNET_CASES:
        over, make = cases.MMGNET_CASES[name]
        for mode in ("eval", "train"):
            net, _ = ref_shims.build_reference_mmgnet(seed=0, overrides=over)
            net.load_state_dict(cases.seeded_state(net, cases.MMGNET_WEIGHT_SEED))
            net.train(mode == "train")
            for m in net.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            b = make()
            outs = net(*b.forward_args(), istrain=True)
            g = grads_of(net, outs[:7], seed=7)
            entry = dict(grads=g, outs=[o.detach().clone() for o in outs[:4]])
            if mode == "train":
                entry["bn_running_mean"] = net.mlp_3d[1].running_mean.detach().clone()
                entry["bn_running_var"] = net.mlp_3d[1].running_var.detach().clone()
            out[f"{name}.{mode}"] = entry
            print(name, mode, len(g), "parameter gradients")

    for name in cases.GRAD_GAT_CASES:
        kw, n, e, iso, seed = cases.GAT_CASES[name]
        layer = GraphEdgeAttenNetwork(**kw).eval()
        layer.load_state_dict(cases.seeded_state(layer, seed))
        x, ef, ei = cases.gat_inputs(name)
        x.requires_grad_(True); ef.requires_grad_(True)
        xo, eo = layer(x, ef, ei)
        g = grads_of(layer, [xo, eo], seed=8)
        g["input.x"], g["input.edge"] = cases.grad_summary(x.grad), cases.grad_summary(ef.grad)
        out[name] = dict(grads=g)
        print(name, len(g))

    for name in ("pointnet_obj", "pointnet_big", "pointnet_rgbn"):
        kw, n, p, seed = cases.POINTNET_CASES[name]
        enc = PointNetfeat(global_feat=True, batch_norm=False, input_transform=False, feature_transform=False, **kw).eval()
        enc.load_state_dict(cases.seeded_state(enc, seed))
        o = enc(cases.pointnet_inputs(name))
        out[name] = dict(grads=grads_of(enc, [o], seed=9))
        print(name)

    for name, (d, h, nq, nk, seed) in cases.MHA_CASES.items():
        att = MultiHeadAttention(d_model=d, d_k=d // h, d_v=d // h, h=h).eval()
        att.load_state_dict(cases.seeded_state(att, seed))
        q, kv = cases.mha_inputs(name)
        q = q.clone().requires_grad_(True)
        kv = q if name == "mha_self" else kv.clone().requires_grad_(True)
        o = att(q.unsqueeze(0), kv.unsqueeze(0), kv.unsqueeze(0)).squeeze(0)
        g = grads_of(att, [o], seed=10)
        g["input.q"] = cases.grad_summary(q.grad)
        if kv is not q:
            g["input.kv"] = cases.grad_summary(kv.grad)
        out[name] = dict(grads=g)
        print(name)

    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
