"""Stage-by-stage gradient comparison of MMG.forward (CUDA differentiable path vs the oracle in float64). Debug aid.
usage: python tests/dev_grad_check_mmg.py [out_index 0..3] [scale_2d]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
import vlsat_b200 as V
from oracle import vlsat_oracle as O
from vlsat_b200 import autograd as A, train_path as T
from vlsat_b200.gat import GraphContext

which = int(sys.argv[1]) if len(sys.argv) > 1 else 1
s2 = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
model = V.Mmgnet(cases.model_config({}), 160, 26)
model.load_state_dict(cases.seeded_state(model, 0))
m = model.mmg
sd = {k: v.clone().double().requires_grad_(True) for k, v in m.state_dict().items()}
b = cases.MMGNET_CASES["mmgnet_cfg1"][1]()
N, E = b.descriptor.shape[0], b.edge_indices.shape[1]
g = torch.Generator().manual_seed(3)
o3, o2 = torch.randn(N, 512, generator=g), torch.randn(N, 512, generator=g) * s2
e3, e2 = torch.randn(E, 512, generator=g).relu(), torch.randn(E, 512, generator=g).relu()
centres = b.descriptor[:, :3].contiguous()
R = torch.randn(max(N, E), 512, generator=g)

def run(fns, o3, o2, e3, e2, dev, dt):
    st = {}
    def keep(name, t):
        t.retain_grad(); st[name] = t; return t
    o3, o2, e3, e2 = (keep(n, t.to(dev, dt).requires_grad_(True)) for n, t in (("in.o3", o3), ("in.o2", o2), ("in.e3", e3), ("in.e2", e2)))
    for i in range(2):
        act = i < 1
        o3 = keep(f"L{i}.o3_sa", fns["sa"](i, o3))
        o2 = keep(f"L{i}.o2_ca", fns["ca"](i, o2, o3))
        o3, e3 = fns["g3"](i, o3, e3); keep(f"L{i}.o3_gcn", o3); keep(f"L{i}.e3_gcn", e3)
        o2, e2 = fns["g2"](i, o2, e2); keep(f"L{i}.o2_gcn", o2); keep(f"L{i}.e2_gcn", e2)
        e2 = keep(f"L{i}.e2_x", fns["x"](i, e2, e3))
        if act:
            o3, o2, e3, e2 = (keep(f"L{i}.relu{j}", fns["relu"](t)) for j, t in enumerate((o3, o2, e3, e2)))
    outs = (o3, o2, e3, e2)
    (outs[which] * R[:outs[which].shape[0]].to(dev, dt)).sum().backward()
    return st

mask, bias = O.distance_bias(sd, "self_attn_fc.", centres.double(), b.batch_ids, 8)
ref = run(dict(sa=lambda i, o3: O.mha(sd, f"self_attn.{i}.", o3, o3, o3, 8, mask, bias),
               ca=lambda i, o2, o3: O.mha(sd, f"cross_attn.{i}.", o2, o3, o3, 8, mask, bias),
               g3=lambda i, o, e: O.gat_layer(sd, f"gcn_3ds.{i}.", o, e, b.edge_indices, 8)[:2],
               g2=lambda i, o, e: O.gat_layer(sd, f"gcn_2ds.{i}.", o, e, b.edge_indices, 8)[:2],
               x=lambda i, e2, e3: O.mha(sd, f"cross_attn_rel.{i}.", e2, e3, e3, 8), relu=torch.relu), o3, o2, e3, e2, "cpu", torch.float64)
m = m.cuda().eval()
sctx = T.SceneContextTrain(b.batch_ids.cuda(), centres.cuda())
bias_c = T.distance_bias(m.self_attn_fc, sctx)
gc = GraphContext(b.edge_indices.cuda(), N, m.flow)
assert torch.equal(gc.perm.cpu(), torch.arange(E, dtype=torch.int32)), "cfg1 edges are already CSR-sorted"
got = run(dict(sa=lambda i, o3: T.mha_scenes(m.self_attn[i], o3, o3, bias_c, sctx),
               ca=lambda i, o2, o3: T.mha_scenes(m.cross_attn[i], o2, o3, bias_c, sctx),
               g3=lambda i, o, e: T.gat_layer(m.gcn_3ds[i], o, e, gc)[:2],
               g2=lambda i, o, e: T.gat_layer(m.gcn_2ds[i], o, e, gc)[:2],
               x=lambda i, e2, e3: T.mha_all(m.cross_attn_rel[i], e2, e3), relu=A.relu), o3, o2, e3, e2, "cuda", torch.float32)
print(f"loss on output {which}, 2-D input scale {s2}")
for k in ref:
    r, q = ref[k], got[k]
    fe = ((q.detach().cpu().double() - r.detach()).norm() / (r.detach().norm() + 1e-30)).item()
    ge = float("nan")
    if r.grad is not None and q.grad is not None:
        ge = ((q.grad.cpu().double() - r.grad).norm() / (r.grad.norm() + 1e-30)).item()
    print(f"{k:12s} fwd rel err {fe:9.2e}   grad rel err {ge:9.2e}   |grad| {0.0 if r.grad is None else r.grad.norm().item():9.2e}")
rows = []
for k, p in m.named_parameters():
    if sd[k].grad is None or p.grad is None:
        continue
    r = sd[k].grad
    rows.append((((p.grad.cpu().double() - r).norm() / (r.norm() + 1e-30)).item(), r.norm().item(), k))
for rel, nr, k in sorted(rows, reverse=True):
    if nr > 1e-9:
        print(f"{rel:10.3e}  |ref|={nr:10.3e}  {k}")
