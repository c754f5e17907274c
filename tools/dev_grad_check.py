"""Per-parameter gradient error of the CUDA path against the oracle's autograd (debug aid).
usage: python tests/dev_grad_check.py [case] [eval|train]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
import vlsat_b200 as V
from oracle import vlsat_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "mmgnet_cfg1"
mode = sys.argv[2] if len(sys.argv) > 2 else "eval"
over, make = cases.MMGNET_CASES[name]
cfg = cases.model_config(over)
model = V.Mmgnet(cfg, 160, 26)
model.load_state_dict(cases.seeded_state(model, cases.MMGNET_WEIGHT_SEED))
sd = {k: (v.clone().double().requires_grad_(True) if v.is_floating_point() else v) for k, v in model.state_dict().items()}
for m in model.modules():
    if isinstance(m, torch.nn.Dropout):
        m.p = 0.0
model = model.to("cuda").train(mode == "train")
b = make()
mc = cfg["MODEL"]
args = [a.double() if a.is_floating_point() else a for a in b.forward_args()]
ref = O.mmgnet_forward(sd, *args, istrain=True, depth=mc["N_LAYERS"], num_heads=mc["NUM_HEADS"], aggr=mc["GCN_AGGR"], train_bn=(mode == "train"))
which = [int(i) for i in os.environ.get("OUTS", "0,1,2,3,4,5,6").split(",")]
got = model(*b.to("cuda").forward_args(), istrain=True)
ws = cases.loss_weights(ref[:7], 7)
sum((ref[i] * ws[i].double()).sum() for i in which).backward()
sum((got[i] * ws[i].cuda()).sum() for i in which).backward()
for i in range(7):
    e = (got[i].detach().cpu().double() - ref[i].detach()).abs().max().item()
    print(f"out{i}: max abs err {e:.3g} (scale {ref[i].detach().abs().max().item():.3g})")
rows = []
for k, p in model.named_parameters():
    r = sd[k].grad
    if r is None or p.grad is None:
        if (r is None) != (p.grad is None) and p.requires_grad:
            print("MISSING", k, r is None, p.grad is None)
        continue
    g = p.grad.detach().cpu().double()
    rel = ((g - r).norm() / (r.norm() + 1e-30)).item()
    rows.append((rel, k, r.norm().item()))
for rel, k, nr in sorted(rows, reverse=True)[:60]:
    print(f"{rel:10.3e}  |ref|={nr:10.3e}  {k}")
