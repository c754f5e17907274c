"""Run each hot kernel of the path a few times at BASELINE config #2 shapes (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops, synth

which = sys.argv[1:] or ["linear", "flash", "gat", "pointnet"]
dev = "cuda"
g = torch.Generator().manual_seed(0)
n_iter = 3
if "linear" in which:
    for (m, n, k) in [(9600, 512, 512), (640, 512, 512), (9600, 1024, 512)]:
        x, w, b = torch.randn(m, k, generator=g).to(dev), torch.randn(n, k, generator=g).to(dev), torch.randn(n, generator=g).to(dev)
        for _ in range(n_iter):
            ops.linear(x, w, b, act=1)
if "flash" in which:
    q, k = torch.randn(9600, 512, generator=g).to(dev), torch.randn(9600, 512, generator=g).to(dev)
    vt = torch.randn(512, 9600, generator=g).to(dev)
    for _ in range(n_iter):
        ops.flash_attn_tc(q, k, vt, 9600, 8)
if "gat" in which or "pointnet" in which:
    model = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)
    synth.load_seeded(model, 0)
    model = model.to(dev).eval()
    b = synth.make_config_batch("cfg2", seed=1).to(dev)
    with torch.no_grad():
        for _ in range(n_iter):
            if "pointnet" in which:
                model.obj_encoder(b.obj_points)
            if "gat" in which:
                layer = model.mmg.gcn_3ds[0]
                x = torch.randn(640, 512, device=dev); e = torch.randn(9600, 512, device=dev)
                layer(x, e, b.edge_indices)
torch.cuda.synchronize()
print("done")
