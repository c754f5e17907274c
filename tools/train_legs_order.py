"""Does the order of the timed legs change their time? forward_backward / full step of ONE TrainStep, alternating. Profiling aid."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import synth, train_glue as G, train_path as T
dev = "cuda"
m = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26); synth.load_seeded(m, 0); m = m.to(dev).train()
bs = [synth.make_config_batch("cfg2", seed=1 + i).to(dev) for i in range(4)]
g = torch.Generator().manual_seed(3)
def tg(b):
    n, e = b.obj_points.shape[0], b.edge_indices.shape[1]
    t = torch.randn(e, 512, generator=g)
    return (torch.randint(0, 160, (n,), generator=g).to(dev), (torch.rand(e, 26, generator=g) < 1 / 26).float().to(dev), (t / t.norm(dim=-1, keepdim=True)).to(dev))
tgs = [tg(b) for b in bs]
st = [T.scene_stats(b.batch_ids) for b in bs]
ts = G.TrainStep(m, G.build_optimizer(m, 1e-4, max_iteration=1000))
flush = torch.empty(64 << 20, device=dev)
def timed(fn, label, n=10):
    for i in range(2): fn(i)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i in range(n):
        flush.zero_(); ev[i][0].record(); fn(i); ev[i][1].record()
    torch.cuda.synchronize()
    t = [a.elapsed_time(b) for a, b in ev]
    print(f"{label}: mean {sum(t)/n:.2f} ms  min {min(t):.2f}  max {max(t):.2f}", flush=True)
fb = lambda i: ts.forward_backward(*bs[i % 4].forward_args(), *tgs[i % 4], scene_stats=st[i % 4])
full = lambda i: ts.step(*bs[i % 4].forward_args(), *tgs[i % 4], scene_stats=st[i % 4])
full(0)
for rep in range(2):
    timed(fb, "forward_backward")
    timed(full, "full step       ")
fb1 = lambda i: ts.forward_backward(*bs[0].forward_args(), *tgs[0], scene_stats=st[0])
timed(fb1, "forward_backward, same batch every step")
