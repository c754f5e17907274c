"""Device time of every vlsat_linear_fwd call of one cfg2 forward (eval), largest first. Profiling aid."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops, synth
dev = "cuda"
model = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)
synth.load_seeded(model, 0)
model = model.to(dev).eval()
b = synth.make_config_batch(sys.argv[1] if len(sys.argv) > 1 else "cfg2", seed=1).to(dev)
with torch.no_grad():
    for _ in range(3):
        model(*b.forward_args(), istrain=False)
    torch.cuda.synchronize()
    timer = ops.KernelTimer(); ops.set_timer(timer)
    torch.cuda._sleep(400_000_000)
    model(*b.forward_args(), istrain=False)
    torch.cuda.synchronize()
ops.set_timer(None)
tot = {}
rows = []
for name, a, c, fl, by in timer.events:
    us = a.elapsed_time(c) * 1e3
    tot[name] = tot.get(name, 0.0) + us
    if name == "vlsat_linear_fwd":
        rows.append((us, fl, by))
print({k: round(v, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])})
print("linear calls:", len(rows), "total us", round(sum(r[0] for r in rows), 1))
import collections
agg = collections.OrderedDict()
for us, fl, by in rows:
    key = (fl, by)
    d = agg.setdefault(key, [0, 0.0]); d[0] += 1; d[1] += us
for (fl, by), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  GFLOP {fl / 1e9:7.3f}  MB {by / 1e6:7.2f}  x{n:2d}  {us / n:7.1f} us each  {us:8.1f} us total  {fl * n / us * 1e-6:7.1f} TFLOP/s")
