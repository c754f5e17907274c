"""Per-entry-point device time of one training step (forward + backward) at BASELINE config #2. Debug/profiling aid."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops, synth, autograd as A

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
dev = "cuda"
model = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)
synth.load_seeded(model, 0)
model = model.to(dev).train()
b = synth.make_config_batch(cfg, seed=1).to(dev)
cot = None

def step():
    global cot
    model.zero_grad(set_to_none=True)
    outs = model(*b.forward_args(), istrain=True)
    if cot is None:
        cot = [torch.randn_like(o) / o.numel() for o in outs[:7]]
    sum((o * c).sum() for o, c in zip(outs[:7], cot)).backward()

for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"eager step wall: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms, peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
timer = ops.KernelTimer(); ops.set_timer(timer)
n = 3
for _ in range(n):
    torch.cuda._sleep(200_000_000)
    step()
torch.cuda.synchronize()
ops.set_timer(None)
summ = timer.summary()
tot = sum(d["ms"] for d in summ.values())
print(f"sum of C-ABI device time per step: {tot / n:.2f} ms")
for name, d in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{d['ms'] / n:9.3f} ms  {d['launches'] / n:7.1f} calls  {d['ms'] / d['launches'] * 1e3:9.1f} us/call  {name}")
# largest individual calls
ev = sorted(timer.events, key=lambda e: -e[1].elapsed_time(e[2]))[:25]
for name, a, b_, fl, by in ev:
    print(f"   {a.elapsed_time(b_) * 1e3:9.1f} us  {name}  flops {fl:.3g}")
