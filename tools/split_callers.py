"""Who asks for a standalone bf16 split pass during one eager training step at config #2: shapes and the two nearest callers
of every ``ops.bf16_split`` call (activations whose producer did not emit the pair). Profiling aid."""
import collections, os, sys, traceback, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops, synth

dev = "cuda"
model = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)
synth.load_seeded(model, 0)
model = model.to(dev).train()
b = synth.make_config_batch("cfg2", seed=1).to(dev)
cot = None


def step():
    global cot
    model.zero_grad(set_to_none=True)
    outs = model(*b.forward_args(), istrain=True)
    if cot is None:
        cot = [torch.randn_like(o) / o.numel() for o in outs[:7]]
    sum((o * c).sum() for o, c in zip(outs[:7], cot)).backward()


for _ in range(2):
    step()
calls = collections.Counter()
orig = ops.bf16_split


def traced(x):
    fr = traceback.extract_stack(limit=7)[:-1]
    where = " < ".join(f"{os.path.basename(f.filename)}:{f.lineno}:{f.name}" for f in reversed(fr) if "vlsat" in f.filename or "tools" in f.filename)[:300]
    calls[(tuple(x.shape), where)] += 1
    return orig(x)


ops.bf16_split = traced
step()
torch.cuda.synchronize()
for (shape, where), n in sorted(calls.items(), key=lambda kv: -kv[0][0][0] * kv[0][0][1] * kv[1]):
    print(n, shape, where)
print("total", sum(calls.values()))
