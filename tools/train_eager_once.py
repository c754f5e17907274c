"""One eager training step (forward + backward of a fixed-cotangent scalar) at config #2 after two warm-up steps: the command
to put under `ncu --metrics gpu__time_duration.sum` for a per-kernel launch list of the training path. Profiling aid."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import synth
import torch.cuda.profiler as prof

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
torch.cuda.synchronize()
prof.start()
step()
torch.cuda.synchronize()
prof.stop()
print("done")
