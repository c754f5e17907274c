"""One eager training step (TrainStep.forward_backward: train-mode forward, the reference's six loss terms, backward) at config #2
after two warm-up steps: the command to put under `ncu --profile-from-start off --metrics gpu__time_duration.sum` for a
per-kernel launch list of the training path as bench.py times it (zero arena, pair hand-offs). Profiling aid."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import autograd as A, synth, train_glue as G
import torch.cuda.profiler as prof

dev = "cuda"
model = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)
synth.load_seeded(model, 0)
model = model.to(dev).train()
b = synth.make_config_batch("cfg2", seed=1).to(dev)
gen = torch.Generator().manual_seed(77)
n, e = b.obj_points.shape[0], b.edge_indices.shape[1]
text = torch.randn(e, 512, generator=gen)
targets = (torch.randint(0, 160, (n,), generator=gen).to(dev), (torch.rand(e, 26, generator=gen) < 1.0 / 26).float().to(dev),
           (text / text.norm(dim=-1, keepdim=True)).to(dev))
A.DropoutState.manual_seed(1234)
ts = G.TrainStep(model, G.build_optimizer(model, lr=1e-4, max_iteration=1000), graphed=False)


def step():
    model.zero_grad(set_to_none=True)
    ts.forward_backward(*b.forward_args(), *targets)


for _ in range(2):
    step()
torch.cuda.synchronize()
prof.start()
step()
torch.cuda.synchronize()
prof.stop()
print("done")
