"""Device time of the full training step (TrainStep: forward, reference losses, backward, fused AdamW) against the
fixed-cotangent forward+backward graph, in both orders, with the SM clock sampled after each phase; then the per-entry
device time of the N1 kernels from an eager step. Profiling aid."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops, synth, train_glue as G, train_path as T

import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
clk = lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)

dev = "cuda"
b = synth.make_config_batch("cfg2", seed=1).to(dev)
g = torch.Generator().manual_seed(3)
n, e = b.obj_points.shape[0], b.edge_indices.shape[1]
text = torch.randn(e, 512, generator=g)
targets = (torch.randint(0, 160, (n,), generator=g).to(dev), (torch.rand(e, 26, generator=g) < 1 / 26).float().to(dev),
           (text / text.norm(dim=-1, keepdim=True)).to(dev))
stats = T.scene_stats(b.batch_ids)
flush = torch.empty(64 << 20, device=dev, dtype=torch.float32)


def new_model():
    m = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)
    synth.load_seeded(m, 0)
    return m.to(dev).train()


def timed(fn, reps=10, label=""):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, c in ev:
        flush.zero_()
        a.record(); fn(); c.record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(c) for a, c in ev]
    print(f"{label}: mean {sum(ts) / reps:.2f} ms  first {ts[0]:.2f}  last {ts[-1]:.2f}  sm clock after {clk()} MHz", flush=True)


model = new_model()
ts = G.TrainStep(model, G.build_optimizer(model, 1e-4, max_iteration=1000))
full = lambda: ts.step(*b.forward_args(), *targets, scene_stats=stats)
cot = None
def loss_fn(outs):
    global cot
    if cot is None:
        gen = torch.Generator(device=dev).manual_seed(5)
        cot = [torch.randn(o.shape, device=dev, generator=gen) / o.numel() for o in outs[:7]]
    return sum((o * c).sum() for o, c in zip(outs[:7], cot))
model2 = new_model()
loss_fn(model2(*b.forward_args(), istrain=True)).backward()
model2.zero_grad(set_to_none=True)
gts = V.GraphedTrainStep(model2, loss_fn)
plain = lambda: gts(*b.forward_args(), scene_stats=stats)

timed(full, label="train_step (losses + AdamW)   ")
timed(plain, label="fwd+bwd fixed cotangents      ")
timed(full, label="train_step again              ")
timed(plain, label="fwd+bwd again                 ")
opt_only = lambda: ts.optimizer.step()
timed(opt_only, reps=20, label="AdamW step alone              ")

m3 = new_model()
ts3 = G.TrainStep(m3, G.build_optimizer(m3, 1e-4, max_iteration=1000), graphed=False)
for _ in range(2):
    ts3.step(*b.forward_args(), *targets)
torch.cuda.synchronize()
timer = ops.KernelTimer(); ops.set_timer(timer)
torch.cuda._sleep(400_000_000)
ts3.step(*b.forward_args(), *targets)
torch.cuda.synchronize()
ops.set_timer(None)
summ = timer.summary()
print(f"eager step, sum of C-ABI device time: {sum(d['ms'] for d in summ.values()):.2f} ms")
for name, d in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
    if any(k in name for k in ("cross_entropy", "bce", "cosine", "l1_unit", "sum_rows", "adamw", "rel_class")):
        print(f"{d['ms']:9.3f} ms  {d['launches']:5d} calls  {d['ms'] / d['launches'] * 1e3:9.1f} us/call  {name}")
