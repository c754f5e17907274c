"""One eager cfg2 forward between cudaProfilerStart/Stop, for `ncu --profile-from-start off` captures
(launch list or --set full of selected kernels). Not a timing tool: numbers under a profiler are never bench values."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import synth
dev = "cuda"
model = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)
synth.load_seeded(model, 0)
model = model.to(dev).eval()
b = synth.make_config_batch(sys.argv[1] if len(sys.argv) > 1 else "cfg2", seed=1).to(dev)
with torch.no_grad():
    for _ in range(3):
        model(*b.forward_args(), istrain=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(*b.forward_args(), istrain=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled one forward")
